"""Shared batched loop: `for t in range(T): u = ctrl(traj[t]); d.ctrl = u; mj_step` (reference controller/move_j.py:76-86,
move_l_task.py:55-69, move_l_mug.py:67-81) for N environments at once, the controller fused into the step kernel."""
import numpy as np
import torch

from .. import _lib, presets
from ..batch import SimBatch
from ..model import Model, asset


def run_trajectory(xml, ctrl_mode, gains, traj, n_envs=1, keyframe="down", device=0, dtype=torch.float32, record_every=1, batch=None):
    """traj: [T, A] (shared by all envs) or [T, N, A] tensor/array of per-step targets.
    Returns (qpos [R, N, nq], qvel [R, N, nv]) recorded every `record_every` steps, and the SimBatch (for further stepping)."""
    model = Model(xml if "/" in xml else asset(xml))
    act_dim = {_lib.CTRL_PD_JOINT: 7 if model.nu > 6 else 6, _lib.CTRL_PID_TASK: 7, _lib.CTRL_PINV: 7, _lib.CTRL_RAW: model.nu}[ctrl_mode]
    if batch is None:
        key = model.key_id(keyframe) if (keyframe is not None and model.nkey) else -1
        cfg = presets.make_config(model, dict(ctrl_mode=ctrl_mode, obs_kind=_lib.OBS_STATE, obs_dim=model.nq + model.nv, act_dim=act_dim,
                                              frame_skip=1, gains=gains, reset_key=key))
        batch = SimBatch(model, cfg, n_envs, device, dtype)
        batch.reset()
    traj = torch.as_tensor(np.asarray(traj) if not isinstance(traj, torch.Tensor) else traj, dtype=dtype, device=batch.device)
    if traj.dim() == 2:
        traj = traj[:, None, :].expand(-1, n_envs, -1)
    if traj.shape[2] < act_dim:
        raise ValueError("trajectory has %d columns, controller needs %d" % (traj.shape[2], act_dim))
    T = traj.shape[0]
    qs, vs = [], []
    nq = model.nq
    for t in range(T):
        obs, *_ = batch.step(traj[t, :, :act_dim].contiguous(), want_final_obs=False)
        if (t + 1) % record_every == 0:
            qs.append(obs[:, :nq].clone()); vs.append(obs[:, nq:].clone())
    return torch.stack(qs), torch.stack(vs), batch
