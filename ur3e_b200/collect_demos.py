"""Batched expert-demonstration collection: the reference's gymnasium_src/scripts/imitation_rl/collect_demos.py:86-209 for N
demonstrations at once on the GPU.

The reference runs, per demonstration, `reset_with_mug(stochastic, 'down', noise)` -> `build_traj_l_pick_place_imitation_augmented`
-> `for t: u = pid_task_ctrl(traj[t]); d.ctrl = u; mj_step; record`, and stores `imitation.data.types.Trajectory(obs, acts, infos,
terminal=True)` objects in a pickle.  Here the N demonstrations are N environments of one batch (controller fused in the kernel, one
launch per step), the trajectories are built on the device per environment, and the result has the same layout:

  obs  [N, T / down_sample + 1, 24]   the 24-dim observation of ImitationEnvIndirect (collect_demos.py:60-84), initial state first
  acts [N, T / down_sample, 4 | 7]    "indirect": [x, y, z of the target, grip ctrl = u[-1]];  "direct": u (7 actuator commands)
  infos, terminal=True                as the reference

`to_trajectories()` yields `imitation.data.types.Trajectory` objects when `imitation` imports, else equivalent light records;
`save_demos` / `load_demos` keep the reference's pickle-of-a-list format (append when `resume_collecting`).
"""
import os
import pickle as pkl
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, presets
from .batch import SimBatch
from .controller import build_traj as BT
from .model import Model, asset


@dataclass
class Trajectory:
    """Field-compatible stand-in for imitation.data.types.Trajectory (used when `imitation` is not installed)."""
    obs: np.ndarray
    acts: np.ndarray
    infos: np.ndarray
    terminal: bool

    def __len__(self):
        return len(self.acts)


def collect_expert_demonstrations(num_demos, action_mode="indirect", reset_mode="stochastic", noise_mag="low", down_sample=1, seed=0, device=0,
                                  dtype=torch.float32, hold=120):
    """Returns dict(obs [N, K + 1, 24], acts [N, K, A], terminal=True) as torch tensors on the device (K = T // down_sample kept steps)."""
    if action_mode not in ("indirect", "direct"):
        raise ValueError("action_mode must be 'indirect' or 'direct'")
    model = Model(asset("main.xml"))
    noise = {"low": _lib.NOISE_LOW, "med": _lib.NOISE_MED, "high": _lib.NOISE_HIGH}[noise_mag] if reset_mode == "stochastic" else _lib.NOISE_NONE
    cfg = presets.make_config(model, dict(ctrl_mode=_lib.CTRL_PID_TASK, obs_kind=_lib.OBS_V2, obs_dim=24, act_dim=7, frame_skip=1, gains=presets.GAINS_L_MUG,
                                          reset_key="down", reset_noise=noise, term_kind=_lib.TERM_NONE, reward_kind=_lib.REW_NONE, max_steps=0))
    batch = SimBatch(model, cfg, num_demos, device, dtype)
    obs0 = batch.reset(seed=seed).clone()
    sens = batch.enable_sensors()
    kw = dict(device=batch.device, dtype=torch.float64)
    rot = torch.tensor(presets.TOOL_ROTVEC, **kw).expand(num_demos, 3)            # init_r = get_site_xrotvec(tcp) at keyframe 'down'
    o = obs0.double()
    start = torch.cat([o[:, 0:3], rot, torch.zeros(num_demos, 1, **kw)], 1)       # get_task_space_state: [xpos, xrotvec, grasp bool]
    pick = torch.cat([o[:, 3:6], rot, torch.zeros(num_demos, 1, **kw)], 1)        # collect_demos.py:112
    place = torch.cat([o[:, 6:9], rot, torch.ones(num_demos, 1, **kw)], 1)        # collect_demos.py:109
    g = torch.Generator(device=batch.device); g.manual_seed(seed)
    traj = BT.build_traj_l_pick_place_imitation_augmented(start, [pick, place], hold, device=batch.device, generator=g)     # [T, N, 7]
    T = traj.shape[0]
    grip_hi = float(model.actuator_ctrlrange[-1][1])
    obs, acts = [obs0.clone()], []
    for t in range(T):
        row = traj[t].to(dtype).contiguous()
        ob, *_ = batch.step(row, want_final_obs=False)
        if t % down_sample == 0:
            if action_mode == "indirect":
                acts.append(torch.cat([row[:, 0:3], row[:, 6:7] * grip_hi], 1))   # [x, y, z, u[-1]] (collect_demos.py:143-149)
            else:
                acts.append(sens[:, 21:28].clone())                               # u of pid_task_ctrl (collect_demos.py:150-151)
            obs.append(ob.clone())
    return dict(obs=torch.stack(obs, 1), acts=torch.stack(acts, 1), terminal=True, traj=traj)


def to_trajectories(demos):
    """List of N Trajectory objects (imitation's class when importable) with float64 numpy arrays, like the reference's pickle."""
    try:
        from imitation.data.types import Trajectory as T
    except Exception:
        T = Trajectory
    obs, acts = demos["obs"].double().cpu().numpy(), demos["acts"].double().cpu().numpy()
    return [T(obs=obs[i], acts=acts[i], infos=np.array([{} for _ in range(acts.shape[1])]), terminal=True) for i in range(obs.shape[0])]


def save_demos(trajectories, save_path, resume_collecting=False):
    """collect_demos.py:192-209: extend the list stored at `save_path` (when resuming) and overwrite the pickle."""
    os.makedirs(os.path.dirname(save_path) or ".", exist_ok=True)
    existing = []
    if os.path.exists(save_path) and resume_collecting:
        with open(save_path, "rb") as f:
            existing = pkl.load(f)
    existing.extend(trajectories)
    with open(save_path, "wb") as f:
        pkl.dump(existing, f)
    return len(existing)


def load_demos(load_fpath):
    with open(load_fpath, "rb") as f:
        return pkl.load(f)
