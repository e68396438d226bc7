"""Batched counterparts of the reference's state readers (utils/utils.py, controller/controller_func.py): same names, the
`d` argument is a SimBatch (or a UR3eVecEnv) instead of an MjData, and every value gains a leading environment axis.

    env = UR3eVecEnv("gymnasium_env/ur3e-v2", 4096); env.batch.enable_sensors(); env.reset(); env.step(a)
    get_jnt_torques(env)              # [N, 7]   utils/utils.py:201-211 (the seven actuatorfrc sensors)
    get_grasp_contact(env)            # ([N], [N])  utils/utils.py:243-245 (left, right touch sensor)
    get_boolean_grasp_contact(env)    # [N] bool    utils/utils.py:238-240
    get_task_space_state(env)         # [N, 7]   controller_func.py:191-200: tcp xpos, tcp rotvec, grasp flag
    get_joint_space_state(env)        # [N, 7]   controller_func.py:203-211: qpos[:6], grasp flag
    get_ctrl(env)                     # [N, 7]   d.ctrl as the fused controller set it (the `u` of pid_task_ctrl / move_j.ctrl)
    get_torque_sensors(env)           # [N, 6, 3]  the <torque> site sensors of assets/main.xml:384-391 (shoulder_pan .. wrist_3)

The sensors are the ones of the step's last mj_step (MuJoCo fills d.sensordata in the forward pass, before integrating),
so a call after `step` returns what the reference's loops log after `mj_step`.  Environments that were auto-reset in that
step report the sensors of their terminal step.
"""
import math

import torch


def _batch(d):
    return getattr(d, "batch", d)


def _sensors(d):
    b = _batch(d)
    if getattr(b, "sensors", None) is None:
        raise RuntimeError("sensor output is off: call batch.enable_sensors() before stepping")
    return b.sensors


def get_jnt_torques(d):
    return _sensors(d)[:, :7]


get_joint_torques = get_jnt_torques   # the name controller/move_*.py import


def get_ctrl(d):
    """d.ctrl of the step's last mj_step, before MuJoCo's ctrlrange clamp (what the reference's loops store in `ctrls[t]`)."""
    return _sensors(d)[:, 21:28]


def get_torque_sensors(d):
    """d.sensor("<joint>_torque").data for the six arm joints: interaction torque between each link and its parent at the joint's
    force-torque site, in the site frame."""
    return _sensors(d)[:, 28:46].reshape(-1, 6, 3)


def get_grasp_contact(d):
    s = _sensors(d)
    return s[:, 8], s[:, 7]           # (left_pad1_contact, right_pad1_contact), the reference's order


def get_boolean_grasp_contact(d, contact_threshold=0.1):
    # the reference compares the TUPLE (left, right) > (thr, thr): lexicographic, i.e. left > thr, or left == thr and right > thr
    left, right = get_grasp_contact(d)
    return (left > contact_threshold) | ((left == contact_threshold) & (right > contact_threshold))


def _rotvec(R):
    """scipy Rotation.from_matrix(R).as_rotvec() for a batch of rotation matrices [N, 3, 3] (angle in [0, pi])."""
    t = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    # quaternion (w, x, y, z) by the largest-component rule, then canonical w >= 0
    q = torch.empty(R.shape[0], 4, dtype=R.dtype, device=R.device)
    c = torch.stack([t, R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]], 1).argmax(1)
    for k in range(4):
        m = c == k
        if not m.any():
            continue
        r = R[m]
        if k == 0:
            w = 1 + r[:, 0, 0] + r[:, 1, 1] + r[:, 2, 2]
            qq = torch.stack([w, r[:, 2, 1] - r[:, 1, 2], r[:, 0, 2] - r[:, 2, 0], r[:, 1, 0] - r[:, 0, 1]], 1)
        else:
            i = k - 1; j = (i + 1) % 3; l = (i + 2) % 3
            v = [None] * 4
            v[1 + i] = 1 + r[:, i, i] - r[:, j, j] - r[:, l, l]
            v[1 + j] = r[:, j, i] + r[:, i, j]
            v[1 + l] = r[:, l, i] + r[:, i, l]
            v[0] = r[:, l, j] - r[:, j, l]
            qq = torch.stack(v, 1)
        q[m] = qq / qq.norm(dim=1, keepdim=True)
    q = torch.where(q[:, :1] < 0, -q, q)
    sn = q[:, 1:].norm(dim=1)
    ang = 2 * torch.atan2(sn, q[:, 0])
    k = torch.where(sn < 1e-12, torch.full_like(sn, 2.0), ang / sn.clamp_min(1e-300))
    return q[:, 1:] * k[:, None]


def get_task_space_state(d):
    """[tcp xpos (3), tcp rotvec (3), grasp flag] per environment (controller_func.py:191-200).  The tcp pose is the one of the
    step's last forward pass, i.e. the pose the next controller call will read."""
    s = _sensors(d)
    pos, mat = s[:, 9:12], s[:, 12:21].reshape(-1, 3, 3)
    g = get_boolean_grasp_contact(d).to(pos.dtype)
    return torch.cat([pos, _rotvec(mat), g[:, None]], 1)


def get_joint_space_state(d):
    """[qpos[:6], grasp flag] per environment (controller_func.py:203-211)."""
    b = _batch(d)
    qpos, _, _ = b.get_state()
    return torch.cat([qpos[:, :6], get_boolean_grasp_contact(d).to(qpos.dtype)[:, None]], 1)
