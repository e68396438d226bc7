"""move_l: Cartesian trajectory tracking with the operational-space controller pid_task_ctrl
(reference controller/move_l_task.py:55-69, move_l_mug.py:67-81, move_l_hold.py, move_l_point.py; controller_func.py:68-117).
`run_pinv` is the pseudo-inverse IK + joint-PD controller of controller/move_l.py:15-78 (gains controller/config/config_l.yml)."""
from .. import _lib, presets
from .loops import run_trajectory


def run(traj, n_envs=1, xml="ur3e_2f85.xml", gains=presets.GAINS_L_TASK, keyframe="down", **kw):
    """traj [T, 7] = x, y, z, rx, ry, rz, g (reference CSV layout).  xml='main.xml' + gains=GAINS_L_MUG gives move_l_mug."""
    return run_trajectory(xml, _lib.CTRL_PID_TASK, gains, traj, n_envs, keyframe, **kw)


def run_pinv(traj, n_envs=1, xml="ur3e_2f85.xml", gains=presets.GAINS_L_PINV, keyframe="down", **kw):
    """controller/move_l.py main loop: traj [T, 7] = x, y, z, rx, ry, rz, g."""
    return run_trajectory(xml, _lib.CTRL_PINV, gains, traj, n_envs, keyframe, **kw)
