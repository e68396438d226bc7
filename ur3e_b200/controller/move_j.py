"""move_j: joint-space PD trajectory tracking (reference controller/move_j.py:14-38,76-86;
controller_func.py:128-167 pd_joint_ctrl).  ctrl = clip(Kp (clip(q*, jnt_range) - q) - Kd qdot, ctrlrange), grip = g * 255."""
from .. import _lib, presets
from .loops import run_trajectory


def run(traj, n_envs=1, xml="ur3e_2f85.xml", gains=presets.GAINS_J, keyframe="down", **kw):
    """traj [T, 7] = j1..j6, g (reference CSV layout, build_traj.py:515-525); ur3e_raw.xml takes the first 6 columns."""
    if xml == "ur3e_raw.xml":
        keyframe = None
    return run_trajectory(xml, _lib.CTRL_PD_JOINT, gains, traj, n_envs, keyframe, **kw)
