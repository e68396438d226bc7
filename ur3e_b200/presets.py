"""Environment presets: the constants the reference hard-codes in its env classes and controller configs.

ids follow the reference's registration (gymnasium_env/envs/register_envs.py:4-25).
"""
import numpy as np

from . import _lib
from .batch import env_config

TOOL_ROTVEC = (-1.209, -1.209, 1.209)                                   # ur3e_env2.py:74 / ur3e_env.py:155
GAINS_L_MUG = [220, 220, 120, 20, 20, 40, 35, 15, 15, 2, 2, 2]          # controller/config/config_l_mug.yml (kp_pos kd_pos kp_rot kd_rot)
GAINS_V0 = [320, 320, 320, 20, 20, 25, 325, 325, 325, 2, 2, 2]          # ur3e_env.py:106-117
GAINS_L_TASK = [120, 120, 120, 20, 20, 20, 35, 15, 15, 2, 2, 2]         # controller/config/config_l_task.yml
GAINS_J = [20, 380, 300, 20, 30, 10, 5, 5, 5, 5, 5, 5]                  # controller/config/config_j.yml (kp[6] kd[6])
GAINS_L_PINV = [20, 60, 20, 20, 20, 10, 5, 15, 5, 5, 5, 20, 5.02, 5.01, 5.80, 5.80, 5.09, 5.80, 5, 50, 10, 5, 5, 5]   # controller/config/config_l.yml (kp_pos kd_pos kp_rot kd_rot, per joint)
MUG_DOWN_XY = (0.29799994, 0.13349916)                                  # assets/main.xml keyframe 'down'

ENV_SPECS = {
    # id: (xml, config kwargs, action low, action high)
    "gymnasium_env/ur3e-v2": ("main.xml", dict(ctrl_mode=_lib.CTRL_PID_TASK_ENV, obs_kind=_lib.OBS_V2, obs_dim=24, act_dim=4, frame_skip=2,
                                                 term_kind=_lib.TERM_V2, reward_kind=_lib.REW_V2, max_steps=2500, gains=GAINS_L_MUG, reset_key="down",
                                                 reset_noise=_lib.NOISE_HIGH),
                              [MUG_DOWN_XY[0] - 0.25, MUG_DOWN_XY[1] - 0.25, 0.0, 0.0], [MUG_DOWN_XY[0] + 0.25, MUG_DOWN_XY[1] + 0.25, 0.5, 1.0]),   # ur3e_env2.py:57-64
    "gymnasium_env/ur3e-v0": ("main.xml", dict(ctrl_mode=_lib.CTRL_PID_TASK_ENV, obs_kind=_lib.OBS_V0, obs_dim=13, act_dim=4, frame_skip=2,
                                                 term_kind=_lib.TERM_V0, reward_kind=_lib.REW_V0, max_steps=500, gains=GAINS_V0, reset_key="down",
                                                 reset_noise=_lib.NOISE_HIGH),
                              [0.28799994, 0.13349916, 0.005, 0.0], [0.35799994, 0.35349916, 0.165, 1.0]),                                            # ur3e_env.py:94-95
    "gymnasium_env/imitation_indirect-v0": ("main.xml", dict(ctrl_mode=_lib.CTRL_PID_TASK_ENV, obs_kind=_lib.OBS_V2, obs_dim=24, act_dim=4, frame_skip=1,
                                                               term_kind=_lib.TERM_NONE, reward_kind=_lib.REW_MINUS1, max_steps=2500, gains=GAINS_L_MUG,
                                                               reset_key="down", reset_noise=_lib.NOISE_HIGH),
                                            [MUG_DOWN_XY[0] - 0.25, MUG_DOWN_XY[1] - 0.25, 0.0, 0.0], [MUG_DOWN_XY[0] + 0.25, MUG_DOWN_XY[1] + 0.25, 0.5, 1.0]),
    "gymnasium_env/imitation_direct-v0": ("main.xml", dict(ctrl_mode=_lib.CTRL_RAW, obs_kind=_lib.OBS_DIRECT, obs_dim=13, act_dim=7, frame_skip=2,
                                                             term_kind=_lib.TERM_NONE, reward_kind=_lib.REW_MINUS1, max_steps=1200, reset_key="down",
                                                             reset_noise=_lib.NOISE_HIGH),
                                          None, None),                                                                                                 # actuator_ctrlrange
}


def make_config(model, spec_kwargs, **override):
    kw = dict(spec_kwargs); kw.update(override)
    key = kw.get("reset_key", -1)
    if isinstance(key, str):
        kw["reset_key"] = model.key_id(key)
    return env_config(**kw)


def action_bounds(model, env_id):
    _, _, lo, hi = ENV_SPECS[env_id]
    if lo is None:
        r = np.array(model.actuator_ctrlrange)
        return r[:, 0].copy(), r[:, 1].copy()
    return np.array(lo, dtype=np.float64), np.array(hi, dtype=np.float64)
