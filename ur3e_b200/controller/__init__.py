"""Batched controller entry points with the reference's names (reference controller/move_j.py, move_l_task.py,
move_l_mug.py, task_space.py, controller_func.py, build_traj.py).  One controller evaluation per mj_step, fused in the kernel."""
from . import build_traj, move_j, move_l, task_space  # noqa: F401
from .loops import run_trajectory  # noqa: F401
