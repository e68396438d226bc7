"""task_space: the reference's controller/task_space.py is a fully commented-out predecessor of pid_task_ctrl
(SURVEY F8c); this module keeps the entry-point name and maps it onto the live controller semantics:
tau = J^T [Kp e_p - Kd Jp qdot ; Kpr e_r - Kdr Jr qdot] + qfrc_bias[:6], no clipping, no integral term, no task-space
mass matrix (controller_func.py:98-117)."""
from .move_l import run  # noqa: F401

pid_task_ctrl = run
