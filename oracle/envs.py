"""ORACLE (test infrastructure, NOT product code): single-environment float64 restatement of the
reference's Gymnasium envs and controller loops on top of oracle/oracle.py.

Follows, line by line, gymnasium_env/envs/ur3e_env2.py:72-123,150-261, ur3e_env.py:137-200,242-462,
imitation_env_indirect.py:73-150, imitation_env_direct.py:74-139, utils/gym_utils.py:8-128,174-201 and
gymnasium's MujocoEnv.do_simulation/set_state.  Stale-kinematics semantics
come for free: the oracle's step() = forward + integrate, exactly like mj_step.
Parity status: UNPINNED against MuJoCo (see ur3e_oracle.h).
"""
import numpy as np

from . import oracle as O

TOOL_ROTVEC = np.array([-1.209, -1.209, 1.209])            # ur3e_env2.py:74
GAINS_MUG = np.array([220, 220, 120, 20, 20, 40, 35, 15, 15, 2, 2, 2], dtype=float)     # controller/config/config_l_mug.yml
GAINS_V0 = np.array([320, 320, 320, 20, 20, 25, 325, 325, 325, 2, 2, 2], dtype=float)   # ur3e_env.py:106-117
GAINS_TASK = np.array([120, 120, 120, 20, 20, 20, 35, 15, 15, 2, 2, 2], dtype=float)    # controller/config/config_l_task.yml
GAINS_J = np.array([20, 380, 300, 20, 30, 10, 5, 5, 5, 5, 5, 5], dtype=float)           # controller/config/config_j.yml


class OracleEnv:
    """kind in {'v2', 'v0', 'indirect', 'direct'}; model = main.xml."""

    def __init__(self, xml, kind="v2"):
        self.m = O.Model(xml); self.d = O.Data(self.m); self.kind = kind
        m = self.m
        self.tcp = m.id("site", "tcp"); self.handle = m.id("site", "handle_site"); self.pad_site = m.id("site", "right_pad1_site")
        self.fish = m.id("body", "fish"); self.ghost = m.id("body", "ghost"); self.lpad = m.id("body", "left_pad"); self.rpad = m.id("body", "right_pad")
        self.table = m.id("body", "table")
        root = m.id("body", "robotiq_base_mount")
        par = m.py["body_parentid"]
        self.gripper = {b for b in range(m.nbody) if self._is_desc(b, root, par)}
        self.arm = {b for b in range(m.nbody) if self._is_desc(b, m.id("body", "robot_base"), par)}      # gym_utils.py:133-143 init_collision_cache
        self.frame_skip = {"v2": 2, "v0": 2, "indirect": 1, "direct": 2}[kind]
        self.max_steps = {"v2": 2500, "v0": 500, "indirect": 2500, "direct": 1200}[kind]
        self.gains = GAINS_V0 if kind == "v0" else GAINS_MUG
        g = [i for i in range(m.ngeom) if m.py["geom_bodyid"][i] == self.fish][0]
        self.mug_size = m.py["geom_size"][g]
        self.t = 0

    @staticmethod
    def _is_desc(b, root, par):
        while b > 0:
            if b == root:
                return True
            b = par[b]
        return False

    def set_state(self, qpos, qvel):
        d = self.d
        d.reset(); d.set_state(qpos, qvel); d.forward(); self.t = 0

    def reset(self, noise_xy=(0.0, 0.0)):
        qp, qv = self.m.key("down")
        qp[14] += noise_xy[0]; qp[15] += noise_xy[1]
        self.set_state(qp, qv)
        return self.obs()

    # ---- gym_utils helpers
    def grasp_count(self):
        pads = set()
        gb = self.m.py["geom_bodyid"]
        for c in self.d.contacts():
            b1, b2 = gb[c.geom1], gb[c.geom2]
            for pad in (self.lpad, self.rpad):
                if pad in (b1, b2) and self.fish in (b1, b2):
                    pads.add(pad)
        return len(pads)

    def table_collision(self):
        gb = self.m.py["geom_bodyid"]
        for c in self.d.contacts():
            b1, b2 = gb[c.geom1], gb[c.geom2]
            if (b1 in self.gripper and b2 == self.table) or (b2 in self.gripper and b1 == self.table):
                return 1
        return 0

    def self_collision(self):
        """gym_utils.py:146-172: two bodies of the robot_base subtree in contact, unless both belong to the gripper subtree."""
        gb = self.m.py["geom_bodyid"]
        for c in self.d.contacts():
            b1, b2 = gb[c.geom1], gb[c.geom2]
            if b1 in self.arm and b2 in self.arm and not (b1 in self.gripper and b2 in self.gripper):
                return 1
        return 0

    def obs(self):
        d = self.d
        sx = d.site_xpos.reshape(-1, 3)
        tcp, mug, ghost = sx[self.tcp].copy(), sx[self.handle].copy(), d.xpos.reshape(-1, 3)[self.ghost].copy()
        if self.kind in ("v2", "indirect"):
            vt = d.site_velocity(self.tcp)[3:]; vm = d.site_velocity(self.handle)[3:]
            gs = self.grasp_count()
            robust = 1 if (gs == 2 and np.all(np.abs(tcp - mug) < [0.01, 0.005, 0.05])) else 0
            return np.hstack([tcp, mug, ghost, tcp - mug, mug - ghost, vt, vt - vm, d.qpos[6], d.qvel[6], robust])
        if self.kind == "v0":
            return np.hstack([tcp, mug, ghost, self.grasp_count(), sx[self.pad_site]])
        return np.hstack([tcp, mug, ghost, self.grasp_count(), d.site_velocity(self.tcp)[3:]])

    def reward_v2(self, o, a):
        mug_z, g2m, m2t, gv, grasped, grip = o[5], o[9:12], o[12:15], o[15:18], o[23], a[-1]
        xy = np.linalg.norm(g2m[:2]); zerr = abs(g2m[2] - 0.02); place = np.linalg.norm(m2t)
        ready = np.exp(-10 * xy) * np.exp(-20 * zerr)
        r = 2.0 * ready + 2.0 * grip * ready + 10.0 * grasped * ready + 8.0 * grasped * np.tanh(8.0 * max(0, mug_z))
        r += grasped * (4.0 * np.exp(-15 * place) - 1.5 * place)
        if grasped and place < 0.05:
            r += 50.0
        r += -1.0 * max(0, -g2m[2]) + -0.01 * np.linalg.norm(gv)
        return r

    def reward_v0(self, o, a):
        gp, bc, tp, gs, pad = o[:3], o[3:6], o[6:9], o[9], o[10:13]
        hh = self.mug_size[-1]
        top, bot = bc[2] + hh, bc[2] - hh
        pad_top = pad[2] - top; g2c = gp[2] - bc[2]; herr = np.linalg.norm(gp[:2] - bc[:2])
        valid = gs == 2 and abs(pad_top) < 0.04 and herr < 0.03
        height_error = g2c - 0.5; z_tol = 0.1
        descent = 1 * ((1 / z_tol) * (height_error + z_tol) * np.exp(-(1 / z_tol) * height_error))
        ready = np.exp(-herr ** 2) * np.exp(-pad_top ** 2) * np.exp(-height_error ** 2) * 100 * np.exp(-a[-1] ** 2)
        align = 4 * np.exp(-60 * herr ** 2); grip = a[-1]
        grasp = 5.5 * (gs >= 1) + 8.5 * (gs == 2) + 23.5 * grip * ready + 28.5 * (gs == 2) * ready + 11.5 * (gs == 2) * ready * np.tanh(8 * grip)
        lift = 12 * (gs == 2) * np.tanh(4 * bot)
        dplace = np.linalg.norm(bc - tp)
        placement = -2 * dplace + 20 * np.exp(-70 * dplace ** 2)
        if dplace < 0.05 and valid:
            placement += 40
        danger = min(0, -100000000000 * (bc[2] - gp[2] + 0.5) ** 3)
        toppled = bc[2] <= max(self.mug_size[0], self.mug_size[1])
        pen = -40 * self.self_collision() + -25 * self.table_collision() + -8 * toppled + -4 * max(0, pad_top) + danger   # ur3e_env.py:339-345
        return descent + align + grasp + lift + placement + 700.5 * grip * ready + 1700.5 * (gs == 2) * ready * np.tanh(10 * grip) + pen

    def step(self, action):
        d = self.d
        action = np.asarray(action, dtype=float)
        if self.kind == "direct":
            u = action
        else:
            traj = np.hstack([action[:3], TOOL_ROTVEC, action[-1]])
            u = d.pid_task_ctrl(self.tcp, traj, self.gains)
        d.ctrl[:] = u
        d.step(self.frame_skip)
        o = self.obs()
        topple_z = max(self.mug_size[0], self.mug_size[1])
        if self.kind == "v2":
            r = self.reward_v2(o, action)
            self.t += 1
            dpick = np.linalg.norm(o[:3] - o[3:6])
            term = bool(dpick > 1 or self.self_collision() or o[5] <= topple_z)
            trunc = self.t >= self.max_steps
            if np.linalg.norm(o[3:6] - o[6:9]) < 0.05:
                term = True; r += 50.0
        elif self.kind == "v0":
            r = self.reward_v0(o, action)
            dpick = np.linalg.norm(o[:3] - o[3:6]); dplace = np.linalg.norm(o[3:6] - o[6:9])
            term = bool(dplace < 0.005 or dpick > 1 or self.self_collision() or o[5] <= topple_z)
            trunc = self.t >= self.max_steps
            self.t += 1
        else:
            r = -1.0; term = False; trunc = self.t >= self.max_steps; self.t += 1
        return o, r, term, trunc


class OracleCtrlLoop:
    """Controller demo loops: one controller evaluation per mj_step (move_j.py:76-86, move_l_task.py:55-69)."""

    def __init__(self, xml, mode, gains):
        self.m = O.Model(xml); self.d = O.Data(self.m); self.mode = mode; self.gains = np.asarray(gains, dtype=float)
        self.tcp = self.m.id("site", "tcp") if "tcp" in self.m.names["site"] else -1

    def set_state(self, qpos, qvel):
        self.d.reset(); self.d.set_state(qpos, qvel); self.d.forward()

    def step(self, target):
        d = self.d
        if self.mode == "pd_joint":
            u = d.pd_joint_ctrl(target[:6], self.gains[:6], self.gains[6:12])
            if self.m.nu > 6:
                u = np.hstack([u, target[6] * self.m.py["actuator_ctrlrange"][-1][1]])
        elif self.mode == "pid_task":
            u = d.pid_task_ctrl(self.tcp, target, self.gains)
        elif self.mode == "pinv":
            u = self.pinv_ctrl(target)
        else:
            u = np.asarray(target, dtype=float)
        d.ctrl[:] = u
        d.step(1)
        return d.qpos.copy(), d.qvel.copy()


def _pinv_ctrl(self, traj):
    """controller/move_l.py:15-78 restated with numpy (np.linalg.pinv, as the reference)."""
    d, m = self.d, self.m
    jp, jr = d.jac_site(self.tcp)
    tcp_pos = d.site_xpos.reshape(-1, 3)[self.tcp]; tcp_mat = d.site_xmat.reshape(-1, 9)[self.tcp]
    errs = [traj[:3] - tcp_pos, O.rot_err(tcp_mat, traj[3:6])]
    jr_rng = m.py["jnt_range"][:6]; cr = m.py["actuator_ctrlrange"][:6]
    u = np.zeros(6); g = self.gains
    for blk, (J, e) in enumerate(zip((jp[:, :6], jr[:, :6]), errs)):
        dth = np.linalg.pinv(J) @ e
        q = d.qpos[:6]
        tq = np.clip(q + dth, jr_rng[:, 0], jr_rng[:, 1])
        ub = g[12 * blk:12 * blk + 6] * (tq - q) + g[12 * blk + 6:12 * blk + 12] * -d.qvel[:6]
        u += np.clip(ub, cr[:, 0], cr[:, 1])
    return np.hstack([u, traj[6] * m.py["actuator_ctrlrange"][-1][1]])


OracleCtrlLoop.pinv_ctrl = _pinv_ctrl
