#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the REFERENCE's own Python verbatim (from /root/reference) on
top of the oracle through oracle/refshim.py.  Run in the build container only:

    python tools/make_golden.py

What the fixtures pin: the reference's controller / observation / reward / termination / truncation code
(controller_func.py, gym_utils.py, ur3e_env2.py, ur3e_env.py, imitation_env_*.py) executed as written,
including scipy's Rotation conventions and the stale-kinematics order of reads.  What they do NOT pin: the
physics, which is the oracle's restatement of mj_step (MuJoCo is not installable here).
"""
import io
import os
import sys
import contextlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import refshim  # noqa: E402

uu, cf, gu = refshim.import_reference()
import yaml  # noqa: E402


def env_episode(cls_name, module, steps, seed, direct=False):
    mod = __import__(module, fromlist=[cls_name])
    np.random.seed(seed)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        env = getattr(mod, cls_name)()
        if direct:
            # ImitationEnvDirect.reset_model raises UnboundLocalError in the reference (get_init(..., "stochastic") without a noise_mag,
            # imitation_env_direct.py:111 vs gym_utils.py:48-60), so the episode starts from the keyframe through MujocoEnv.set_state
            # through MujocoEnv.set_state -- from a state of the ur3e-v0 fixture in which both pads touch the mug, so that the
            # observation's grasp count and the tcp velocity are exercised by raw actuator commands
            v0 = np.load(os.path.join(OUT, "env_v0.npz"))
            k0 = int(np.argmax(v0["obs"][:, 9] == 2)) + 5
            env.set_state(v0["qpos"][k0].copy(), v0["qvel"][k0].copy()); env.t = 0
            obs0 = env._get_obs()
        else:
            obs0, _ = env.reset()
    qpos0, qvel0 = env.data.qpos.copy(), env.data.qvel.copy()
    rng = np.random.default_rng(seed)
    rec = dict(qpos0=qpos0, qvel0=qvel0, obs0=np.asarray(obs0, dtype=np.float64), actions=[], obs=[], reward=[], terminated=[], truncated=[], qpos=[], qvel=[],
               action_low=np.asarray(env.action_space.low, dtype=np.float64), action_high=np.asarray(env.action_space.high, dtype=np.float64),
               frame_skip=np.int64(env.frame_skip))
    mug = obs0[3:6].copy()
    for k in range(steps):
        if direct:
            # raw actuator commands: gravity-compensating torques (the controller's stale bias term) plus noise, gripper closed for the first half
            a = np.zeros(7); a[:6] = env.data.qfrc_bias[:6] + rng.uniform(-1.5, 1.5, 6); a[6] = 255.0 if k < steps // 2 else 0.0
        else:
            # scripted approach, close, lift: exercises pad-mug contacts, grasp flags and the reward branches
            z = mug[2] + 0.02 + max(0.0, 0.1 - 0.002 * k) + (0.0 if k < 170 else 0.0005 * (k - 170))
            a = np.hstack([mug[:2] + rng.normal(0, 0.001, 2), z, 1.0 if k > 90 else 0.0])
        with contextlib.redirect_stdout(sink):
            o, r, te, tr, _ = env.step(a)
        rec["actions"].append(a); rec["obs"].append(np.asarray(o, dtype=np.float64)); rec["reward"].append(float(r))
        rec["terminated"].append(bool(te)); rec["truncated"].append(bool(tr)); rec["qpos"].append(env.data.qpos.copy()); rec["qvel"].append(env.data.qvel.copy())
        if te or tr:
            break
    return {k: np.asarray(v) for k, v in rec.items()}


def pick_place_episode():
    """The reference's scripted expert (controller/move_l_mug.py:36-41 targets through build_traj_l_pick_place, build_traj.py:28-59:
    approach the mug's centre closing to 0.5, lift 0.15 m closing fully, carry to the ghost + 0.025, release), fed to UR3eEnv2.step
    as [x, y, z, grip] with 15 waypoints per segment held 60 env-steps (= the script's hold of 120 mj_steps at frame_skip 2), from
    the deterministic keyframe reset the script uses.  Runs until the env reports success (mug within 0.05 of the ghost)."""
    mod = __import__("gymnasium_env.envs.ur3e_env2", fromlist=["UR3eEnv2"])
    sink = io.StringIO()
    np.random.seed(3)
    with contextlib.redirect_stdout(sink):
        env = mod.UR3eEnv2()
        env.reset()
        qp, qv = gu.get_init(env.model, "deterministic", "down")
        env.set_state(np.array(qp, dtype=np.float64), np.array(qv, dtype=np.float64)); env.t = 0
        o = env._get_obs()
    start = np.hstack([o[0:3], 0.0]); pick = np.hstack([o[3:6], 0.5]); up = pick + [0, 0, 0.15, 0.5]
    place = np.hstack([o[6:9], 1.0]) + [0, 0, 0.025, 0]; drop = place.copy(); drop[3] = 0.0
    way = [start, pick, up, place, drop]
    H, N, EVERY = 60, 15, 25
    rec = dict(qpos0=env.data.qpos.copy(), qvel0=env.data.qvel.copy(), obs0=np.asarray(o, dtype=np.float64), actions=[], obs=[], reward=[], terminated=[], truncated=[],
               seed_step=[], seed_qpos=[], seed_qvel=[], seed_ws=[], ncon=[])
    for k in range(4 * H * N):
        if k % EVERY == 0:
            rec["seed_step"].append(k); rec["seed_qpos"].append(env.data.qpos.copy()); rec["seed_qvel"].append(env.data.qvel.copy())
            rec["seed_ws"].append(np.array(env.data.qacc_warmstart, dtype=np.float64).copy())
        seg, r = divmod(k, H * N)
        a = way[seg] + (way[seg + 1] - way[seg]) * ((r // H + 1) / N)
        with contextlib.redirect_stdout(sink):
            o, rew, te, tr, _ = env.step(a)
        rec["actions"].append(a); rec["obs"].append(np.asarray(o, dtype=np.float64)); rec["reward"].append(float(rew))
        rec["terminated"].append(bool(te)); rec["truncated"].append(bool(tr)); rec["ncon"].append(int(env.data.ncon))
        if te or tr:
            break
    return {k: np.asarray(v) for k, v in rec.items()}


def collision_vectors():
    """get_self_collision / get_table_collision (utils/gym_utils.py:146-201) as the reference's own env classes use them: one step of
    UR3eEnv2 and UR3eEnv from injected arm poses -- keyframe 'down', elbow folded onto the shoulder (self-collision: both classes
    terminate, ur3e_env2.py:244-246 / ur3e_env.py:441-443, v0's reward carries -40), tool pressed into the table (table collision: v0's
    -25 penalty, no termination)."""
    out = {}
    sink = io.StringIO()
    for name, cls_name, module in (("v2", "UR3eEnv2", "gymnasium_env.envs.ur3e_env2"), ("v0", "UR3eEnv", "gymnasium_env.envs.ur3e_env")):
        mod = __import__(module, fromlist=[cls_name])
        np.random.seed(1)
        with contextlib.redirect_stdout(sink):
            env = getattr(mod, cls_name)()
            env.reset()
        qp0, qv0 = gu.get_init(env.model, "deterministic", "down")
        poses = []
        for dq in ([0, 0, 0, 0, 0, 0], [0, 0, 1.2292, 0, 0, 0], [0, 0, 1.3292, 0, 0, 0], [0.5, 0.6, 0.0, -0.6, 0, 0], [0.5, 0.75, 0.1, -0.85, 0, 0]):
            q = np.array(qp0, dtype=np.float64); q[:6] += dq; poses.append(q)
        rec = dict(qpos=[], action=[], obs=[], reward=[], terminated=[], self_collision=[], table_collision=[], ncon=[])
        for q in poses:
            with contextlib.redirect_stdout(sink):
                env.set_state(q, np.array(qv0, dtype=np.float64)); env.t = 0
                o0 = env._get_obs()
                a = np.hstack([o0[:3], 0.0])
                rec["self_collision"].append(int(gu.get_self_collision(env.model, env.data, env.collision_cache)))
                rec["table_collision"].append(int(gu.get_table_collision(env.model, env.data, env.collision_cache)))
                rec["ncon"].append(int(env.data.ncon))
                o, r, te, tr, _ = env.step(a)
            rec["qpos"].append(q); rec["action"].append(a); rec["obs"].append(np.asarray(o, dtype=np.float64)); rec["reward"].append(float(r)); rec["terminated"].append(bool(te))
        for k, v in rec.items():
            out[name + "_" + k] = np.asarray(v)
        print("collision", name, "self", rec["self_collision"], "table", rec["table_collision"], "terminated", rec["terminated"], "ncon", rec["ncon"])
    out["qvel"] = np.array(qv0, dtype=np.float64)
    return out


def truncation_steps():
    """Step index (1-based) at which each reference env class first reports truncated=True under a hold-still action:
    ur3e_env2.py:89-92 increments t before the test (2500), the others test first (ur3e_env.py:183-194 -> 501,
    imitation_env_indirect.py:97-101 -> 2501, imitation_env_direct.py:99-103 -> 1201).  SURVEY 8a row a15."""
    out = {}
    sink = io.StringIO()
    for name, cls_name, module, direct in (("v2", "UR3eEnv2", "gymnasium_env.envs.ur3e_env2", False), ("v0", "UR3eEnv", "gymnasium_env.envs.ur3e_env", False),
                                           ("indirect", "ImitationEnvIndirect", "gymnasium_env.envs.imitation_env_indirect", False),
                                           ("direct", "ImitationEnvDirect", "gymnasium_env.envs.imitation_env_direct", True)):
        mod = __import__(module, fromlist=[cls_name])
        np.random.seed(0)
        with contextlib.redirect_stdout(sink):
            env = getattr(mod, cls_name)()
            if not direct:    # ImitationEnvDirect.reset_model raises (see env_episode); the others need it for their error buffers
                env.reset()
            qp, qv = gu.get_init(env.model, "deterministic", "down")
            env.set_state(np.array(qp, dtype=np.float64), np.array(qv, dtype=np.float64)); env.t = 0
            obs = env._get_obs()
        hold = np.hstack([obs[:3], 0.0])
        first, term_any = -1, False
        for k in range(1, 2600):
            a = np.hstack([env.data.qfrc_bias[:6], 0.0]) if direct else hold
            with contextlib.redirect_stdout(sink):
                o, r, te, tr, _ = env.step(a)
            term_any |= bool(te)
            if tr:
                first = k; break
        out[name] = np.array([first, int(term_any)], dtype=np.int64)
        print("truncation", name, "first truncated step", first, "terminated before:", term_any)
    return out


def controller_vectors(seed):
    """pid_task_ctrl / pd_joint_ctrl / get_rot_err of the reference on random ur3e_2f85.xml states."""
    import mujoco
    m, d = uu.load_model("assets/ur3e_2f85.xml")
    rng = np.random.default_rng(seed)
    with open("controller/config/config_l_task.yml") as f:
        yml = yaml.safe_load(f)
    pos_g = {k: np.diag(v) for k, v in yml["pos"].items()}; rot_g = {k: np.diag(v) for k, v in yml["rot"].items()}
    with open("controller/config/config_j.yml") as f:
        yj = yaml.safe_load(f)
    jg = {k: np.diag(v) for k, v in yj["qpos"].items()}     # move_j.py:50
    from controller.move_j import ctrl as move_j_ctrl      # move_j.py:14-27 (pd_joint_ctrl + grip_ctrl)
    from controller.move_l import ctrl as move_l_ctrl      # move_l.py:15-31 (pinv IK + two pd_joint_ctrl)
    with open("controller/config/config_l.yml") as f:
        yl = yaml.safe_load(f)
    lp = {k: np.diag(v) for k, v in yl["pos"].items()}; lr = {k: np.diag(v) for k, v in yl["rot"].items()}
    out = dict(qpos=[], qvel=[], traj=[], u_task=[], rot_err=[], target_j=[], u_joint=[], u_pinv=[])
    for _ in range(24):
        uu.reset(m, d, "down")
        d.qpos[:6] += rng.uniform(-0.5, 0.5, 6); d.qvel[:] = rng.uniform(-1, 1, m.nv)
        mujoco.mj_forward(m, d)
        tcp = uu.get_site_xpos(m, d, "tcp")
        traj = np.hstack([tcp + rng.uniform(-0.1, 0.1, 3), rng.uniform(-2.5, 2.5, 3), rng.uniform(0, 1)])
        u = cf.pid_task_ctrl(0, m, d, traj, pos_g, rot_g, np.zeros((1, 3)), np.zeros((1, 3)), np.zeros(3), np.zeros(3))
        e = cf.get_rot_err(0, m, d, traj[3:6], np.zeros((1, 3)))
        tj = np.hstack([d.qpos[:6] + rng.uniform(-0.2, 0.2, 6), 0.3])
        uj = move_j_ctrl(0, m, d, tj, jg, np.zeros((1, 6)))
        out["u_pinv"].append(move_l_ctrl(0, m, d, traj, lp, lr, np.zeros((1, 3)), np.zeros((1, 3))))
        out["qpos"].append(d.qpos.copy()); out["qvel"].append(d.qvel.copy()); out["traj"].append(traj); out["u_task"].append(u)
        out["rot_err"].append(e); out["target_j"].append(tj); out["u_joint"].append(uj)
    g = dict(gains_task=np.hstack([np.diag(pos_g["kp"]), np.diag(pos_g["kd"]), np.diag(rot_g["kp"]), np.diag(rot_g["kd"])]))
    g["gains_pinv"] = np.hstack([np.diag(lp["kp"]), np.diag(lp["kd"]), np.diag(lr["kp"]), np.diag(lr["kd"])])
    if jg:
        g["gains_j"] = np.hstack([np.diag(jg["kp"]), np.diag(jg["kd"])])
    return {**{k: np.asarray(v) for k, v in out.items()}, **g}


def main():
    os.makedirs(OUT, exist_ok=True)
    jobs = [("env_v2", "UR3eEnv2", "gymnasium_env.envs.ur3e_env2", 260, 11, False),
            ("env_v0", "UR3eEnv", "gymnasium_env.envs.ur3e_env", 260, 12, False),
            ("env_indirect", "ImitationEnvIndirect", "gymnasium_env.envs.imitation_env_indirect", 260, 13, False),
            ("env_direct", "ImitationEnvDirect", "gymnasium_env.envs.imitation_env_direct", 260, 14, True)]
    only = sys.argv[1:]
    for name, cls, module, steps, seed, direct in jobs:
        if only and name not in only:
            continue
        rec = env_episode(cls, module, steps, seed, direct)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "steps", len(rec["reward"]), "sum reward %.6f" % rec["reward"].sum(), "grasp max", rec["obs"][:, 23 if rec["obs"].shape[1] == 24 else 9].max(),
              "term", rec["terminated"].any(), "trunc", rec["truncated"].any())
    if not only or "env_v2_pick" in only:
        rec = pick_place_episode()
        np.savez_compressed(os.path.join(OUT, "env_v2_pick.npz"), **rec)
        print("env_v2_pick steps", len(rec["reward"]), "sum reward %.4f" % rec["reward"].sum(), "robust-grasp steps", int(rec["obs"][:, 23].sum()), "max ncon", rec["ncon"].max(),
              "max mug z %.4f" % rec["obs"][:, 5].max(), "terminated", bool(rec["terminated"][-1]))
    if not only or "collision" in only:
        np.savez_compressed(os.path.join(OUT, "collision.npz"), **collision_vectors())
    if not only or "truncation" in only:
        np.savez_compressed(os.path.join(OUT, "truncation.npz"), **truncation_steps())
    if only and "controllers" not in only:
        return
    cv = controller_vectors(5)
    np.savez_compressed(os.path.join(OUT, "controllers.npz"), **cv)
    print("controllers", cv["u_task"].shape, cv["u_joint"].shape)
    # the one golden vector the reference itself records: tcp site position at keyframe 'down' (assets/main.xml:415)
    import mujoco
    m, d = uu.load_model("assets/main.xml"); uu.reset(m, d, "down")
    print("tcp@down", uu.get_site_xpos(m, d, "tcp"), "rotvec", uu.get_site_xrotvec(m, d, "tcp"))


if __name__ == "__main__":
    main()
