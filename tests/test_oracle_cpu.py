"""CPU suite part 1: the oracle against every known answer available for this path.

(i) the one golden vector the reference records (tcp@'down', assets/main.xml:415) and the model sizes it comments;
(ii) fixtures produced by running the reference's own Python verbatim over the oracle (tools/make_golden.py);
(iii) analytic invariants of the restated dynamics.  Parity against a real MuJoCo build stays UNPINNED.
"""
import os

import numpy as np
import pytest

from oracle import envs as OE
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b, floor=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor)))


def test_model_sizes(assets):
    # controller/move_l_mug.py:29-32: ur3e 6, 2f85 8, mug 7/6 -> 21 nq / 20 nv / 7 nu
    dims = {"ur3e_raw.xml": (6, 6, 6, 8), "ur3e_2f85.xml": (14, 14, 7, 23), "main.xml": (21, 20, 7, 25)}
    for xml, (nq, nv, nu, nb) in dims.items():
        m = O.Model(assets + "/" + xml)
        assert (m.nq, m.nv, m.nu, m.nbody) == (nq, nv, nu, nb)


def test_tcp_golden_vector(assets):
    m = O.Model(assets + "/main.xml"); d = O.Data(m)
    qp, qv = m.key("down"); d.set_state(qp, qv); d.forward()
    tcp = d.site_xpos.reshape(-1, 3)[m.id("site", "tcp")]
    assert np.allclose(tcp, [0.29799994, 0.13349916, 0.1682003], atol=5e-9)            # assets/main.xml:415
    assert np.allclose(d.site_xpos.reshape(-1, 3)[m.id("site", "handle_site")], [0.29799994, 0.13349916, 0.055111], atol=1e-12)
    assert np.allclose(d.xpos.reshape(-1, 3)[m.id("body", "ghost")], [0.29799994, 0.25, 0.055111], atol=1e-12)
    # the hard-coded tool orientation of ur3e_env2.py:74 is the tcp rotvec at 'down': zero rotation error there
    assert np.abs(O.rot_err(d.site_xmat.reshape(-1, 9)[m.id("site", "tcp")], [-1.2092, -1.2092, 1.2092])).max() < 1e-4


def test_rot_err_matches_scipy():
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(0)
    for _ in range(200):
        Rs = R.from_rotvec(rng.uniform(-3, 3, 3)); rv = rng.uniform(-3, 3, 3)
        want = (R.from_rotvec(rv) * R.from_matrix(Rs.as_matrix()).inv()).as_rotvec()    # controller_func.py:38-45
        assert np.allclose(O.rot_err(Rs.as_matrix().ravel(), rv), want, atol=1e-10)


def test_controllers_against_reference_fixture(assets):
    g = np.load(GOLD + "/controllers.npz")
    m = O.Model(assets + "/ur3e_2f85.xml"); d = O.Data(m); tcp = m.id("site", "tcp")
    for i in range(len(g["qpos"])):
        d.reset(); d.set_state(g["qpos"][i], g["qvel"][i]); d.forward()
        assert rel(d.pid_task_ctrl(tcp, g["traj"][i], g["gains_task"]), g["u_task"][i]) < 1e-10
        assert np.allclose(O.rot_err(d.site_xmat.reshape(-1, 9)[tcp], g["traj"][i][3:6]), g["rot_err"][i], atol=1e-10)
        assert rel(d.pd_joint_ctrl(g["target_j"][i][:6], g["gains_j"][:6], g["gains_j"][6:]), g["u_joint"][i][:6]) < 1e-12
        assert abs(g["u_joint"][i][6] - g["target_j"][i][6] * 255.0) < 1e-12                      # grip_ctrl
    loop = OE.OracleCtrlLoop(assets + "/ur3e_2f85.xml", "pinv", g["gains_pinv"])                  # move_l.ctrl (pinv IK)
    for i in range(len(g["qpos"])):
        loop.set_state(g["qpos"][i], g["qvel"][i])
        assert rel(loop.pinv_ctrl(g["traj"][i]), g["u_pinv"][i]) < 1e-9


@pytest.mark.parametrize("kind", ["v2", "v0", "indirect"])
def test_env_restatement_against_reference_fixture(assets, kind):
    """oracle/envs.py (the restatement the GPU tests use) == the reference's env classes run verbatim, same physics."""
    g = np.load(GOLD + "/env_%s.npz" % kind)
    env = OE.OracleEnv(assets + "/main.xml", kind)
    env.set_state(g["qpos0"], g["qvel0"])
    assert rel(env.obs(), g["obs0"]) < 1e-12
    for k in range(len(g["reward"])):
        o, r, te, tr = env.step(g["actions"][k])
        assert rel(o, g["obs"][k], 1e-3) < 1e-7, k      # scipy-vs-C rounding differences are amplified by the contact-rich episode
        assert abs(r - g["reward"][k]) <= 1e-6 * max(1.0, abs(g["reward"][k])), k
        assert te == bool(g["terminated"][k]) and tr == bool(g["truncated"][k])
        assert rel(env.d.qpos, g["qpos"][k], 1e-3) < 1e-7


def test_mass_matrix_and_jacobian_invariants(assets):
    m = O.Model(assets + "/main.xml"); d = O.Data(m)
    rng = np.random.default_rng(1)
    qp, _ = m.key("down"); qp[:14] += rng.uniform(-0.3, 0.3, 14); qv = rng.uniform(-1, 1, m.nv)
    d.set_state(qp, qv); d.forward()
    M = d.fullM()
    assert np.allclose(M, M.T, atol=1e-14) and np.linalg.eigvalsh(M).min() > 0
    tcp = m.id("site", "tcp"); jp, jr = d.jac_site(tcp)
    assert np.allclose(np.r_[jr @ d.qvel, jp @ d.qvel], d.site_velocity(tcp), atol=1e-13)     # J qdot == mj_objectVelocity
    p0 = d.site_xpos.reshape(-1, 3)[tcp].copy(); num = np.zeros((3, 14)); eps = 1e-6
    for i in range(14):
        q = qp.copy(); q[i] += eps; d.set_state(q, qv); d.forward(); num[:, i] = (d.site_xpos.reshape(-1, 3)[tcp] - p0) / eps
    assert np.abs(num - jp[:, :14]).max() < 1e-5


def test_gravity_bias_is_potential_gradient(assets):
    m = O.Model(assets + "/ur3e_2f85.xml"); d = O.Data(m)
    qp, _ = m.key("down"); zero = np.zeros(m.nv)

    def pe(q):
        d.set_state(q, zero); d.forward()
        return sum(m.py["body_mass"][b] * 9.81 * d.xipos.reshape(-1, 3)[b, 2] for b in range(m.nbody))
    eps = 1e-6
    g = np.array([(pe(qp + eps * np.eye(14)[i]) - pe(qp - eps * np.eye(14)[i])) / (2 * eps) for i in range(14)])
    d.set_state(qp, zero); d.forward()
    assert np.abs(g - d.qfrc_bias).max() < 1e-7


def test_energy_conservation_raw_arm(assets):
    """ur3e_raw.xml has no damping / friction / constraints: semi-implicit Euler at dt = 1e-4 conserves energy to O(dt)."""
    m = O.Model(assets + "/ur3e_raw.xml"); d = O.Data(m)
    rng = np.random.default_rng(0)
    d.set_state(rng.uniform(-0.5, 0.5, 6), rng.uniform(-0.5, 0.5, 6))

    def energy():
        d.forward(); M = d.fullM()
        return 0.5 * d.qvel @ M @ d.qvel + sum(m.py["body_mass"][b] * 9.81 * d.xipos.reshape(-1, 3)[b, 2] for b in range(m.nbody))
    e0 = energy(); d.step(3000)
    assert d.nefc == 0 and abs(energy() - e0) < 0.02 * abs(e0) + 0.05


def test_constraint_solution_is_a_minimum(assets):
    """At the solver's qacc the primal gradient vanishes: M a - qfrc_smooth - J^T f = 0 with f from the row states."""
    m = O.Model(assets + "/main.xml"); d = O.Data(m)
    qp, qv = m.key("down"); qp[16] -= 0.001; qp[6] = qp[10] = 0.3
    d.set_state(qp, np.random.default_rng(2).uniform(-0.3, 0.3, m.nv)); d.forward()
    nv, nefc = m.nv, d.nefc
    assert d.ncon == 4 and nefc >= 25
    J = d.efc_J[:nefc * nv].reshape(nefc, nv)
    grad = d.fullM() @ d.qacc - d.qfrc_smooth - J.T @ d.efc_force[:nefc]
    assert np.abs(grad).max() < 1e-8 * max(1.0, np.abs(d.qfrc_smooth).max())
    assert np.allclose(J.T @ d.efc_force[:nefc], d.qfrc_constraint, atol=1e-10)


def test_mug_rests_on_table(assets):
    m = O.Model(assets + "/main.xml"); d = O.Data(m)
    qp, qv = m.key("down"); d.set_state(qp, qv); d.forward()
    for _ in range(150):     # the arm is held up by gravity compensation (without it the gripper's hull lands on the mug)
        d.ctrl[:6] = d.qfrc_bias[:6]; d.step(10)
    assert d.ncon >= 4 and abs(d.qpos[16] - 0.055111) < 2e-4 and np.abs(d.qvel[14:]).max() < 1e-3
    assert d.warn_bad == 0


def test_torque_sensors_equal_holding_torque_at_rest(assets):
    """The <torque> site sensors (reference assets/main.xml:384-391) against a statics identity: with the arm held still by gravity
    compensation, the component of each joint's interaction torque along its own axis is the torque that holds the link, qfrc_bias[j]
    (the sites sit at the link origins, on the joint axes: main.xml:113,119,125,131,137,143)."""
    m = O.Model(assets + "/main.xml"); d = O.Data(m)
    qp, qv = m.key("down"); qp[:6] += [0.2, 0.1, -0.2, 0.3, 0.1, -0.1]
    d.set_state(qp, qv); d.forward()
    for _ in range(1500):                      # let the gripper's springs settle while the arm is held
        d.ctrl[:6] = d.qfrc_bias[:6]; d.step(1)
    d.ctrl[:6] = d.qfrc_bias[:6]; d.forward()
    assert np.abs(d.qacc[:14]).max() < 1e-3 and np.abs(d.qvel[:14]).max() < 1e-3
    tq = d.torque_sensors()
    assert tq.shape == (6, 3)
    axis = [2, 1, 1, 1, 2, 1]                  # shoulder_pan z, shoulder_lift y, elbow y, wrist_1 y, wrist_2 z, wrist_3 y (main.xml:112-142)
    for j in range(6):
        assert abs(tq[j, axis[j]] - d.qfrc_bias[j]) < 2e-3 * max(1.0, abs(d.qfrc_bias[j])), (j, tq[j], d.qfrc_bias[j])
    # the wrench through the shoulder carries the whole arm: its vertical force component is the moving mass times g -- checked
    # through the torque it produces about the base is not available here, so check the lift joint against the closed form instead:
    # torque about the lift axis = sum over the links above of m g x (horizontal lever), i.e. qfrc_bias of that joint (already above)
