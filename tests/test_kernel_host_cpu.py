"""CPU suite part 2: the kernel's own source (engine.cuh), compiled as plain C++ by tests/hostcheck, against the oracle.
This checks the CUDA path's mathematics without a GPU (the GPU parity tests re-check the compiled kernel through the C ABI);
it also checks the C++ MJCF loader against the oracle's independent Python loader."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.hostcheck import build as HC

XMLS = ["ur3e_raw.xml", "ur3e_2f85.xml", "main.xml"]


@pytest.mark.parametrize("xml", XMLS)
def test_cpp_loader_matches_python_loader(assets, xml):
    path = assets + "/" + xml
    m = O.Model(path)
    for f in O.DBL_FIELDS:
        a, b = np.asarray(m.arr(f)), HC.model_array(path, f)
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-12, atol=1e-14), f
    for f in O.INT_FIELDS:
        a, b = np.asarray(m.py[f]).ravel(), HC.model_array(path, f)
        assert a.shape == b.shape and np.array_equal(a, b), f


def _state(m, rng, xml):
    if m.nkey:
        qp, _ = m.key("down")
    else:
        qp = np.zeros(m.nq)
    qp = qp.copy(); qp[:6] += rng.uniform(-0.3, 0.3, 6)
    if m.nq >= 14:
        qp[6] = qp[10] = rng.uniform(0, 0.5)
    if xml == "main.xml":
        qp[16] -= rng.uniform(0, 0.001)
    return qp, rng.uniform(-1, 1, m.nv), rng.uniform(-5, 5, m.nu)


@pytest.mark.parametrize("xml", XMLS)
def test_forward_matches_oracle_f64(assets, xml):
    path = assets + "/" + xml
    m = O.Model(path); d = O.Data(m); rng = np.random.default_rng(5)
    for _ in range(4):
        qp, qv, ctrl = _state(m, rng, xml)
        d.reset(); d.set_state(qp, qv); d.ctrl[:] = ctrl; d.forward()
        r = HC.run(path, qp, qv, ctrl, np.zeros(m.nv), 0)
        assert (r["ncon"], r["nefc"]) == (d.ncon, d.nefc)
        assert np.abs(r["M"] - d.fullM()).max() < 1e-13
        assert np.abs(r["bias"] - d.qfrc_bias).max() < 1e-12
        assert np.abs(r["qacc"] - d.qacc).max() < 1e-9 * max(1.0, np.abs(d.qacc).max())
        assert np.abs(r["fc"] - d.qfrc_constraint).max() < 1e-9 * max(1.0, np.abs(d.qfrc_constraint).max())


@pytest.mark.parametrize("xml", XMLS)
def test_trajectory_matches_oracle_f64(assets, xml):
    path = assets + "/" + xml
    m = O.Model(path); d = O.Data(m); rng = np.random.default_rng(9)
    qp, qv, ctrl = _state(m, rng, xml)
    d.reset(); d.set_state(qp, qv); d.ctrl[:] = ctrl; d.step(300)
    r = HC.run(path, qp, qv, ctrl, np.zeros(m.nv), 300)
    assert r["warn"] == 0 and r["overflow"] == 0
    assert np.abs(r["qpos"] - d.qpos).max() < 1e-10 and np.abs(r["qvel"] - d.qvel).max() < 1e-8


def test_trajectory_f32_drift_bound(assets):
    """float32 arithmetic of the same source: drift vs float64 oracle after 300 mj_steps of main.xml stays below 1e-3 (qpos)."""
    path = assets + "/main.xml"
    m = O.Model(path); d = O.Data(m); rng = np.random.default_rng(9)
    qp, qv, ctrl = _state(m, rng, "main.xml")
    d.reset(); d.set_state(qp, qv); d.ctrl[:] = ctrl; d.step(300)
    r = HC.run(path, qp, qv, ctrl, np.zeros(m.nv), 300, use_float=1, max_iter=8, tol=1e-7)
    assert np.isfinite(r["qpos"]).all() and np.abs(r["qpos"] - d.qpos).max() < 1e-3


def test_gripper_closing_contacts_pads(assets):
    """closing the 2F85 on nothing brings the pad boxes into contact (pad-pad box-box pairs)."""
    path = assets + "/ur3e_2f85.xml"
    m = O.Model(path); d = O.Data(m)
    qp, qv = m.key("down"); ctrl = np.zeros(7); ctrl[6] = 255
    d.reset(); d.set_state(qp, qv); d.ctrl[:] = ctrl; d.step(600)
    r = HC.run(path, qp, qv, ctrl, np.zeros(m.nv), 600)
    assert d.ncon > 0 and r["ncon"] == d.ncon
    assert np.abs(r["qpos"] - d.qpos).max() < 1e-5     # 600 steps without re-seeding through a 2-point edge contact


def test_logging_record_matches_oracle_f64(assets):
    """The step kernel's cold sensor path compiled for the CPU: actuator forces, controller output and the six <torque> site sensors
    (mj_rnePostConstraint restated) against the oracle, contact-free and with both pads and the gripper's hull pressing on the mug."""
    path = assets + "/main.xml"
    m = O.Model(path); d = O.Data(m)
    qp, qv = m.key("down"); d.reset(); d.set_state(qp, qv)
    u = np.zeros(7); u[6] = 255.0
    seen_contact = False
    for k in range(450):
        d.ctrl[:] = u; d.step(1)
        if k % 90 == 89:
            q, v = d.qpos.copy(), d.qvel.copy()
            o = O.Data(m); o.reset(); o.set_state(q, v); o.ctrl[:] = u; o.forward()
            got = HC.sensors(path, q, v, u)
            assert np.abs(got[:7] - o.sensors()[:7]).max() < 1e-9 and np.array_equal(got[21:28], u)
            assert np.abs(got[28:46] - o.torque_sensors().ravel()).max() < 1e-9
            seen_contact |= o.ncon > 4
    assert seen_contact
