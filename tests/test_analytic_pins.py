"""Second-source pins for the physics restatement: closed-form answers derived from MuJoCo's published
constraint model (documentation chapter "Computation": impedance d(r), reference acceleration, R = (1 - d) / d * A_hat, elliptic
cones with impratio) and from rigid-body mechanics -- NOT from oracle/ur3e_oracle.c or SURVEY App. B pseudo-code.  MuJoCo 3.3.3
itself cannot be installed in this image or on the GPU box (profiles/r2_mujoco_probe.log), so these closed forms are what pins
the oracle, the kernel's source compiled for the CPU (tests/hostcheck) and, in the GPU-marked tests, the float64 CUDA build.

Model: tests/models/box_on_plane.xml (a physical pendulum + a 0.1 kg box on a plane, MuJoCo-default contact parameters).

  free fall        semi-implicit Euler: z_n = z_0 - g h^2 n (n + 1) / 2 exactly
  pendulum         small-angle period 2 pi sqrt(I / (m g l)), I = I_yy + m l^2, and energy conservation
  mj_setConst      dof_invweight0 = 1 / I (hinge), 1 / m and mean(1 / I_k) (free joint); body_invweight0 = tr(J M^-1 J^T) / 3
  impedance        d(r) of solimp (0.9, 0.95, 0.001, 0.5, 2): 0.9, 0.90625, 0.925, 0.94375, 0.95 at r / width = 0, 1/4, 1/2, 3/4, >= 1;
                   K = 1 / (dmax^2 tc^2 dr^2), B = 2 / (dmax tc) with tc = max(solref[0], 2 h)
  box at rest      four corner contacts share m g: each normal force m g / 4, penetration r solving r = (1 - d) g / (4 K d^2)
                   (f = D aref with D = d / ((1 - d) A_hat), A_hat = 1 / m, aref = K d r)
  sticking creep   on an incline below the friction angle the soft friction rows (K = 0, R_t = R_n / impratio) let the box creep at
                   v = g sin(theta) / (impratio B sum_i d_i / (1 - d_i))
  sliding          above the friction angle the box accelerates at about g (sin(theta) - mu cos(theta))   [coarse: 10 %]

Model: tests/models/double_pendulum.xml: mass matrix (CRBA) and bias forces (RNE) against the textbook closed form.

Model: tests/models/hinge_pins.xml (three independent unit hinges, no gravity): the joint-space rows and the affine actuator.

  joint limit      a motor torque tau pushes the hinge into its stop: penetration r solving r = tau (1 - d) A_hat / (K d^2), A_hat = 1 / I
  friction loss    tau below frictionloss: creep v = tau (1 - d0) A_hat / (d0 B) (pos = 0, so d = solimp[0]); above: acceleration (tau - F) / I
  affine actuator  the 2F85's `general` form: equilibrium gain ctrl = -b1 q, and q = -gain ctrl / b1 again once the force range saturates
                   transiently (damped joint)
"""
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = 9.81
M_BOX, I_HINGE, M_PEND, L_PEND = 0.1, 0.03 + 1.5 * 0.4 ** 2, 1.5, 0.4
DMIN, DMAX, WIDTH, TC = 0.9, 0.95, 0.001, 0.02
K_REF, B_REF = 1.0 / (DMAX ** 2 * TC ** 2), 2.0 / (DMAX * TC)
IMPRATIO = 10.0


def model_path(tmp_path, theta=0.0, halfz=0.01, z0=None):
    src = open(os.path.join(HERE, "models", "box_on_plane.xml")).read()
    z0 = halfz if z0 is None else z0
    src = src.replace("GRAVITY", "%.17g 0 %.17g" % (G * np.sin(theta), -G * np.cos(theta))).replace("HALFZ", repr(halfz)).replace("Z0", repr(z0))
    p = str(tmp_path / ("pins_%.4f_%.4f_%.4f.xml" % (theta, halfz, z0)))
    open(p, "w").write(src)
    return p


def impedance(r):
    x = min(abs(r) / WIDTH, 1.0)
    y = x * x / 0.5 if x <= 0.5 else 1.0 - (1.0 - x) ** 2 / 0.5          # power 2, midpoint 0.5
    return DMIN + (DMAX - DMIN) * y


def rest_penetration(gn):
    r = 1e-4
    for _ in range(200):
        d = impedance(r); r = (1.0 - d) * gn / (4.0 * K_REF * d * d)
    return r


class OracleSim:
    name = "oracle"

    def __init__(self, path):
        self.m = O.Model(path); self.d = O.Data(self.m); self.d.reset()

    def set_state(self, qpos, qvel):
        self.d.reset(); self.d.set_state(qpos, qvel)

    def step(self, n):
        self.d.step(n); return self.d.qpos.copy(), self.d.qvel.copy()


class KernelSourceSim:
    """engine.cuh compiled as plain C++ (tests/hostcheck): the CUDA path's arithmetic without a GPU."""
    name = "kernel-source"

    def __init__(self, path):
        from tests.hostcheck import build as HC
        self.HC, self.path, self.m = HC, path, O.Model(path)
        self.qpos, self.qvel, self.ws = np.array(self.m.arr("qpos0")).copy(), np.zeros(self.m.nv), np.zeros(self.m.nv)

    def set_state(self, qpos, qvel):
        self.qpos, self.qvel, self.ws = np.array(qpos, dtype=float), np.array(qvel, dtype=float), np.zeros(self.m.nv)

    def step(self, n):
        r = self.HC.run(self.path, self.qpos, self.qvel, np.zeros(self.m.nu), self.ws, n)
        assert r["warn"] == 0 and r["overflow"] == 0
        self.qpos, self.qvel, self.ws = r["qpos"].copy(), r["qvel"].copy(), r["qacc"].copy()
        return self.qpos.copy(), self.qvel.copy()


class GpuSim:
    """float64 CUDA build through the C ABI."""
    name = "cuda-f64"

    def __init__(self, path):
        import torch
        import ur3e_b200._lib as lib
        from ur3e_b200.batch import SimBatch, env_config
        from ur3e_b200.model import Model
        self.torch = torch
        m = Model(path)
        self.nq, self.nv = m.nq, m.nv
        self.b = SimBatch(m, env_config(ctrl_mode=lib.CTRL_RAW, obs_kind=lib.OBS_STATE, obs_dim=m.nq + m.nv, act_dim=m.nu, frame_skip=1), 2, 0, torch.float64)
        self.b.reset()
        self.a = torch.zeros(2, m.nu, dtype=torch.float64, device="cuda")

    def set_state(self, qpos, qvel):
        t = self.torch
        self.b.set_state(t.tensor(np.tile(qpos, (2, 1)), device="cuda"), t.tensor(np.tile(qvel, (2, 1)), device="cuda"))

    def step(self, n):
        for _ in range(n):
            obs, *_ = self.b.step(self.a)
        o = obs[0].cpu().numpy()
        assert self.torch.equal(obs[0], obs[1])
        return o[:self.nq].copy(), o[self.nq:].copy()


CPU_SIMS = [OracleSim, KernelSourceSim]
SIMS = [pytest.param(OracleSim, id="oracle"), pytest.param(KernelSourceSim, id="kernel-source"), pytest.param(GpuSim, id="cuda-f64", marks=pytest.mark.gpu)]


# ------------------------------------------------------------------------------------------------ model constants (oracle's mj_setConst)
def test_invweights_closed_form(tmp_path, assets):
    m = O.Model(model_path(tmp_path))
    bw = np.asarray(m.arr("body_invweight0")).reshape(-1, 2); dw = np.asarray(m.arr("dof_invweight0"))
    link, box = m.id("body", "link"), m.id("body", "box")
    # hinge: M = I_hinge; the link's centre of mass moves along one direction with lever l
    assert abs(dw[0] - 1.0 / I_HINGE) < 1e-12
    assert abs(bw[link, 0] - L_PEND ** 2 / I_HINGE / 3.0) < 1e-12 and abs(bw[link, 1] - 1.0 / I_HINGE / 3.0) < 1e-12
    # free box: translation 1 / m, rotation mean of the inverse principal inertias of a solid box
    sx, sy, sz = 0.03, 0.02, 0.01
    inertia = M_BOX / 3.0 * np.array([sy * sy + sz * sz, sx * sx + sz * sz, sx * sx + sy * sy])
    assert np.allclose(dw[1:4], 1.0 / M_BOX, rtol=1e-12) and np.allclose(dw[4:7], np.mean(1.0 / inertia), rtol=1e-9)
    assert abs(bw[box, 0] - 1.0 / M_BOX) < 1e-9 and abs(bw[box, 1] - np.mean(1.0 / inertia)) / bw[box, 1] < 1e-9
    assert np.all(bw[0] == 0) and np.all(bw[m.id("body", "base")] == 0)                        # static bodies
    # the reference's gripper tendon (assets/main.xml:349-354): invweight = J M^-1 J^T with J = 0.5 on each driver joint
    g = O.Model(assets + "/ur3e_2f85.xml"); d = O.Data(g); d.reset(); d.forward()
    Minv = np.linalg.inv(d.fullM())
    J = np.zeros(g.nv); J[6] = J[10] = 0.5
    assert abs(np.asarray(g.arr("tendon_invweight0"))[0] - J @ Minv @ J) / (J @ Minv @ J) < 1e-9
    assert np.allclose(np.asarray(g.arr("dof_invweight0")), np.diag(Minv), rtol=1e-9)          # all hinges: the diagonal of M^-1 at qpos0


def test_impedance_and_reference_parameters(tmp_path):
    """d(r), K and B of the box-plane contact rows at prescribed penetrations against the documented formulas; friction rows: K = 0,
    R_t = R_n / impratio, regularised mu = friction * sqrt(R_t / R_n)."""
    for frac, want in ((0.25, 0.90625), (0.5, 0.925), (0.75, 0.94375), (1.0, 0.95), (2.0, 0.95)):
        r = frac * WIDTH
        m = O.Model(model_path(tmp_path, z0=0.01 - r)); d = O.Data(m); d.reset(); d.forward()
        assert d.ncon == 4 and d.nefc == 12
        kbip = np.asarray(d.arr("efc_KBIP")).reshape(-1, 4)[:12]; R = np.asarray(d.arr("efc_R"))[:12]
        assert abs(impedance(r) - want) < 1e-15
        assert np.allclose(kbip[0::3, 2], want, atol=1e-12)                                    # impedance of the normal rows
        assert np.allclose(kbip[0::3, 0], K_REF, rtol=1e-12) and np.allclose(kbip[:, 1], B_REF, rtol=1e-12)
        assert np.all(kbip[1::3, 0] == 0) and np.all(kbip[2::3, 0] == 0)                       # friction rows have no position term
        Rn = (1.0 - want) / want / M_BOX
        assert np.allclose(R[0::3], Rn, rtol=1e-12) and np.allclose(R[1::3], Rn / IMPRATIO, rtol=1e-12) and np.allclose(R[2::3], Rn / IMPRATIO, rtol=1e-12)
        for c in d.contacts()[:4]:
            assert abs(c.mu - 1.0 * np.sqrt(1.0 / IMPRATIO)) < 1e-12 and abs(c.dist + r) < 1e-12
        aref = np.asarray(d.arr("efc_aref"))[:12]
        assert np.allclose(aref[0::3], K_REF * want * r, rtol=1e-9)                             # at rest: aref = -K d (pos - margin)


# ------------------------------------------------------------------------------------------------ trajectories on every implementation
@pytest.mark.parametrize("Sim", SIMS)
def test_free_fall_is_semi_implicit_euler(tmp_path, Sim):
    s = Sim(model_path(tmp_path, z0=0.5))
    m = O.Model(model_path(tmp_path, z0=0.5))
    qp = np.array(m.arr("qpos0")).copy(); qv = np.zeros(m.nv)
    s.set_state(qp, qv)
    n, h = 150, 0.001
    q, v = s.step(n)
    assert abs(q[3] - (0.5 - G * h * h * n * (n + 1) / 2)) < 1e-12 and abs(v[3] + G * h * n) < 1e-12
    assert np.abs(q[1:3]).max() < 1e-15 and np.abs(q[4:8] - [1, 0, 0, 0]).max() < 1e-15


@pytest.mark.parametrize("Sim", SIMS)
def test_pendulum_period_and_energy(tmp_path, Sim):
    s = Sim(model_path(tmp_path, z0=0.5))
    m = O.Model(model_path(tmp_path, z0=0.5))
    qp = np.array(m.arr("qpos0")).copy(); qp[0] = 0.02
    s.set_state(qp, np.zeros(m.nv))
    T = 2 * np.pi * np.sqrt(I_HINGE / (M_PEND * G * L_PEND)) * (1 + 0.02 ** 2 / 16)
    h = 0.001
    th, om = [0.02], [0.0]
    for _ in range(340):
        q, v = s.step(10 if Sim is not KernelSourceSim else 10)
        th.append(q[0]); om.append(v[0])
    th, om = np.array(th), np.array(om)
    t = np.arange(len(th)) * 10 * h
    # downward zero crossings of the angle, linearly interpolated: two consecutive ones are a period apart
    idx = [i for i in range(len(th) - 1) if th[i] > 0 >= th[i + 1]]
    cross = [t[i] + (t[i + 1] - t[i]) * th[i] / (th[i] - th[i + 1]) for i in idx]
    assert len(cross) >= 2
    assert abs((cross[1] - cross[0]) - T) / T < 2e-4
    energy = 0.5 * I_HINGE * om ** 2 + M_PEND * G * L_PEND * (1 - np.cos(th))
    assert np.abs(energy - energy[0]).max() / energy[0] < 5e-3                                  # symplectic Euler: bounded O(h) oscillation, no drift


@pytest.mark.parametrize("Sim", SIMS)
def test_box_at_rest_penetration_and_forces(tmp_path, Sim):
    p = model_path(tmp_path)
    s = Sim(p); m = O.Model(p)
    qp = np.array(m.arr("qpos0")).copy()
    s.set_state(qp, np.zeros(m.nv))
    q, v = s.step(3000)
    r = rest_penetration(G)
    assert abs((0.01 - q[3]) - r) < 1e-11 and np.abs(v[1:]).max() < 1e-10
    if Sim is OracleSim:
        f = np.asarray(s.d.arr("efc_force"))[:12]
        assert np.allclose(f[0::3], M_BOX * G / 4, rtol=1e-9) and np.abs(f[1::3]).max() < 1e-12 and np.abs(f[2::3]).max() < 1e-12


@pytest.mark.parametrize("Sim", SIMS)
@pytest.mark.parametrize("tan_theta", [0.2, 0.5])
def test_sticking_creep_velocity_on_incline(tmp_path, Sim, tan_theta):
    theta = np.arctan(tan_theta)
    p = model_path(tmp_path, theta)
    s = Sim(p); m = O.Model(p)
    s.set_state(np.array(m.arr("qpos0")).copy(), np.zeros(m.nv))
    q, v = s.step(3000)
    # per-contact impedance from the settled corner depths (the incline loads the downhill corners more)
    o = O.Data(m); o.reset(); o.set_state(q, v); o.forward()
    assert o.ncon == 4
    ssum = sum(impedance(c.dist) / (1.0 - impedance(c.dist)) for c in o.contacts()[:4])
    want = G * np.sin(theta) / (IMPRATIO * B_REF * ssum)
    assert abs(v[1] - want) / want < 1e-6 and abs(v[2]) < 1e-12, (v[1], want)
    fn = np.asarray(o.arr("efc_force"))[:12:3]
    assert abs(fn.sum() - M_BOX * G * np.cos(theta)) < 1e-8                                      # normal forces carry the weight's normal component


@pytest.mark.parametrize("Sim", SIMS)
def test_sliding_above_friction_angle(tmp_path, Sim):
    theta = np.arctan(1.5)                                                                       # friction coefficient 1 -> slides
    p = model_path(tmp_path, theta)
    s = Sim(p); m = O.Model(p)
    s.set_state(np.array(m.arr("qpos0")).copy(), np.zeros(m.nv))
    q, v = s.step(1000)
    want = G * (np.sin(theta) - 1.0 * np.cos(theta)) * 1.0
    assert abs(v[1] - want) / want < 0.10, (v[1], want)
    assert v[1] < G * np.sin(theta) * 0.5                                                        # far from frictionless (8.2 m/s)


@pytest.mark.parametrize("source", ["oracle", "kernel-source"])
def test_double_pendulum_mass_matrix_and_bias_textbook(source):
    """CRBA and RNE against the closed-form double pendulum (any robotics text): M11 = I1 + I2 + m1 lc1^2 + m2 (l1^2 + lc2^2 + 2 l1 lc2 c2),
    M12 = I2 + m2 (lc2^2 + l1 lc2 c2), M22 = I2 + m2 lc2^2; bias = Coriolis / centrifugal (h = -m2 l1 lc2 s2) + gravity."""
    path = os.path.join(HERE, "models", "double_pendulum.xml")
    m1, m2, l1, lc1, lc2, I1, I2 = 2.0, 1.5, 0.5, 0.3, 0.25, 0.04, 0.02
    rng = np.random.default_rng(4)
    for _ in range(6):
        q = rng.uniform(-2.5, 2.5, 2); v = rng.uniform(-3, 3, 2)
        c2, s2 = np.cos(q[1]), np.sin(q[1])
        M = np.array([[I1 + I2 + m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * c2), I2 + m2 * (lc2 ** 2 + l1 * lc2 * c2)],
                      [I2 + m2 * (lc2 ** 2 + l1 * lc2 * c2), I2 + m2 * lc2 ** 2]])
        h = -m2 * l1 * lc2 * s2
        grav = np.array([m1 * G * lc1 * np.sin(q[0]) + m2 * G * (l1 * np.sin(q[0]) + lc2 * np.sin(q[0] + q[1])), m2 * G * lc2 * np.sin(q[0] + q[1])])
        bias = np.array([h * (2 * v[0] * v[1] + v[1] ** 2), -h * v[0] ** 2]) + grav
        if source == "oracle":
            mo = O.Model(path); d = O.Data(mo); d.reset(); d.set_state(q, v); d.forward()
            gotM, gotb = d.fullM(), np.asarray(d.qfrc_bias)
        else:
            from tests.hostcheck import build as HC
            r = HC.run(path, q, v, np.zeros(2), np.zeros(2), 0)
            gotM, gotb = r["M"], r["bias"]
        assert np.abs(gotM - M).max() < 1e-12 and np.abs(gotb - bias).max() < 1e-11, (gotM, M, gotb, bias)


class _HingeGpu:
    def __init__(self, path):
        import torch
        import ur3e_b200._lib as lib
        from ur3e_b200.batch import SimBatch, env_config
        from ur3e_b200.model import Model
        self.torch = torch
        self.b = SimBatch(Model(path), env_config(ctrl_mode=lib.CTRL_RAW, obs_kind=lib.OBS_STATE, obs_dim=6, act_dim=3, frame_skip=1), 1, 0, torch.float64)
        self.b.reset()

    def run(self, ctrl, n):
        a = self.torch.tensor(np.asarray(ctrl, dtype=float)[None], device="cuda")
        for _ in range(n):
            obs, *_ = self.b.step(a)
        o = obs[0].cpu().numpy()
        return o[:3], o[3:]


def _hinge_run(Sim, ctrl, n):
    path = os.path.join(HERE, "models", "hinge_pins.xml")
    if Sim is GpuSim:
        return _HingeGpu(path).run(ctrl, n)
    if Sim is KernelSourceSim:
        from tests.hostcheck import build as HC
        r = HC.run(path, np.zeros(3), np.zeros(3), np.asarray(ctrl, dtype=float), np.zeros(3), n)
        return r["qpos"], r["qvel"]
    m = O.Model(path); d = O.Data(m); d.reset(); d.ctrl[:] = ctrl; d.step(n)
    return d.qpos.copy(), d.qvel.copy()


@pytest.mark.parametrize("Sim", SIMS)
def test_joint_limit_frictionloss_and_affine_actuator(Sim):
    inertia, a_hat = 0.5, 2.0
    q, v = _hinge_run(Sim, [2.0, 0.1, 10.0], 4000)
    # joint limit: constant torque 2 against the upper stop at 0.5
    r = 1e-4
    for _ in range(200):
        d = impedance(r); r = 2.0 * (1.0 - d) * a_hat / (K_REF * d * d)
    assert abs((q[0] - 0.5) - r) < 1e-11 and abs(v[0]) < 1e-10
    # friction loss 0.2 under torque 0.1: sticking, with the soft row's creep
    assert abs(v[1] - 0.1 * (1.0 - DMIN) * a_hat / (DMIN * B_REF)) < 1e-12
    # affine actuator: 0.3137255 * ctrl - 100 q - 10 qdot = 0 at rest
    assert abs(q[2] - 0.3137255 * 10.0 / 100.0) < 1e-10 and abs(v[2]) < 1e-10
    q, v = _hinge_run(Sim, [0.0, 0.5, 255.0], 2000)
    assert abs(v[1] / 2.0 - (0.5 - 0.2) / inertia) < 1e-9                                        # sliding: Coulomb torque subtracted
    assert abs(q[2] - 0.3137255 * 255.0 / 100.0) < 1e-6                                          # force range (+-5) only limits the approach
    assert abs(q[0]) < 1e-15 and abs(v[0]) < 1e-15


@pytest.mark.parametrize("Sim", SIMS)
def test_gripper_closure_residual(assets, Sim):
    """2F85 four-bar loops (connect equalities, reference assets/ur3e_2f85.xml:294-298) under full closing force: the soft constraints
    hold the loop closed to a fraction of a millimetre and the two drivers stay mirrored (joint equality)."""
    p = assets + "/ur3e_2f85.xml"
    m = O.Model(p)
    if Sim is GpuSim:
        import torch
        import ur3e_b200._lib as lib
        from ur3e_b200.batch import SimBatch, env_config
        from ur3e_b200.model import Model
        b = SimBatch(Model(p), env_config(ctrl_mode=lib.CTRL_RAW, obs_kind=lib.OBS_STATE, obs_dim=28, act_dim=7, frame_skip=1, reset_key=1), 1, 0, torch.float64)
        b.reset()
        a = torch.zeros(1, 7, dtype=torch.float64, device="cuda"); a[0, 6] = 255.0
        for _ in range(800):
            obs, *_ = b.step(a)
        q = obs[0, :14].cpu().numpy()
    elif Sim is KernelSourceSim:
        from tests.hostcheck import build as HC
        qp, qv = m.key("down"); u = np.zeros(7); u[6] = 255.0
        q = HC.run(p, qp, qv, u, np.zeros(m.nv), 800)["qpos"]
    else:
        d = O.Data(m); qp, qv = m.key("down"); d.reset(); d.set_state(qp, qv); d.ctrl[:] = 0; d.ctrl[6] = 255.0; d.step(800); q = d.qpos.copy()
    d = O.Data(m); d.reset(); d.set_state(q, np.zeros(m.nv)); d.forward()
    xpos, xmat = d.xpos.reshape(-1, 3), d.xmat.reshape(-1, 3, 3)
    eq = np.asarray(m.arr("eq_data")).reshape(m.neq, -1)
    for e, (b1, b2) in enumerate((("right_follower", "right_coupler"), ("left_follower", "left_coupler"))):
        i1, i2 = m.id("body", b1), m.id("body", b2)
        p1 = xpos[i1] + xmat[i1] @ eq[e, 0:3]; p2 = xpos[i2] + xmat[i2] @ eq[e, 3:6]
        assert np.linalg.norm(p1 - p2) < 3e-4, (b1, np.linalg.norm(p1 - p2))
    assert abs(q[6] - q[10]) < 2e-3 and q[6] > 0.5                                               # closed, mirrored
