"""ORACLE (test infrastructure, NOT product code): ctypes front-end of libur3e_oracle.so.

`Model(xml_path)` parses the MJCF with oracle/mjcf_model.py, uploads the arrays into the C
oracle and runs mj_setConst's restatement; `Data(model)` exposes the MuJoCo-named arrays as
numpy views.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
Parity status: UNPINNED against a real MuJoCo build (see ur3e_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import mjcf_model

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

INT_FIELDS = """body_parentid body_rootid body_weldid body_jntadr body_jntnum body_dofadr body_dofnum
jnt_type jnt_bodyid jnt_qposadr jnt_dofadr jnt_limited dof_bodyid dof_jntid dof_parentid geom_type geom_bodyid
site_bodyid tendon_adr tendon_num wrap_jnt eq_type eq_obj1id eq_obj2id actuator_trntype actuator_trnid
actuator_ctrllimited actuator_forcelimited pair_geom1 pair_geom2 pair_condim""".split()
DBL_FIELDS = """body_pos body_quat body_ipos body_iquat body_mass body_inertia body_invweight0 jnt_pos jnt_axis
jnt_range jnt_stiffness jnt_margin jnt_solref jnt_solimp qpos0 qpos_spring dof_armature dof_damping dof_frictionloss
dof_invweight0 dof_solref dof_solimp geom_pos geom_quat geom_size site_pos site_quat wrap_coef tendon_invweight0
eq_data eq_solref eq_solimp actuator_gainprm actuator_biasprm actuator_ctrlrange actuator_forcerange actuator_gear
pair_friction pair_solref pair_solimp pair_margin pair_gap key_qpos key_qvel""".split()


class OContact(C.Structure):
    _fields_ = [("dist", C.c_double), ("pos", C.c_double * 3), ("frame", C.c_double * 9), ("friction", C.c_double * 5),
                ("solref", C.c_double * 2), ("solimp", C.c_double * 5), ("mu", C.c_double), ("includemargin", C.c_double),
                ("geom1", C.c_int), ("geom2", C.c_int), ("dim", C.c_int), ("efc_address", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "_build", "libur3e_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("ur3e_oracle.c", "ur3e_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, cp, ip, dp = C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
        L.o_model_new.restype = vp; L.o_model_new.argtypes = [ip, dp]
        L.o_model_set_int.argtypes = [vp, cp, ip, C.c_int]; L.o_model_set_dbl.argtypes = [vp, cp, dp, C.c_int]
        L.o_model_get_dbl.argtypes = [vp, cp, C.POINTER(dp)]
        L.o_model_free.argtypes = [vp]; L.o_set_const.argtypes = [vp]
        L.o_model_meaninertia.restype = C.c_double; L.o_model_meaninertia.argtypes = [vp]
        L.o_data_new.restype = vp; L.o_data_new.argtypes = [vp]; L.o_data_free.argtypes = [vp]
        L.o_data_get_dbl.argtypes = [vp, vp, cp, C.POINTER(dp), ip]; L.o_data_get_int.argtypes = [vp, vp, cp, C.POINTER(ip), ip]
        L.o_data_contacts.restype = C.POINTER(OContact); L.o_data_contacts.argtypes = [vp]
        L.o_data_info.argtypes = [vp, C.c_int]; L.o_data_time.restype = C.c_double; L.o_data_time.argtypes = [vp]
        for f in ("o_reset_data", "o_forward", "o_step"):
            getattr(L, f).argtypes = [vp, vp]
        L.o_step_n.argtypes = [vp, vp, C.c_int]
        L.o_jac.argtypes = [vp, vp, dp, dp, dp, C.c_int]; L.o_jac_site.argtypes = [vp, vp, dp, dp, C.c_int]
        L.o_object_velocity_site.argtypes = [vp, vp, C.c_int, dp, C.c_int]
        L.o_fullM.argtypes = [vp, vp, dp]
        L.o_pid_task_ctrl.argtypes = [vp, vp, C.c_int, dp, dp, dp]; L.o_pd_joint_ctrl.argtypes = [vp, vp, dp, dp, dp, dp]
        L.o_rot_err.argtypes = [dp, dp, dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Model:
    def __init__(self, xml_path):
        L = lib()
        m = mjcf_model.load_mjcf(xml_path)
        self.py = m
        self.names = m["names"]
        pc = m["pair_candidates"]
        m["npair"] = len(pc)
        m["pair_geom1"] = np.array([r["g1"] for r in pc], dtype=np.int32)
        m["pair_geom2"] = np.array([r["g2"] for r in pc], dtype=np.int32)
        m["pair_condim"] = np.array([r["condim"] for r in pc], dtype=np.int32)
        m["pair_friction"] = np.array([r["friction"] for r in pc], dtype=np.float64).reshape(len(pc), 5)
        m["pair_solref"] = np.array([r["solref"] for r in pc], dtype=np.float64).reshape(len(pc), 2)
        m["pair_solimp"] = np.array([r["solimp"] for r in pc], dtype=np.float64).reshape(len(pc), 5)
        m["pair_margin"] = np.array([r["margin"] for r in pc], dtype=np.float64)
        m["pair_gap"] = np.array([r["gap"] for r in pc], dtype=np.float64)
        m["nwrap"] = len(m["wrap_jnt"])
        m["body_invweight0"] = np.zeros(2 * m["nbody"]); m["dof_invweight0"] = np.zeros(m["nv"])
        m["tendon_invweight0"] = np.zeros(m["ntendon"])
        sizes = np.array([m[k] for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "neq", "ntendon", "nwrap", "npair", "nkey")], dtype=np.int32)
        o = m["opt"]
        opt = np.array([o["timestep"], *o["gravity"], o["impratio"], o["tolerance"], o["ls_tolerance"],
                        1.0 if o["cone"] == "elliptic" else 0.0, o["iterations"], o["ls_iterations"]], dtype=np.float64)
        self.ptr = L.o_model_new(_ip(sizes), _dp(opt))
        for f in INT_FIELDS:
            a = np.ascontiguousarray(m[f], dtype=np.int32).ravel()
            rc = L.o_model_set_int(self.ptr, f.encode(), _ip(a), a.size)
            assert rc == 0, (f, rc, a.size)
        for f in DBL_FIELDS:
            a = np.ascontiguousarray(m[f], dtype=np.float64).ravel()
            rc = L.o_model_set_dbl(self.ptr, f.encode(), _dp(a), a.size)
            assert rc == 0, (f, rc, a.size)
        L.o_set_const(self.ptr)
        for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "neq", "ntendon", "npair", "nkey"):
            setattr(self, k, int(m[k]))
        self.timestep = o["timestep"]
        self.meaninertia = L.o_model_meaninertia(self.ptr)

    def arr(self, name):
        """numpy view of a float64 model array held by the C oracle (post set_const)."""
        p = C.POINTER(C.c_double)()
        n = lib().o_model_get_dbl(self.ptr, name.encode(), C.byref(p))
        assert n >= 0, name
        return np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n]

    def id(self, kind, name):
        return self.names[kind].index(name)

    def key(self, name):
        k = self.names["key"].index(name)
        return self.py["key_qpos"][k].copy(), self.py["key_qvel"][k].copy()

    def __del__(self):
        try:
            lib().o_model_free(self.ptr)
        except Exception:
            pass


class Data:
    def __init__(self, model):
        self.m = model
        self.ptr = lib().o_data_new(model.ptr)
        self._cache = {}

    def arr(self, name):
        if name not in self._cache:
            p = C.POINTER(C.c_double)(); n = C.c_int()
            if lib().o_data_get_dbl(self.m.ptr, self.ptr, name.encode(), C.byref(p), C.byref(n)) == 0:
                self._cache[name] = np.ctypeslib.as_array(p, shape=(max(n.value, 1),))[:n.value]
            else:
                q = C.POINTER(C.c_int)()
                assert lib().o_data_get_int(self.m.ptr, self.ptr, name.encode(), C.byref(q), C.byref(n)) == 0, name
                self._cache[name] = np.ctypeslib.as_array(q, shape=(max(n.value, 1),))[:n.value]
        return self._cache[name]

    def __getattr__(self, name):
        if name.startswith("_") or name in ("m", "ptr"):
            raise AttributeError(name)
        return self.arr(name)

    # scalars
    @property
    def ncon(self): return lib().o_data_info(self.ptr, 0)
    @property
    def nefc(self): return lib().o_data_info(self.ptr, 1)
    @property
    def solver_iter(self): return lib().o_data_info(self.ptr, 2)
    @property
    def warn_bad(self): return lib().o_data_info(self.ptr, 3)
    @property
    def time(self): return lib().o_data_time(self.ptr)

    def contacts(self):
        c = lib().o_data_contacts(self.ptr)
        return [c[i] for i in range(self.ncon)]

    def reset(self): lib().o_reset_data(self.m.ptr, self.ptr)
    def forward(self): lib().o_forward(self.m.ptr, self.ptr)
    def step(self, n=1): lib().o_step_n(self.m.ptr, self.ptr, n)

    def sensors(self):
        """main.xml's logging sensors after forward()/step(): 7 x actuatorfrc (= actuator_force) and the two touch sensors
        right_pad1_contact, left_pad1_contact (reference assets/main.xml:392-408; readers utils/utils.py:201-245).
        Touch (MuJoCo's mjSENS_TOUCH, restated): sum of the normal forces of the contacts that involve the site's body, have a
        positive normal force, and whose line through the contact point along the contact normal crosses the site's box."""
        m = self.m
        out = np.zeros(9)
        out[:min(m.nu, 7)] = self.arr("actuator_force")[:7]
        gb = m.py["geom_bodyid"]; sb = m.py["site_bodyid"]; ssz = m.py["site_size"]
        sx = self.arr("site_xpos").reshape(-1, 3); sm = self.arr("site_xmat").reshape(-1, 3, 3); f = self.arr("efc_force")
        for k, name in enumerate(("right_pad1_site", "left_pad1_site")):
            if name not in m.names["site"]:
                continue
            j = m.names["site"].index(name)
            for c in self.contacts():
                if c.efc_address < 0 or (gb[c.geom1] != sb[j] and gb[c.geom2] != sb[j]):
                    continue
                fn = f[c.efc_address]
                if fn <= 0:
                    continue
                o = sm[j].T @ (np.array(c.pos[:]) - sx[j]); d = sm[j].T @ np.array(c.frame[:3])
                lo, hi = -np.inf, np.inf          # slab test of the (two-sided) line o + t d against the box |x_i| <= size_i
                for i in range(3):
                    if abs(d[i]) < 1e-12:
                        if abs(o[i]) > ssz[j][i]:
                            lo, hi = 1.0, 0.0
                    else:
                        t1, t2 = (-ssz[j][i] - o[i]) / d[i], (ssz[j][i] - o[i]) / d[i]
                        lo, hi = max(lo, min(t1, t2)), min(hi, max(t1, t2))
                if lo <= hi:
                    out[7 + k] += fn
        return out

    def torque_sensors(self):
        """The <torque> site sensors of main.xml:384-391 after forward()/step() (MuJoCo's mj_rnePostConstraint + mjSENS_TORQUE, restated):
        cacc from qacc, cfrc_int = subtree sum of (I cacc + v x* I v - external wrenches from contacts and connect equalities), torque
        moved from the tree's reference point to the site and rotated into the site frame.  [n_sensors, 3].
        Reads qvel, so call it after forward() (MuJoCo evaluates sensors inside the forward pass, before integrating), not after step()."""
        m = self.m; nb, nv = m.nbody, m.nv
        par = m.py["body_parentid"]; root = m.py["body_rootid"]; dofbody = m.py["dof_bodyid"]; gb = m.py["geom_bodyid"]
        cin = self.arr("cinert").reshape(nb, 10); cvel = self.arr("cvel").reshape(nb, 6); cdof = self.arr("cdof").reshape(nv, 6); cdd = self.arr("cdof_dot").reshape(nv, 6)
        com = self.arr("subtree_com").reshape(nb, 3); xpos = self.arr("xpos").reshape(nb, 3); xmat = self.arr("xmat").reshape(nb, 3, 3)
        qacc, qvel, f = self.arr("qacc"), self.arr("qvel"), self.arr("efc_force")

        def mul_inert(i, v):
            I = np.array([[i[0], i[3], i[4]], [i[3], i[1], i[5]], [i[4], i[5], i[2]]]); h = i[6:9]; mass = i[9]
            return np.hstack([I @ v[:3] + np.cross(h, v[3:]), mass * v[3:] - np.cross(h, v[:3])])

        cacc = np.zeros((nb, 6)); cacc[0, 3:] = -np.asarray(m.py["opt"]["gravity"], dtype=float)
        cfrc = np.zeros((nb, 6))
        for b in range(1, nb):
            cacc[b] = cacc[par[b]]
            for d in range(nv):
                if dofbody[d] == b:
                    cacc[b] = cacc[b] + cdd[d] * qvel[d] + cdof[d] * qacc[d]
            Iv = mul_inert(cin[b], cvel[b])
            cfrc[b] = mul_inert(cin[b], cacc[b]) + np.hstack([np.cross(cvel[b, :3], Iv[:3]) + np.cross(cvel[b, 3:], Iv[3:]), np.cross(cvel[b, :3], Iv[3:])])

        def ext(b, p, F, sign):
            if b > 0:
                cfrc[b, :3] -= sign * np.cross(p - com[root[b]], F); cfrc[b, 3:] -= sign * F
        for c in self.contacts():
            if c.efc_address < 0:
                continue
            fr = np.array(c.frame[:]).reshape(3, 3); F = fr.T @ f[c.efc_address:c.efc_address + 3]; p = np.array(c.pos[:])
            ext(gb[c.geom2], p, F, 1.0); ext(gb[c.geom1], p, F, -1.0)
        et, eid = self.arr("efc_type"), self.arr("efc_id")
        eq = np.asarray(m.arr("eq_data")).reshape(m.neq, -1) if m.neq else np.zeros((0, 11))
        r = 0
        while r < self.nefc:
            if et[r] == 0 and m.py["eq_type"][eid[r]] == 0:        # connect equality: three rows along the world axes
                e = eid[r]; b1, b2 = m.py["eq_obj1id"][e], m.py["eq_obj2id"][e]
                F = f[r:r + 3].copy()
                ext(b1, xpos[b1] + xmat[b1] @ eq[e, 0:3], F, 1.0); ext(b2, xpos[b2] + xmat[b2] @ eq[e, 3:6], F, -1.0)
                r += 3
            else:
                r += 1
        for b in range(nb - 1, 0, -1):
            if par[b] > 0:
                cfrc[par[b]] += cfrc[b]
        sx = self.arr("site_xpos").reshape(-1, 3); sm = self.arr("site_xmat").reshape(-1, 3, 3)
        out = []
        for sid in m.py["sensor_torque_site"]:
            b = m.py["site_bodyid"][sid]
            tau = cfrc[b, :3] - np.cross(sx[sid] - com[root[b]], cfrc[b, 3:])
            out.append(sm[sid].T @ tau)
        return np.array(out).reshape(-1, 3)

    def set_state(self, qpos, qvel):
        self.arr("qpos")[:] = qpos; self.arr("qvel")[:] = qvel

    def fullM(self):
        nv = self.m.nv; out = np.zeros((nv, nv)); lib().o_fullM(self.m.ptr, self.ptr, _dp(out)); return out

    def jac_site(self, site):
        nv = self.m.nv; jp = np.zeros((3, nv)); jr = np.zeros((3, nv))
        lib().o_jac_site(self.m.ptr, self.ptr, _dp(jp), _dp(jr), site); return jp, jr

    def jac(self, point, body):
        nv = self.m.nv; jp = np.zeros((3, nv)); jr = np.zeros((3, nv)); pt = np.ascontiguousarray(point, dtype=np.float64)
        lib().o_jac(self.m.ptr, self.ptr, _dp(jp), _dp(jr), _dp(pt), body); return jp, jr

    def site_velocity(self, site, local=0):
        out = np.zeros(6); lib().o_object_velocity_site(self.m.ptr, self.ptr, site, _dp(out), local); return out

    def pid_task_ctrl(self, tcp_site, traj7, gains12):
        u = np.zeros(7); t = np.ascontiguousarray(traj7, dtype=np.float64); g = np.ascontiguousarray(gains12, dtype=np.float64)
        lib().o_pid_task_ctrl(self.m.ptr, self.ptr, tcp_site, _dp(t), _dp(g), _dp(u)); return u

    def pd_joint_ctrl(self, target6, kp6, kd6):
        u = np.zeros(6)
        a, b, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (target6, kp6, kd6))
        lib().o_pd_joint_ctrl(self.m.ptr, self.ptr, _dp(a), _dp(b), _dp(c), _dp(u)); return u

    def __del__(self):
        try:
            lib().o_data_free(self.ptr)
        except Exception:
            pass


def rot_err(xmat9, rotvec3):
    out = np.zeros(3); a = np.ascontiguousarray(xmat9, dtype=np.float64).ravel(); b = np.ascontiguousarray(rotvec3, dtype=np.float64)
    lib().o_rot_err(_dp(a), _dp(b), _dp(out)); return out
