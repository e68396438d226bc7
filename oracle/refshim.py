"""ORACLE (test infrastructure, NOT product code): stand-ins for the third-party modules the reference
imports (`mujoco`, `gymnasium`, `matplotlib`), backed by the float64 oracle, so that the reference's own
Python -- controller/controller_func.py, utils/utils.py, utils/gym_utils.py, gymnasium_env/envs/*.py --
runs VERBATIM in the build container (none of those packages is installable there).

Used only by tools/make_golden.py to generate tests/golden/*.npz.  The MuJoCo API subset is the one listed
in SURVEY 8(b)-2.  Known reference breakages that are shimmed rather than fixed:
`utils.utils.get_joint_torques` (missing alias of get_jnt_torques).
"""
import os
import sys
import types

import numpy as np

from . import oracle as O

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ur3e_b200", "assets")


class _Opt:
    def __init__(self, ts): self.timestep = ts


class _Key:
    def __init__(self, qpos, qvel): self.qpos, self.qvel = qpos, qvel


class MjModel:
    def __init__(self, path):
        # the reference's XML needs meshes that are not shipped: load the mesh-stripped twin
        self._o = O.Model(os.path.join(ASSETS, os.path.basename(path)))
        p = self._o.py
        for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite"):
            setattr(self, k, int(p[k]))
        self.opt = _Opt(p["opt"]["timestep"])
        self.actuator_ctrlrange = np.array(p["actuator_ctrlrange"]).reshape(-1, 2)
        self.jnt_range = np.array(p["jnt_range"]).reshape(-1, 2)
        self.geom_bodyid = np.array(p["geom_bodyid"]); self.geom_size = np.array(p["geom_size"]).reshape(-1, 3)
        self.body_parentid = np.array(p["body_parentid"]); self.body_mass = np.array(p["body_mass"])

    @staticmethod
    def from_xml_path(path): return MjModel(path)

    def keyframe(self, name):
        qp, qv = self._o.key(name)
        return _Key(qp, qv)


class _Named:
    def __init__(self, **kw): self.__dict__.update(kw)


class _Contact:
    def __init__(self, c): self.geom1, self.geom2, self.dist = c.geom1, c.geom2, c.dist


class MjData:
    def __init__(self, m):
        self._m = m; self._o = O.Data(m._o)

    # assignable state vectors (the reference does `d.qpos = init_qpos`, `d.ctrl = u`)
    def _get(self, n): return self._o.arr(n)
    qpos = property(lambda s: s._get("qpos"), lambda s, v: s._get("qpos").__setitem__(slice(None), v))
    qvel = property(lambda s: s._get("qvel"), lambda s, v: s._get("qvel").__setitem__(slice(None), v))
    ctrl = property(lambda s: s._get("ctrl"), lambda s, v: s._get("ctrl").__setitem__(slice(None), v))
    qacc = property(lambda s: s._get("qacc"))
    qacc_warmstart = property(lambda s: s._get("qacc_warmstart"))
    qfrc_bias = property(lambda s: s._get("qfrc_bias"))
    xpos = property(lambda s: s._get("xpos").reshape(-1, 3))
    xmat = property(lambda s: s._get("xmat").reshape(-1, 9))
    cvel = property(lambda s: s._get("cvel").reshape(-1, 6))
    ncon = property(lambda s: s._o.ncon)
    contact = property(lambda s: [_Contact(c) for c in s._o.contacts()])
    time = property(lambda s: s._o.time)

    def site(self, i): return _Named(xpos=self._get("site_xpos").reshape(-1, 3)[i], xmat=self._get("site_xmat").reshape(-1, 9)[i])
    def geom(self, i): return _Named(xpos=self._get("geom_xpos").reshape(-1, 3)[i], xmat=self._get("geom_xmat").reshape(-1, 9)[i])
    def sensor(self, name): return _Named(data=np.zeros(3))   # sensors are logging-only in the reference


class _mjtObj:
    mjOBJ_BODY, mjOBJ_JOINT, mjOBJ_GEOM, mjOBJ_SITE = 1, 3, 5, 6


_KIND = {1: "body", 3: "joint", 5: "geom", 6: "site"}


def _name2id(m, t, name):
    try:
        return m._o.names[_KIND[t]].index(name)
    except ValueError:
        return -1


def _jac_site(m, d, jacp, jacr, sid):
    jp, jr = d._o.jac_site(sid)
    if jacp is not None: jacp[:] = jp
    if jacr is not None: jacr[:] = jr


def _object_velocity(m, d, objtype, oid, res, flg_local=0):
    assert objtype == _mjtObj.mjOBJ_SITE
    res[:] = d._o.site_velocity(oid, flg_local)


def _step(m, d, nstep=1): d._o.step(nstep)


def install():
    """Put the stand-in modules into sys.modules (idempotent)."""
    if "mujoco" in sys.modules and getattr(sys.modules["mujoco"], "_ur3e_shim", False):
        return
    mj = types.ModuleType("mujoco"); mj._ur3e_shim = True
    mj.MjModel, mj.MjData, mj.mjtObj = MjModel, MjData, _mjtObj
    mj.mj_resetData = lambda m, d: d._o.reset()
    mj.mj_forward = lambda m, d: d._o.forward()
    mj.mj_step = _step
    mj.mj_name2id = _name2id
    mj.mj_id2name = lambda m, t, i: m._o.names[_KIND[t]][i]
    mj.mj_jacSite = _jac_site
    mj.mj_objectVelocity = _object_velocity
    mj.mj_rnePostConstraint = lambda m, d: None
    viewer = types.ModuleType("mujoco.viewer"); mj.viewer = viewer
    sys.modules["mujoco"] = mj; sys.modules["mujoco.viewer"] = viewer

    # matplotlib: plotting is out of scope; the reference imports it at module load
    mpl = types.ModuleType("matplotlib"); mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    tk = types.ModuleType("mpl_toolkits"); tk3 = types.ModuleType("mpl_toolkits.mplot3d"); tk3.Axes3D = object
    mpl.pyplot = plt
    for n, mod in (("matplotlib", mpl), ("matplotlib.pyplot", plt), ("mpl_toolkits", tk), ("mpl_toolkits.mplot3d", tk3)):
        sys.modules.setdefault(n, mod)

    # gymnasium: spaces.Box + MujocoEnv glue (do_simulation / set_state / reset)
    gym = types.ModuleType("gymnasium"); spaces = types.ModuleType("gymnasium.spaces")
    envs = types.ModuleType("gymnasium.envs"); gmj = types.ModuleType("gymnasium.envs.mujoco")

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float64):
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy() if shape is not None else np.asarray(low, dtype=dtype)
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy() if shape is not None else np.asarray(high, dtype=dtype)
            self.shape = self.low.shape; self.dtype = dtype

    class MujocoEnv:
        def __init__(self, model_path, frame_skip, observation_space, render_mode=None, **kw):
            self.model = MjModel(model_path); self.data = MjData(self.model)
            self.frame_skip = frame_skip; self.observation_space = observation_space; self.render_mode = render_mode
            self.init_qpos = self.data.qpos.copy(); self.init_qvel = self.data.qvel.copy()

        def do_simulation(self, ctrl, n_frames):
            if np.array(ctrl).shape != (self.model.nu,):
                raise ValueError("Action dimension mismatch")
            self.data.ctrl[:] = ctrl
            _step(self.model, self.data, n_frames)

        def set_state(self, qpos, qvel):
            assert qpos.shape == (self.model.nq,) and qvel.shape == (self.model.nv,)
            self.data.qpos[:] = np.copy(qpos); self.data.qvel[:] = np.copy(qvel)
            self.data._o.forward()

        def reset(self, *, seed=None, options=None):
            self.data._o.reset()
            return self.reset_model(), {}

        def render(self): return None
        def close(self): pass

    spaces.Box = Box; gym.spaces = spaces; gym.envs = envs; envs.mujoco = gmj; gmj.MujocoEnv = MujocoEnv
    gym.register = lambda *a, **k: None; gym.Env = object
    for n, mod in (("gymnasium", gym), ("gymnasium.spaces", spaces), ("gymnasium.envs", envs), ("gymnasium.envs.mujoco", gmj)):
        sys.modules[n] = mod


def import_reference(ref_root="/root/reference"):
    """chdir into the reference (its code opens ./assets and controller/config by relative path), apply the F5 alias,
    and return its modules."""
    install()
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    os.chdir(ref_root)
    import utils.utils as uu
    if not hasattr(uu, "get_joint_torques"):
        uu.get_joint_torques = uu.get_jnt_torques      # controller/move_l_task.py:7 imports a name that does not exist
    import controller.controller_func as cf
    import utils.gym_utils as gu
    return uu, cf, gu
