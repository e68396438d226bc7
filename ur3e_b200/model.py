"""Model handle: the C++ MJCF loader's flat arrays behind MuJoCo's names.

Replaces `mujoco.MjModel.from_xml_path` for the reference's scenes (reference utils/utils.py:9-12) and
the `m.*` reads its Python performs: nq/nv/nu, opt.timestep, actuator_ctrlrange,
jnt_range, geom_bodyid, geom_size, body_parentid, body_mass, keyframes, name<->id lookups.
"""
import ctypes as C
import os

import numpy as np

from . import _lib

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def asset(name):
    """Path of a bundled (mesh-stripped) scene: 'main.xml', 'ur3e_2f85.xml', 'ur3e_raw.xml'."""
    p = os.path.join(ASSET_DIR, name)
    if not os.path.exists(p):
        raise FileNotFoundError(p)
    return p


class _Opt:
    def __init__(self, timestep):
        self.timestep = timestep


class _Key:
    def __init__(self, qpos, qvel):
        self.qpos, self.qvel = qpos, qvel


class Model:
    def __init__(self, xml_path):
        L = _lib.load()
        self._L = L
        self.path = os.path.abspath(xml_path)
        self.ptr = L.ur3e_model_load(self.path.encode())
        if not self.ptr:
            raise ValueError("ur3e_model_load: " + _lib.last_error())
        d = _lib.ModelDims()
        _lib.check(L.ur3e_model_info(self.ptr, C.byref(d)), "ur3e_model_info")
        for n, _ in _lib.ModelDims._fields_:
            setattr(self, n, getattr(d, n))
        self.opt = _Opt(d.timestep)
        self.warnings = [L.ur3e_model_warning(self.ptr, i).decode() for i in range(L.ur3e_model_num_warnings(self.ptr))]
        self._arrays = {}

    @classmethod
    def from_xml_path(cls, path):
        return cls(path)

    def array(self, field):
        """Read-only numpy view of a model array by its mjModel name."""
        if field not in self._arrays:
            ptr = C.c_void_p(); shape = (C.c_int64 * 2)(); nd = C.c_int(); is_int = C.c_int()
            rc = self._L.ur3e_model_array(self.ptr, field.encode(), C.byref(ptr), shape, C.byref(nd), C.byref(is_int))
            if rc != 0:
                raise AttributeError(field)
            shp = tuple(int(shape[i]) for i in range(nd.value))
            n = int(np.prod(shp)) if shp else 0
            if n == 0:
                a = np.zeros(shp, dtype=np.int32 if is_int.value else np.float64)
            else:
                ct = C.c_int32 if is_int.value else C.c_double
                a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).reshape(shp)
            a.flags.writeable = False
            self._arrays[field] = a
        return self._arrays[field]

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return self.array(name)

    def name2id(self, objtype, name):
        return self._L.ur3e_model_name2id(self.ptr, objtype, name.encode())

    def id2name(self, objtype, i):
        s = self._L.ur3e_model_id2name(self.ptr, objtype, i)
        return s.decode() if s is not None else None

    def body_id(self, name): return self.name2id(_lib.OBJ_BODY, name)
    def site_id(self, name): return self.name2id(_lib.OBJ_SITE, name)
    def joint_id(self, name): return self.name2id(_lib.OBJ_JOINT, name)
    def geom_id(self, name): return self.name2id(_lib.OBJ_GEOM, name)

    def key_id(self, name):
        k = self.name2id(_lib.OBJ_KEY, name)
        if k < 0:
            raise KeyError("no keyframe named %r" % name)
        return k

    def keyframe(self, name):
        """m.keyframe(name).qpos/.qvel (reference utils/utils.py:15-24, gym_utils.py:63-79); copies."""
        k = self.key_id(name)
        return _Key(np.array(self.array("key_qpos")[k]), np.array(self.array("key_qvel")[k]))

    def close(self):
        if getattr(self, "ptr", None):
            self._L.ur3e_model_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
