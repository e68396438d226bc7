"""TEST-ONLY helper: compile ur3e_b200/csrc/engine.cuh as plain C++ (see host_engine.cpp) and bind it."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libhostcheck.so")
_L = None


def lib():
    global _L
    if _L is None:
        srcs = [os.path.join(HERE, "host_engine.cpp"), os.path.join(ROOT, "ur3e_b200", "csrc", "mjcf.cpp")]
        deps = srcs + [os.path.join(ROOT, "ur3e_b200", "csrc", f) for f in ("engine.cuh", "warp_model.cuh", "dev_model.h", "compile_model.h", "host_model.h", "xml_mini.h", "merge_bodies.h")]
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wno-unused", "-ffp-contract=off", "-o", SO] + srcs)
        _L = C.CDLL(SO)
    return _L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def run(xml, qpos, qvel, ctrl, ws, nsteps, use_float=0, max_iter=50, tol=1e-15):
    L = lib()
    dims = np.zeros(11, dtype=np.int32)
    assert L.hc_model_dims(xml.encode(), dims.ctypes.data_as(C.POINTER(C.c_int))) == 0
    nq, nv = int(dims[0]), int(dims[1])
    oq = np.zeros(nq); ov = np.zeros(nv); oa = np.zeros(nv); oM = np.zeros((nv, nv)); ob = np.zeros(nv); ofc = np.zeros(nv); info = np.zeros(8, dtype=np.int32)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (qpos, qvel, ctrl, ws)]
    rc = L.hc_run(xml.encode(), use_float, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), nsteps, max_iter, C.c_double(tol), _p(oq), _p(ov), _p(oa), _p(oM), _p(ob), _p(ofc),
                  info.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return dict(qpos=oq, qvel=ov, qacc=oa, M=oM, bias=ob, fc=ofc, ncon=int(info[0]), nefc=int(info[1]), iters=int(info[2]), warn=int(info[3]), overflow=int(info[4]), arena=int(info[5]))


def sensors(xml, qpos, qvel, ctrl):
    """The step kernel's logging record (include/ur3e_b200.h UR3E_NSENSOR) after a forward pass at the given state."""
    out = np.zeros(46)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (qpos, qvel, ctrl)]
    assert lib().hc_sensors(xml.encode(), _p(a[0]), _p(a[1]), _p(a[2]), _p(out)) == 0
    return out


def model_array(xml, name, cap=8192):
    out = np.zeros(cap)
    k = lib().hc_model_array(xml.encode(), name.encode(), _p(out), cap)
    assert k >= 0, name
    return out[:k]
