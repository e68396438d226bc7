"""SimBatch: N environments resident on one B200, stepped by one fused kernel launch.

Thin torch front-end over the C ABI (include/ur3e_b200.h): tensors are passed by `data_ptr()`, zero-copy;
torch supplies device memory and streams only.  dtype float32 = production, float64 = validation build.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .model import Model


def env_config(**kw):
    c = _lib.EnvConfig()
    c.frame_skip = 1; c.reset_key = -1; c.auto_reset = 0
    gains = kw.pop("gains", None)
    rotvec = kw.pop("tool_rotvec", (-1.209, -1.209, 1.209))   # reference ur3e_env2.py:74
    for k, v in kw.items():
        if not hasattr(c, k):
            raise TypeError("unknown env_config field %r" % k)
        setattr(c, k, v)
    if gains is not None:
        g = list(np.asarray(gains, dtype=np.float64).ravel())
        if len(g) > 24:
            raise ValueError("at most 24 gains")
        for i, v in enumerate(g):
            c.gains[i] = v
    for i in range(3):
        c.tool_rotvec[i] = rotvec[i]
    return c


class SimBatch:
    def __init__(self, model, cfg, n_envs, device=0, dtype=torch.float32):
        if not torch.cuda.is_available():
            raise RuntimeError("ur3e_b200 needs a CUDA device (there is no CPU fallback)")
        self._L = _lib.load()
        self.model = model if isinstance(model, Model) else Model(model)
        self.cfg = cfg
        self.n = int(n_envs)
        self.device = torch.device("cuda", device)
        self.dtype = dtype
        code = {torch.float32: _lib.F32, torch.float64: _lib.F64}[dtype]
        self.ptr = self._L.ur3e_batch_create(self.model.ptr, C.byref(cfg), self.n, device, code)
        if not self.ptr:
            raise RuntimeError("ur3e_batch_create: " + _lib.last_error())
        self.obs_dim, self.act_dim = cfg.obs_dim, cfg.act_dim
        kw = dict(device=self.device)
        self.obs = torch.zeros(self.n, self.obs_dim, dtype=dtype, **kw)
        self.final_obs = torch.zeros(self.n, self.obs_dim, dtype=dtype, **kw)
        self.reward = torch.zeros(self.n, dtype=dtype, **kw)
        self.terminated = torch.zeros(self.n, dtype=torch.uint8, **kw)
        self.truncated = torch.zeros(self.n, dtype=torch.uint8, **kw)
        self._stats = torch.zeros(16, dtype=torch.float64, **kw)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, t, shape, dtype=None):
        dtype = dtype or self.dtype
        if not (t.is_cuda and t.device == self.device and t.dtype == dtype and t.is_contiguous() and tuple(t.shape) == tuple(shape)):
            raise ValueError("expected a contiguous %s tensor of shape %s on %s, got %s %s on %s" % (dtype, tuple(shape), self.device, t.dtype, tuple(t.shape), t.device))
        return C.c_void_p(t.data_ptr())

    def reset(self, seed=0, mask=None):
        mp = None
        if mask is not None:
            mask = mask.to(torch.uint8).contiguous()
            mp = self._chk(mask, (self.n,), torch.uint8)
        _lib.check(self._L.ur3e_batch_reset(self.ptr, mp, seed, C.c_void_p(self.obs.data_ptr()), self._stream()), "ur3e_batch_reset")
        return self.obs

    def step(self, actions, want_final_obs=True):
        """One env-step for every environment (one kernel launch).  Returns views of the batch-owned output tensors."""
        a = self._chk(actions, (self.n, self.act_dim))
        fo = C.c_void_p(self.final_obs.data_ptr()) if want_final_obs else None
        _lib.check(self._L.ur3e_batch_step(self.ptr, a, C.c_void_p(self.obs.data_ptr()), C.c_void_p(self.reward.data_ptr()),
                                           C.c_void_p(self.terminated.data_ptr()), C.c_void_p(self.truncated.data_ptr()), fo, self._stream()), "ur3e_batch_step")
        return self.obs, self.reward, self.terminated, self.truncated

    def step_host(self, actions, obs, reward, terminated, truncated):
        """Host-buffer entry point (numpy arrays or pinned CPU tensors): H2D, step, D2H, synchronise."""
        def p(x):
            return C.c_void_p(x.data_ptr() if isinstance(x, torch.Tensor) else x.ctypes.data)
        _lib.check(self._L.ur3e_batch_step_host(self.ptr, p(actions), p(obs), p(reward), p(terminated), p(truncated)), "ur3e_batch_step_host")

    def get_state(self):
        qpos = torch.empty(self.n, self.model.nq, dtype=self.dtype, device=self.device)
        qvel = torch.empty(self.n, self.model.nv, dtype=self.dtype, device=self.device)
        ws = torch.empty(self.n, self.model.nv, dtype=self.dtype, device=self.device)
        _lib.check(self._L.ur3e_batch_get_state(self.ptr, C.c_void_p(qpos.data_ptr()), C.c_void_p(qvel.data_ptr()), C.c_void_p(ws.data_ptr()), self._stream()), "ur3e_batch_get_state")
        return qpos, qvel, ws

    def set_state(self, qpos, qvel, qacc_warmstart=None):
        """MujocoEnv.set_state: write qpos/qvel for every env and run the forward pass (fresh kinematics cache)."""
        qp = self._chk(qpos, (self.n, self.model.nq)); qv = self._chk(qvel, (self.n, self.model.nv))
        ws = self._chk(qacc_warmstart, (self.n, self.model.nv)) if qacc_warmstart is not None else None
        _lib.check(self._L.ur3e_batch_set_state(self.ptr, qp, qv, ws, self._stream()), "ur3e_batch_set_state")

    def enable_sensors(self, on=True):
        """Attach (or detach) the [n, 21] logging output: after every step `self.sensors` holds, from the step's last mj_step,
        7 x actuatorfrc, the two touch sensors (reference assets/main.xml:392-408) and the tcp site pose; see ur3e_b200/utils.py."""
        self.sensors = torch.zeros(self.n, _lib.NSENSOR, dtype=self.dtype, device=self.device) if on else None
        _lib.check(self._L.ur3e_batch_set_sensor_buffer(self.ptr, C.c_void_p(self.sensors.data_ptr()) if on else None), "ur3e_batch_set_sensor_buffer")
        return self.sensors

    def stats(self, reset=True):
        _lib.check(self._L.ur3e_batch_stats(self.ptr, C.c_void_p(self._stats.data_ptr()), int(reset), self._stream()), "ur3e_batch_stats")
        return self._stats

    def stats_dict(self, reset=True):
        v = self.stats(reset).cpu().numpy()
        return {k: float(v[i]) for i, k in enumerate(_lib.STAT_NAMES)}

    def debug_forward(self, env=0):
        """Forward pass (with constraint solve, zero ctrl) of one environment at its current state; float64 host copies."""
        nv = self.model.nv
        M = np.zeros((nv, nv)); bias = np.zeros(nv); qacc = np.zeros(nv); fc = np.zeros(nv)
        info = np.zeros(8, dtype=np.int32); con = np.zeros((_lib.MAXCON, 4)); cache = np.zeros(_lib.CACHE_SIZE)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        torch.cuda.synchronize(self.device)
        _lib.check(self._L.ur3e_batch_debug_forward(self.ptr, env, dp(M), dp(bias), dp(qacc), dp(fc), info.ctypes.data_as(C.POINTER(C.c_int32)), dp(con), dp(cache)), "ur3e_batch_debug_forward")
        return dict(M=M, qfrc_bias=bias, qacc=qacc, qfrc_constraint=fc, ncon=int(info[0]), nefc=int(info[1]), solver_iter=int(info[2]),
                    overflow=int(info[3]), contacts=con[:int(info[0])], tcp_pos=cache[:3], tcp_mat=cache[3:12].reshape(3, 3), J_arm=cache[12:48].reshape(6, 6), bias_arm=cache[48:54])

    @property
    def launch_count(self):
        return int(self._L.ur3e_batch_launch_count(self.ptr))

    def kernel_info(self):
        v = [C.c_int32() for _ in range(4)]
        _lib.check(self._L.ur3e_batch_kernel_info(self.ptr, *[C.byref(x) for x in v]), "ur3e_batch_kernel_info")
        t = (C.c_int64 * 8)()
        _lib.check(self._L.ur3e_batch_tier_info(self.ptr, t), "ur3e_batch_tier_info")
        d = dict(arena_bytes=v[0].value, warps_per_block=v[1].value, blocks_per_sm=v[2].value, regs_per_thread=v[3].value,
                 state_bytes=int(self._L.ur3e_batch_state_bytes(self.ptr)))
        if t[0]:
            d["lite"] = dict(arena_bytes=int(t[0]), warps_per_block=int(t[1]), blocks_per_sm=int(t[2]), regs_per_thread=int(t[3]),
                             lite_tier_steps=int(t[4]), full_only_steps=int(t[5]), last_overflow_envs=int(t[6]))
            mv = [C.c_int32() for _ in range(3)]
            _lib.check(self._L.ur3e_batch_mid_tier_info(self.ptr, *[C.byref(x) for x in mv]), "ur3e_batch_mid_tier_info")
            if mv[0].value:
                d["grasp_tier"] = dict(arena_bytes=mv[0].value, warps_per_block=mv[1].value, regs_per_thread=mv[2].value)
        return d

    def kernel_timing(self, enable=True):
        """Bracket every step-kernel launch with CUDA events on the launching stream (measurement aid, see kernel_times)."""
        _lib.check(self._L.ur3e_batch_kernel_timing(self.ptr, int(enable)), "ur3e_batch_kernel_timing")

    def kernel_times(self):
        """{tier: (total ms, launches)} of the step kernels since kernel_timing(True); synchronises."""
        v = (C.c_double * 6)()
        _lib.check(self._L.ur3e_batch_kernel_times(self.ptr, v), "ur3e_batch_kernel_times")
        return {"lite": (v[0], int(v[1])), "full": (v[2], int(v[3])), "side": (v[4], int(v[5]))}

    def close(self):
        if getattr(self, "ptr", None):
            self._L.ur3e_batch_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
