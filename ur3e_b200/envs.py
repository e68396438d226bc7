"""Vectorised drop-in environments with the reference's Gymnasium ids and spaces.

    env = UR3eVecEnv("gymnasium_env/ur3e-v2", num_envs=65536)      # torch tensors in/out, zero-copy
    obs, info = env.reset(seed=0)
    obs, reward, terminated, truncated, info = env.step(actions)   # actions: cuda tensor [N, act_dim]

replaces N copies of the reference's env classes behind SB3's SubprocVecEnv
(reference gymnasium_src/scripts/regular_rl/rl/train_rl.py:38-44; env classes gymnasium_env/envs/*.py;
registration gymnasium_env/envs/register_envs.py:4-25).  Auto-reset follows the SB3 VecEnv convention: a finished
environment returns its reset observation and the terminal one is available as `final_obs` / infos[i]["terminal_observation"].
gymnasium and stable_baselines3 are optional imports (neither is present in the build container).
"""
import numpy as np
import torch

from . import _lib, presets
from .batch import SimBatch
from .model import Model, asset

ENV_IDS = tuple(presets.ENV_SPECS.keys())


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (used when gymnasium is not installed)."""

    def __init__(self, low, high, shape=None, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy() if shape is not None else np.asarray(low, dtype=dtype)
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy() if shape is not None else np.asarray(high, dtype=dtype)
        self.shape = self.low.shape

    def sample(self, rng=None):
        rng = rng or np.random.default_rng()
        lo = np.where(np.isfinite(self.low), self.low, -1.0); hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


def _box(low, high, shape=None, dtype=np.float64):
    try:
        from gymnasium import spaces
        return spaces.Box(low=low, high=high, shape=shape, dtype=dtype)
    except Exception:
        return Box(low, high, shape, dtype)


class UR3eVecEnv:
    """N environments of one of the reference's four ids on one GPU (torch API)."""

    metadata = {"render_modes": [], "autoreset_mode": "same_step"}

    def __init__(self, env_id="gymnasium_env/ur3e-v2", num_envs=1, device=0, dtype=torch.float32, auto_reset=True, env_id_base=0, **config_overrides):
        if env_id not in presets.ENV_SPECS:
            raise ValueError("unknown env id %r (known: %s)" % (env_id, ", ".join(ENV_IDS)))
        xml, kw, _, _ = presets.ENV_SPECS[env_id]
        self.env_id, self.num_envs = env_id, int(num_envs)
        self.model = Model(asset(xml))
        self.cfg = presets.make_config(self.model, kw, auto_reset=int(bool(auto_reset)), env_id_base=env_id_base, **config_overrides)
        self.batch = SimBatch(self.model, self.cfg, self.num_envs, device, dtype)
        lo, hi = presets.action_bounds(self.model, env_id)
        self.single_action_space = _box(lo, hi, dtype=np.float64)
        self.single_observation_space = _box(-np.inf, np.inf, (self.cfg.obs_dim,), np.float64)     # ur3e_env2.py:44-48
        self.action_space = _box(np.tile(lo, (self.num_envs, 1)), np.tile(hi, (self.num_envs, 1)), dtype=np.float64)
        self.observation_space = _box(-np.inf, np.inf, (self.num_envs, self.cfg.obs_dim), np.float64)
        self.frame_skip = self.cfg.frame_skip
        self.dt = self.model.opt.timestep * self.frame_skip
        self.metadata = dict(self.metadata, render_fps=int(round(1.0 / self.dt)))                  # ur3e_env2.py:19-28
        self.device, self.dtype = self.batch.device, dtype
        self._seed = 0

    # ---- gymnasium.vector-style API (torch tensors)
    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._seed = int(seed)
        return self.batch.reset(self._seed), {}

    def step(self, actions):
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions), dtype=self.dtype, device=self.device)
        if actions.dtype != self.dtype or actions.device != self.device or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=self.dtype).contiguous()
        if tuple(actions.shape) != (self.num_envs, self.cfg.act_dim):
            raise ValueError("Action dimension mismatch: expected %s, got %s" % ((self.num_envs, self.cfg.act_dim), tuple(actions.shape)))   # MujocoEnv.do_simulation
        obs, rew, term, trunc = self.batch.step(actions)
        return obs, rew, term.bool(), trunc.bool(), {"final_obs": self.batch.final_obs}

    def set_state(self, qpos, qvel):
        """MujocoEnv.set_state for every env (qpos [N,nq], qvel [N,nv])."""
        self.batch.set_state(qpos, qvel)

    def get_state(self):
        return self.batch.get_state()

    def episode_stats(self, reset=True):
        return self.batch.stats_dict(reset)

    def close(self):
        self.batch.close()


class SB3VecEnv:
    """stable_baselines3.common.vec_env.VecEnv-shaped adapter (numpy at the boundary, as SB3 expects).

    Provides what `make_vec_env(..., vec_env_cls=SubprocVecEnv)` + `Monitor` give the reference's trainers
    (train_rl.py:38-57): step_async/step_wait, auto-reset with infos[i]["terminal_observation"],
    infos[i]["TimeLimit.truncated"], infos[i]["episode"] = {"r", "l"}, get_attr/set_attr/env_method/seed/close.
    """

    def __init__(self, env_id="gymnasium_env/ur3e-v2", n_envs=1, device=0, dtype=torch.float32, **kw):
        self.venv = UR3eVecEnv(env_id, n_envs, device, dtype, auto_reset=True, **kw)
        self.num_envs = n_envs
        self.observation_space = self.venv.single_observation_space
        self.action_space = self.venv.single_action_space
        self.render_mode = None
        self._actions = None
        self._ret = np.zeros(n_envs); self._len = np.zeros(n_envs, dtype=np.int64)
        self._seed = 0

    def seed(self, seed=None):
        self._seed = 0 if seed is None else int(seed)
        return [self._seed + i for i in range(self.num_envs)]

    def reset(self):
        obs, _ = self.venv.reset(seed=self._seed)
        self._ret[:] = 0; self._len[:] = 0
        return obs.double().cpu().numpy()

    def step_async(self, actions):
        self._actions = np.asarray(actions)

    def step_wait(self):
        obs, rew, term, trunc, info = self.venv.step(self._actions)
        obs_n = obs.double().cpu().numpy(); rew_n = rew.double().cpu().numpy()
        term_n = term.cpu().numpy(); trunc_n = trunc.cpu().numpy(); done = term_n | trunc_n
        self._ret += rew_n; self._len += 1
        infos = [{} for _ in range(self.num_envs)]
        if done.any():
            fo = info["final_obs"].double().cpu().numpy()
            for i in np.nonzero(done)[0]:
                infos[i]["terminal_observation"] = fo[i]
                infos[i]["TimeLimit.truncated"] = bool(trunc_n[i] and not term_n[i])
                infos[i]["episode"] = {"r": float(self._ret[i]), "l": int(self._len[i])}
                self._ret[i] = 0; self._len[i] = 0
        return obs_n, rew_n, done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def get_attr(self, name, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self.venv, name)] * n

    def set_attr(self, name, value, indices=None):
        setattr(self.venv, name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self.venv, method_name)(*args, **kwargs)] * n

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [False] * n

    def close(self):
        self.venv.close()


def register_envs():
    """Register the reference's four ids with gymnasium (vector entry points); no-op without gymnasium."""
    try:
        import gymnasium
    except Exception:
        return False
    for env_id in ENV_IDS:
        def make_vec(num_envs=1, _id=env_id, **kw):
            return UR3eVecEnv(_id, num_envs, **kw)
        try:
            gymnasium.register(id=env_id, vector_entry_point=make_vec)
        except Exception:
            pass
    return True
