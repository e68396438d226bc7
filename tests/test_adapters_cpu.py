"""CPU suite: the trainer-facing adapters against stand-ins for their consumers' base classes.

stable_baselines3 and gymnasium are not installable here, so minimal ABCs with the same names are injected into sys.modules: the
adapters must subclass them (isinstance is what SB3's `_wrap_env` / `VecNormalize` and gymnasium's `make_vec` check), and the
VecEnv contract (auto-reset infos, per-index attributes, Monitor's episode record) is exercised on a CPU stand-in for the batch."""
import abc
import importlib
import sys
import types

import numpy as np
import pytest
import torch


class _FakeBatchEnv:
    """UR3eVecEnv's torch API on the CPU: obs = [t, env id], reward 1, truncation every `horizon` steps, auto-reset."""

    def __init__(self, n, horizon=3):
        from ur3e_b200.envs import Box
        self.num_envs, self.horizon = n, horizon
        self.single_observation_space = Box(-np.inf, np.inf, (2,)); self.single_action_space = Box(-1.0, 1.0, (4,))
        self.observation_space = Box(-np.inf, np.inf, (n, 2)); self.action_space = Box(-1.0, 1.0, (n, 4))
        self.metadata = {"render_fps": 500}; self.device = torch.device("cpu"); self.render_mode = None
        self.t = torch.zeros(n)
        self.some_attr = "shared"

    def _obs(self):
        return torch.stack([self.t, torch.arange(self.num_envs, dtype=torch.float32)], 1)

    def reset(self, *, seed=None, options=None):
        self.t.zero_(); return self._obs(), {}

    def step(self, actions):
        a = torch.as_tensor(np.asarray(actions), dtype=torch.float32) if not isinstance(actions, torch.Tensor) else actions
        assert tuple(a.shape) == (self.num_envs, 4)
        self.t += 1 + torch.arange(self.num_envs) % 2            # odd environments age twice as fast
        trunc = self.t >= self.horizon
        final = self._obs().clone()
        self.t[trunc] = 0
        return self._obs(), torch.ones(self.num_envs), torch.zeros(self.num_envs, dtype=torch.bool), trunc, {"final_obs": final}

    def hello(self, x):
        return x * 2

    def close(self):
        self.closed = True


@pytest.fixture
def fake_consumers():
    """Inject minimal `stable_baselines3.common.vec_env.VecEnv`, `gymnasium.Env`, `gymnasium.vector.VectorEnv` and reload the adapters."""
    class VecEnv(abc.ABC):
        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space
            self.sb3_base_initialised = True

        @abc.abstractmethod
        def reset(self): ...
        @abc.abstractmethod
        def step_async(self, actions): ...
        @abc.abstractmethod
        def step_wait(self): ...
        @abc.abstractmethod
        def close(self): ...
        @abc.abstractmethod
        def get_attr(self, attr_name, indices=None): ...
        @abc.abstractmethod
        def set_attr(self, attr_name, value, indices=None): ...
        @abc.abstractmethod
        def env_method(self, method_name, *method_args, indices=None, **method_kwargs): ...
        @abc.abstractmethod
        def env_is_wrapped(self, wrapper_class, indices=None): ...

    class Env: ...
    class VectorEnv: ...
    registered = {}
    mods = {"stable_baselines3": types.ModuleType("stable_baselines3"), "stable_baselines3.common": types.ModuleType("stable_baselines3.common"),
            "stable_baselines3.common.vec_env": types.ModuleType("stable_baselines3.common.vec_env"),
            "gymnasium": types.ModuleType("gymnasium"), "gymnasium.vector": types.ModuleType("gymnasium.vector")}
    mods["stable_baselines3.common.vec_env"].VecEnv = VecEnv
    mods["gymnasium"].Env = Env; mods["gymnasium.vector"].VectorEnv = VectorEnv; mods["gymnasium"].vector = mods["gymnasium.vector"]
    mods["gymnasium"].register = lambda id, entry_point=None, vector_entry_point=None: registered.__setitem__(id, (entry_point, vector_entry_point))
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    import ur3e_b200.envs as E
    E = importlib.reload(E)
    yield E, VecEnv, Env, VectorEnv, registered
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    importlib.reload(E)


def test_adapters_subclass_their_consumers_base_classes(fake_consumers):
    E, VecEnv, Env, VectorEnv, registered = fake_consumers
    assert issubclass(E.SB3VecEnv, VecEnv) and issubclass(E.UR3eVecEnv, VectorEnv) and issubclass(E.UR3eEnv, Env)
    sb = E.SB3VecEnv(venv=_FakeBatchEnv(4))
    assert isinstance(sb, VecEnv) and sb.sb3_base_initialised and sb.num_envs == 4            # SB3's _wrap_env would not re-wrap it
    assert E.register_envs() is True
    assert set(registered) == set(E.ENV_IDS)                                                  # register_envs.py:4-25: the same four ids
    assert all(callable(ep) and callable(vep) for ep, vep in registered.values())             # gym.make(id) and gym.make_vec(id) both resolve


def test_sb3_vecenv_contract_on_cpu_stand_in(fake_consumers):
    E = fake_consumers[0]
    sb = E.SB3VecEnv(venv=_FakeBatchEnv(4, horizon=3))
    obs = sb.reset()
    assert obs.shape == (4, 2) and obs.dtype == np.float64
    seen = {}
    for k in range(4):
        sb.step_async(np.zeros((4, 4)))
        obs, rew, done, infos = sb.step_wait()
        assert rew.dtype == np.float64 and done.dtype == bool and len(infos) == 4
        for i in np.nonzero(done)[0]:
            ep = infos[i]["episode"]
            assert set(ep) == {"r", "l", "t"} and ep["r"] == ep["l"] and ep["t"] >= 0         # Monitor's record, reward 1 per step
            assert infos[i]["TimeLimit.truncated"] is True and infos[i]["terminal_observation"][0] >= 3 and obs[i, 0] == 0
            seen.setdefault(i, ep["l"])
        assert all(not infos[i] for i in np.nonzero(~done)[0])
    assert seen == {1: 2, 3: 2, 0: 3, 2: 3}                                                   # odd environments finish after 2 steps, even after 3
    # per-index attributes and methods
    assert sb.get_attr("some_attr") == ["shared"] * 4
    sb.set_attr("some_attr", "mine", indices=[1, 3])
    assert sb.get_attr("some_attr") == ["shared", "mine", "shared", "mine"] and sb.get_attr("some_attr", 1) == ["mine"]
    assert sb.env_method("hello", 21, indices=[0, 2]) == [42, 42] and sb.env_is_wrapped(object) == [False] * 4
    assert sb.seed(7) == [7, 8, 9, 10]
    sb.close(); assert sb.venv.closed


def test_device_side_vecnormalize_matches_sb3_formulas(tmp_path):
    """VecNormalizeGPU against a numpy restatement of SB3's RunningMeanStd.update_from_moments / VecNormalize.step."""
    from ur3e_b200.envs import VecNormalizeGPU
    env = _FakeBatchEnv(8, horizon=5)
    vn = VecNormalizeGPU(env, norm_obs=True, norm_reward=True, clip_obs=5.0, clip_reward=3.0, gamma=0.9)
    mean, var, count = np.zeros(2), np.ones(2), 1e-4
    rmean, rvar, rcount, ret = 0.0, 1.0, 1e-4, np.zeros(8)

    def upd(mean, var, count, x):
        bm, bv, bc = x.mean(0), x.var(0), x.shape[0]
        d = bm - mean; tot = count + bc
        return mean + d * bc / tot, (var * count + bv * bc + d * d * count * bc / tot) / tot, tot

    o, _ = vn.reset()
    mean, var, count = upd(mean, var, count, vn.get_original_obs().double().numpy())
    for k in range(12):
        o, r, te, tr, info = vn.step(torch.zeros(8, 4))
        raw_o, raw_r = vn.get_original_obs().double().numpy(), vn.get_original_reward().double().numpy()
        mean, var, count = upd(mean, var, count, raw_o)
        ret = ret * 0.9 + raw_r
        rmean, rvar, rcount = upd(rmean, rvar, rcount, ret)
        assert np.allclose(o.double().numpy(), np.clip((raw_o - mean) / np.sqrt(var + 1e-8), -5, 5), atol=1e-6)
        assert np.allclose(r.double().numpy(), np.clip(raw_r / np.sqrt(rvar + 1e-8), -3, 3), atol=1e-6)
        ret[(te | tr).numpy()] = 0
    path = str(tmp_path / "vecnormalize.pkl")
    vn.save(path)
    vn2 = VecNormalizeGPU.load(path, _FakeBatchEnv(8, horizon=5))
    assert torch.equal(vn2.obs_rms.mean, vn.obs_rms.mean) and torch.equal(vn2.ret_rms.var, vn.ret_rms.var) and vn2.clip_obs == 5.0
    vn2.training = False
    before = vn2.obs_rms.mean.clone(); vn2.reset(); vn2.step(torch.zeros(8, 4))
    assert torch.equal(before, vn2.obs_rms.mean)                                               # evaluation mode freezes the statistics
    assert vn2.hello(4) == 8                                                                   # everything else is forwarded to the wrapped env
