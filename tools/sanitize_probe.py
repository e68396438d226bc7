"""Small contact-rich run for compute-sanitizer (racecheck / memcheck): 2 blocks of the lite and full step kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ur3e_b200._lib as lib
from ur3e_b200.envs import UR3eVecEnv
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "env_v0.npz"))
n = 30
for kw in (dict(lite_max_contacts=4), dict(single_tier=1)):
    env = UR3eVecEnv("gymnasium_env/ur3e-v0", n, dtype=torch.float32, auto_reset=True, reset_noise=lib.NOISE_HIGH, max_steps=60, **kw)
    env.reset(seed=1)
    env.set_state(torch.tensor(np.tile(g["qpos0"], (n, 1)), device="cuda", dtype=torch.float32), torch.tensor(np.tile(g["qvel0"], (n, 1)), device="cuda", dtype=torch.float32))
    for k in range(150):
        a = torch.tensor(np.tile(g["actions"][k], (n, 1)), device="cuda", dtype=torch.float32)
        a[n // 4:, 3] = 0.0
        env.step(a)
    torch.cuda.synchronize()
    print("ok", kw, env.episode_stats())
