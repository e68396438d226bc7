"""How many environments of the ur3e-v2 random-action rollout exceed a given lite-tier contact / row cap (they are re-stepped by
the full tier): informs the choice of the lite size class's caps.  usage: python tools/overflow_probe.py [envs]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ur3e_b200.envs import UR3eVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for caps in (dict(), dict(lite_max_contacts=7), dict(lite_max_contacts=6), dict(lite_max_contacts=5), dict(lite_max_rows=40), dict(lite_max_rows=36)):
    env = UR3eVecEnv("gymnasium_env/ur3e-v2", n, **caps)
    env.reset(seed=0)
    lo = torch.tensor(env.single_action_space.low, device="cuda", dtype=torch.float32); hi = torch.tensor(env.single_action_space.high, device="cuda", dtype=torch.float32)
    g = torch.Generator(device="cuda").manual_seed(1)
    worst = 0; tot = 0
    for k in range(300):
        a = lo + (hi - lo) * torch.rand(n, 4, device="cuda", generator=g)
        env.step(a)
        if k % 10 == 9:
            torch.cuda.synchronize()
            o = env.batch.kernel_info()["lite"]["last_overflow_envs"]; worst = max(worst, o); tot += o
    print(caps, "max overflow envs / step: %d (%.3f %%), mean %.1f" % (worst, 100.0 * worst / n, tot / 30.0))
    env.close()
