import time, torch, numpy as np, sys
sys.path.insert(0,'/root/repo')
import ur3e_b200._lib as lib
from ur3e_b200.batch import SimBatch, env_config
from ur3e_b200.model import Model, asset
G=[220,220,120,20,20,40,35,15,15,2,2,2]
for n in (4096, 65536):
    m=Model(asset('main.xml'))
    b=SimBatch(m, env_config(ctrl_mode=lib.CTRL_PID_TASK_ENV, obs_kind=lib.OBS_V2, obs_dim=24, act_dim=4, gains=G, frame_skip=2, reset_key=1, term_kind=lib.TERM_V2, reward_kind=lib.REW_V2, max_steps=2500, auto_reset=1, reset_noise=lib.NOISE_HIGH), n)
    print(b.kernel_info())
    o=b.reset(seed=1)
    lo=torch.tensor([0.29799994-0.25,0.13349916-0.25,0,0],device='cuda'); hi=torch.tensor([0.29799994+0.25,0.13349916+0.25,0.5,1],device='cuda')
    acts=[(lo+(hi-lo)*torch.rand(n,4,device='cuda')).contiguous() for _ in range(8)]
    for k in range(20): b.step(acts[k%8])
    torch.cuda.synchronize(); t=time.time()
    K=100
    for k in range(K): b.step(acts[k%8])
    torch.cuda.synchronize(); dt=time.time()-t
    print(n,'envs:',n*K/dt,'env-steps/s', dt/K*1e3,'ms/step', b.stats_dict())
