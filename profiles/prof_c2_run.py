"""ncu target: RandomAccess-style episodes of D2DEnv (c2: N = 4, deadlines 7) at the named 4,096 envs, the whole episode
in one launch (sc_run_lanes_kernel).  usage: python profiles/prof_c2_run.py [B=4096]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from d2d_ppo_b200 import presets
from d2d_ppo_b200.envs import D2DEnv

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
env = D2DEnv(n_envs=B, device=dev, seed=10, **presets.d2d_c2_kwargs())
rew = torch.zeros(B, dtype=torch.int32, device=dev)
for _ in range(3):
    env.reset(with_state=False)
    env.run_random_access(0.2, env.episode_length, out_reward=rew, accumulate=True)
torch.cuda.synchronize()
