"""ncu target: a few learned-policy rollout steps at 65,536 envs (c3: GRU actor + GRU critic, H 64, L 6)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from d2d_ppo_b200 import presets
from d2d_ppo_b200.algorithms.ippo import iPPO
from d2d_ppo_b200.envs import CombinatorialEnv
dev = torch.device("cuda", 0)
B, T = 65536, 12
kw = presets.combinatorial_kwargs("setup_8_channels", load=1 / 3, episode_length=T)
env = CombinatorialEnv(n_envs=B, device=dev, seed=7, **kw)
agent = iPPO(env, hidden_size=64, gamma=0.4, policy_lr=3e-4, value_lr=1e-3, useRNN=True, combinatorial=True,
             history_len=6, early_stopping=False, seed=1, scratch_bytes=6 << 30)
agent.create_rollouts(B)
torch.cuda.synchronize()
torch.cuda.profiler.start()
agent.create_rollouts(B)
torch.cuda.profiler.stop()
torch.cuda.synchronize()
