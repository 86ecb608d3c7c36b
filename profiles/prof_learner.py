"""Small fixed workload for ncu: one c3-shaped iPPO rollout (T steps) and one update epoch.
usage: python profiles/prof_learner.py [B] [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from d2d_ppo_b200 import presets
from d2d_ppo_b200.algorithms.ippo import iPPO
from d2d_ppo_b200.envs import CombinatorialEnv

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda", 0)
kw = presets.combinatorial_kwargs("setup_8_channels", load=1 / 3, episode_length=T)
env = CombinatorialEnv(n_envs=B, device=dev, seed=7, **kw)
agent = iPPO(env, hidden_size=64, gamma=0.4, policy_lr=3e-4, value_lr=1e-3, useRNN=True, combinatorial=True,
             history_len=6, early_stopping=False, seed=1, scratch_bytes=int(__import__("os").environ.get("SCRATCH_GB", "6")) << 30)
agent.create_rollouts(B)
agent.update_epoch()          # warm-up: scratch allocation happens here
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
agent.create_rollouts(B)
e1.record()
torch.cuda.profiler.start()       # ncu --profile-from-start off: capture the update epoch only
agent.update_epoch()
torch.cuda.profiler.stop()
e2.record()
torch.cuda.synchronize()
print(f"B={B} T={T} rollout {e0.elapsed_time(e1):.2f} ms  epoch {e1.elapsed_time(e2):.2f} ms")
