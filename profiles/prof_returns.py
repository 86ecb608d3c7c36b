"""ncu target: the two returns-scan passes on a c3-shaped rollout (65,536 envs x 200 steps x 6 columns)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from d2d_ppo_b200.algorithms._nets import returns_emit, returns_stats
dev = torch.device("cuda", 0)
T, N, B = 200, 6, 65536
reward = torch.randint(0, 4, (T, B), dtype=torch.int32, device=dev)
value = torch.randn((T, N, B), device=dev)
one = torch.ones(N, dtype=torch.int32, device=dev)
mean = torch.zeros(N, dtype=torch.float64, device=dev); std = torch.ones(N, dtype=torch.float64, device=dev)
adv = torch.empty((T, N, B), device=dev); ret = torch.empty((T, N, B), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.fill_(1)
    returns_stats(reward, value, 0.4, 0.97, 1)
    flush.fill_(1)
    returns_emit(reward, value, 0.4, 0.97, 1, (mean, std, one), (mean, std, one), adv, ret)
torch.cuda.synchronize()
