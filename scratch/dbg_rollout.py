import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from _helpers import load_ppo_case, params_from, rel_err
from d2d_ppo_b200 import _lib as L
from d2d_ppo_b200.algorithms._nets import NetSet
dev = torch.device("cuda", 0)
for H, Lh, I, E, T in [(16, 4, 11, 4, 12), (64, 6, 30, 4, 20), (64, 4, 30, 4, 12), (16, 6, 30, 4, 20), (32, 6, 30, 8, 9), (64, 6, 30, 128, 9), (64,2,30,4,5)]:
    N = 2
    torch.manual_seed(0)
    ns = NetSet(L.NET_GRU, L.OUT_SIGMOID, N, E, [I]*N, [k*I for k in range(N)], N*I, H, 8, Lh, dev, 1e-3)
    x = torch.zeros((Lh - 1 + T, N * I, E), device=dev)
    x[Lh-1:] = torch.randint(-1, 3, (T, N*I, E), device=dev).float()
    full = ns.forward(x, Lh-1, 0, T, padded=0)
    one = torch.cat([ns.forward(x, Lh-1, t, t+1, padded=0) for t in range(T)])
    step = torch.cat([ns.rollout_step(x, Lh-1, t) for t in range(T)])
    e1 = [(full[t]-one[t]).abs().max().item() for t in range(T)]
    e2 = [(full[t]-step[t]).abs().max().item() for t in range(T)]
    print(H, Lh, I, E, T, "full-vs-per-t", ["%.1e"%e for e in e1])
    print("           full-vs-step ", ["%.1e"%e for e in e2])
