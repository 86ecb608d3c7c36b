import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from d2d_ppo_b200 import _lib as L
from d2d_ppo_b200.algorithms._nets import NetSet
dev = torch.device("cuda", 0)
for H, Lh, I, E, T, N in [(64, 6, 30, 256, 9, 2), (64, 6, 30, 300, 9, 3), (16, 4, 11, 40, 7, 1), (32, 3, 30, 1000, 5, 6), (64, 6, 30, 16384, 4, 6)]:
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1)
    a = NetSet(L.NET_GRU, L.OUT_SIGMOID, N, E, [I]*N, [k*I for k in range(N)], N*I, H, 8, Lh, dev, 1e-3, generator=gen, inputs_bf16_exact=True)
    b = NetSet(L.NET_GRU, L.OUT_SIGMOID, N, E, [I]*N, [k*I for k in range(N)], N*I, H, 8, Lh, dev, 1e-3, inputs_bf16_exact=False)
    b.params.copy_(a.params)
    x = torch.zeros((Lh - 1 + T, N * I, E), device=dev)
    x[Lh-1:] = torch.randint(-1, 4, (T, N*I, E), device=dev).float()
    for padded in (0, 1):
        ya = a.forward(x, Lh-1, 0, T, padded=padded); yb = b.forward(x, Lh-1, 0, T, padded=padded)
        torch.cuda.synchronize()
        err = (ya - yb).abs().max().item() / yb.abs().max().item()
        print(f"H={H} L={Lh} I={I} E={E} T={T} N={N} padded={padded}: tc-vs-ffma rel err {err:.2e}", flush=True)
    sa = torch.cat([a.rollout_step(x, Lh-1, t) for t in range(T)]); sb = b.forward(x, Lh-1, 0, T, padded=0)
    print("   rollout_step tc vs ffma forward:", ((sa - sb).abs().max() / sb.abs().max()).item(), flush=True)
# timing at c3 size
N, E, I, H, Lh = 6, 65536, 30, 64, 6
a = NetSet(L.NET_GRU, L.OUT_SIGMOID, N, E, [I]*N, [k*I for k in range(N)], N*I, H, 8, Lh, dev, 1e-3, inputs_bf16_exact=True)
b = NetSet(L.NET_GRU, L.OUT_SIGMOID, N, E, [I]*N, [k*I for k in range(N)], N*I, H, 8, Lh, dev, 1e-3, inputs_bf16_exact=False)
x = torch.randint(-1, 4, (Lh - 1 + 8, N*I, E), device=dev).float()
for name, ns in (("tc", a), ("ffma", b)):
    for _ in range(2): ns.rollout_step(x, Lh-1, 6)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(8): ns.rollout_step(x, Lh-1, t if name == "ffma" else 7)
    e1.record(); torch.cuda.synchronize()
    print(name, "rollout_step ms per step at B=65536:", e0.elapsed_time(e1) / 8)
