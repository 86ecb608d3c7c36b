import sys
sys.path.insert(0, "/root/repo")
import torch
from d2d_ppo_b200 import _lib as L
from d2d_ppo_b200.algorithms._nets import NetSet
dev = torch.device("cuda", 0)
N, E, I, H, Lh = 6, 65536, 30, 64, 6
a = NetSet(L.NET_GRU, L.OUT_SIGMOID, N, E, [I]*N, [k*I for k in range(N)], N*I, H, 8, Lh, dev, 1e-3, inputs_bf16_exact=True)
x = torch.randint(-1, 4, (Lh - 1 + 8, N*I, E), device=dev).float()
for _ in range(3): a.rollout_step(x, Lh-1, 7)
torch.cuda.synchronize()
