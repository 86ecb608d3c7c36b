"""Determinism stress: the same policy_grad call repeated; report tensors whose gradient moves by > 1e-5 (norm-wise)."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from test_learner_gpu import _netset, _env_minor
from d2d_ppo_b200 import _lib as L
from d2d_ppo_b200.algorithms._nets import action_dtype, policy_head
dev = torch.device("cuda", 0)
H, Lh, E, T, I, C, N = 64, 6, int(os.environ.get("E", 300)), 5, 30, 8, 3
gen = torch.Generator().manual_seed(1)
obs = torch.randint(-1, 4, (E * T, N, I), generator=gen).float()
acts = torch.randint(0, 2, (E * T, N, C), generator=gen).float()
adv = torch.randn(E * T, N, generator=gen)
lead = Lh - 1
x = _env_minor(obs.reshape(E * T, N * I).numpy(), E, T, lead, dev)
in_dim, in_off = [I] * N, [k * I for k in range(N)]
pol = _netset(dev, "gru", "sigmoid", N, E, in_dim, in_off, N * I, H, C, Lh, exact=True)
packed = (acts.long() * (1 << torch.arange(C))).sum(-1)
actions = packed.reshape(E, T, N).permute(1, 2, 0).contiguous().to(action_dtype(0, C)).to(dev)
em = lambda a: a.reshape(E, T, N).permute(1, 2, 0).contiguous().to(dev)
logits = pol.forward(x, lead, 0, T, padded=1)
logp = torch.empty((T, N, E), device=dev)
policy_head(logits, N, E, C, L.OUT_SIGMOID, L.DIST_BERNOULLI, L.ACT_GIVEN, actions, logp)
logp_old = logp + 0.05 * em(torch.randn(E * T, N, generator=gen))
advd = em(adv)
ref = None
bad = 0
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
for rep in range(reps):
    sums = torch.zeros((N, 2), dtype=torch.float64, device=dev)
    pol.zero_grad()
    pol.policy_grad(x, lead, 0, T, L.DIST_BERNOULLI, actions, logp_old, advd, 1, None, 1.0 / (E * T), 0.1, 0.01, sums)
    g = pol.grads.clone()
    if ref is None:
        ref = g
        continue
    rows = []
    for i in range(N):
        for name in pol.keys:
            a, b = pol.tensor_view(g, i, name), pol.tensor_view(ref, i, name)
            d = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)
            if d > 1e-5:
                rows.append((i, name, f"{d:.1e}"))
    if rows:
        bad += 1
        print(rep, rows)
print("bad", bad, "of", reps - 1, {k: v for k, v in os.environ.items() if k.startswith("D2D_")})
