"""Repeat the multi-tile tcgen05 training-path check; print the worst gradient error per repetition."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
src = open("/root/repo/tests/test_learner_gpu.py").read()
src = src.replace('assert rel_err(pol.tensor_view(pol.grads, i, name), gr) < 2e-5, ("policy", i, name)',
                  'REC.append(("pol", i, name, rel_err(pol.tensor_view(pol.grads, i, name), gr)))')
src = src.replace('assert rel_err(val.tensor_view(val.grads, i, name), gr) < 2e-5, ("value", i, name)',
                  'REC.append(("val", i, name, rel_err(val.tensor_view(val.grads, i, name), gr)))')
ns = {"__name__": "m", "REC": []}
exec(compile(src, "t", "exec"), ns)
dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
junk = torch.empty(1 << 28, device=dev).normal_()      # dirty the allocator's memory
del junk
bad = 0
for rep in range(reps):
    ns["REC"].clear()
    ns["test_tensor_core_training_path_vs_autograd"](64, 6, 300, 5, 30, 8, dev)
    worst = max(ns["REC"], key=lambda r: r[3])
    if worst[3] > 2e-5:
        bad += 1
        print(rep, "BAD", [(r[0], r[1], r[2], f"{r[3]:.2e}") for r in ns["REC"] if r[3] > 2e-5][:6])
    junk = torch.empty(1 << 27, device=dev).uniform_(-1e3, 1e3); del junk
print("env", {k: v for k, v in os.environ.items() if k.startswith("D2D_")}, "bad", bad, "of", reps)
