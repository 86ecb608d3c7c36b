"""Print the gradient errors of the ippo c3_gru fixture per tensor (max over agents), for the current library."""
import os, sys, re
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
src = open("/root/repo/tests/test_learner_gpu.py").read()
# turn the per-tensor gradient asserts into recordings
src = src.replace("assert rel_err(mine, gr) < 2e-5, (i, name)", "REC.setdefault('pol/' + name, []).append(rel_err(mine, gr))")
src = src.replace("assert rel_err(val.tensor_view(val.grads, i, name), gr) < 2e-5, (i, name)",
                  "REC.setdefault('val/' + name, []).append(rel_err(val.tensor_view(val.grads, i, name), gr))")
ns = {"__name__": "gradmod", "REC": {}}
exec(compile(src, "test_learner_gpu_rec", "exec"), ns)
dev = torch.device("cuda", 0)
for tag in ("c3_gru", "small_gru"):
    ns["REC"].clear()
    ns["test_ippo_gradients_and_adam"].__wrapped__(tag, 1 << 20, True, True, dev) if hasattr(ns["test_ippo_gradients_and_adam"], "__wrapped__") else ns["test_ippo_gradients_and_adam"](tag, 1 << 20, True, True, dev)
    print(tag, "TC disabled" if os.environ.get("D2D_DISABLE_TCGEN05") == "1" else "TC enabled")
    for k, v in ns["REC"].items():
        print(f"   {k:28s} max {max(v):.2e}  mean {sum(v)/len(v):.2e}")
