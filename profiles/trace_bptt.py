"""Phase timeline of gru_bptt_tc_kernel (CTA 0, thread 0, clock64 at the phase boundaries), from a library built with
-DD2D_BPTT_TRACE into profiles/trace/ (see DESIGN.md section 4); prints the mean cycles of every segment of a step."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from d2d_ppo_b200 import _lib

_lib.LIB_PATH = os.path.join(ROOT, "profiles", "trace", "libd2d_b200.so")
from d2d_ppo_b200 import presets
from d2d_ppo_b200.algorithms.ippo import iPPO
from d2d_ppo_b200.envs import CombinatorialEnv

B, T = 4096, 200
dev = torch.device("cuda", 0)
kw = presets.combinatorial_kwargs("setup_8_channels", load=1 / 3, episode_length=T)
env = CombinatorialEnv(n_envs=B, device=dev, seed=7, **kw)
agent = iPPO(env, hidden_size=64, gamma=0.4, policy_lr=3e-4, value_lr=1e-3, useRNN=True, combinatorial=True,
             history_len=6, early_stopping=False, seed=1, scratch_bytes=24 << 30)
agent.create_rollouts(B)
agent.update_epoch()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (16 * 512))()
lib = ctypes.CDLL(_lib.LIB_PATH)
assert lib.d2d_debug_bptt_trace(buf) == 0
tr = np.frombuffer(buf, dtype=np.int64).reshape(512, 16)
L = 6
names = ["0 top", "1 staged + arrived (a_ready)", "2 recompute done (r_ready)", "3 first chunk maths done",
         "4 w_done seen", "5 phase B done, arrived (g_ready)", "6 next loads issued", "7 d_ready", "8 phase C done"]
NS = len(names)
mid = [i for i in range(40, 500) if (i % L) not in (0, L - 1) and tr[i, NS - 1] > 0]   # steps that have all 14 stamps
d = tr[mid]
print(f"{len(mid)} interior steps; mean cycles between stamps")
for i in range(1, NS):
    print(f"  {names[i - 1]:32s} -> {names[i]:32s} {np.mean(d[:, i] - d[:, i - 1]):8.0f}")
nxt = tr[[i + 1 for i in mid], 0] - d[:, NS - 1]
print(f"  last -> next top {np.mean(nxt):8.0f}")
print(f"  whole step {np.mean(tr[[i + 1 for i in mid], 0] - d[:, 0]):8.0f} cycles")
