"""Host-side action packing alone (csrc/host_pack.cpp): u8 [B, N, C] -> bitmasks [N, B] at the bench's e2e shape, cycling
through 4 input buffers (200 MB, larger than the host's caches) as bench.py does.  No GPU involved.
usage: python profiles/time_host_pack.py"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import d2d_ppo_b200  # noqa: F401
from d2d_ppo_b200 import _lib as L

lib = L.lib()
B, N, Cn = 1048576, 6, 8
bufs = [(np.random.default_rng(i).random((B, N, Cn)) < 0.2).astype(np.uint8) for i in range(4)]
out = np.zeros((N, B), np.uint8)
print("cpus", len(os.sched_getaffinity(0)), "default threads", lib.d2d_get_host_threads())
for th in (1, 2, 4, 8, 12, 16):
    lib.d2d_set_host_threads(th)
    for a in bufs:
        lib.d2d_pack_actions_host(C.c_void_p(a.ctypes.data), C.c_void_p(out.ctypes.data), B, N, Cn)
    ts = []
    for i in range(40):
        a = bufs[i % 4]
        t0 = time.perf_counter()
        lib.d2d_pack_actions_host(C.c_void_p(a.ctypes.data), C.c_void_p(out.ctypes.data), B, N, Cn)
        ts.append(time.perf_counter() - t0)
    med = sorted(ts)[len(ts) // 2]
    print(f"{th:2d} threads: median {med * 1e3:.3f} ms = {bufs[0].nbytes / med / 1e9:.1f} GB/s of action bytes "
          f"(min {min(ts) * 1e3:.3f} ms)", flush=True)
b = np.empty_like(bufs[0])
ts = []
for i in range(8):
    t0 = time.perf_counter()
    np.copyto(b, bufs[i % 4])
    ts.append(time.perf_counter() - t0)
print(f"np.copyto of one buffer (1 thread): {min(ts) * 1e3:.2f} ms = {bufs[0].nbytes / min(ts) / 1e9:.1f} GB/s")
