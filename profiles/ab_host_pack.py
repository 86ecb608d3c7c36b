"""A/B inside one process (same box, same PCIe link): the host-buffer step of bench.py's e2e leg with the reference
action layout, for several host thread counts and with host packing switched off.
usage: python profiles/ab_host_pack.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from d2d_ppo_b200 import _lib as L, presets
from d2d_ppo_b200.envs import CombinatorialEnv

dev = torch.device("cuda", 0)
B, N, Cn, T = 1048576, 6, 8, 200
kw = presets.combinatorial_kwargs("setup_8_channels", load=1 / 3)
env = CombinatorialEnv(n_envs=B, device=dev, seed=7, **kw)
acts = [torch.from_numpy(np.random.default_rng(i).binomial(1, 0.2, (B, N, Cn)).astype(np.uint8)).pin_memory()
        for i in range(4)]
rew = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
obs = torch.empty((env.obs_layout[0], B), dtype=torch.float32, device=dev)


def run(n):
    pending = None
    for i in range(n):
        if env.timestep >= T:
            env.reset(with_state=False)
        tk = env.step_host(acts[i % 4], rew[i % 2], layout="reference", with_state=False, out_obs=obs)
        if pending is not None:
            env.host_wait(pending)
        pending = tk
    env.host_wait(pending)


env.reset(with_state=False)
for rnd in range(2):
    for threads, on in ((16, True), (12, True), (8, True), (12, False)):
        L.check(L.lib().d2d_set_host_threads(threads))
        L.set_kernel_switch(L.SWITCH_HOST_PACK, on)
        run(8)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(150)
        dt = time.perf_counter() - t0
        print(f"round {rnd}: {'host pack, ' + str(threads) + ' threads' if on and threads >= 12 else 'unpacked copy (' + str(threads) + ' threads)':28s} "
              f"{B * N * 150 / dt:.3e} agent-steps/s, {dt / 150 * 1e3:.3f} ms per step", flush=True)
L.set_kernel_switch(L.SWITCH_HOST_PACK, True)
