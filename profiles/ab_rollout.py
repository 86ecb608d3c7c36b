"""A/B of one kernel switch on the learned-policy rollout (bench.py's rollout_learned workload) inside ONE process, so
that both arms see the same box and clocks: python profiles/ab_rollout.py <switch id> [envs] [rounds]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from d2d_ppo_b200 import _lib, presets
from d2d_ppo_b200.algorithms.ippo import iPPO
from d2d_ppo_b200.envs import CombinatorialEnv

which = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
kw = presets.combinatorial_kwargs("setup_8_channels", load=1 / 3)
env = CombinatorialEnv(n_envs=B, device=dev, seed=7, **kw)
agent = iPPO(env, hidden_size=64, gamma=0.4, policy_lr=3e-4, value_lr=1e-3, useRNN=True, combinatorial=True,
             history_len=6, early_stopping=False, seed=1, scratch_bytes=6 << 30)
agent.create_rollouts(B)
steps = B * env.n_agents * kw["episode_length"]
for r in range(rounds):
    for on in (1, 0):
        _lib.set_kernel_switch(which, on)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            agent.create_rollouts(B)
        e1.record()
        torch.cuda.synchronize()
        print(f"round {r} switch {which} = {on}: {2 * steps / (e0.elapsed_time(e1) * 1e-3):.4e} agent-steps/s", flush=True)
_lib.set_kernel_switch(which, 1)
