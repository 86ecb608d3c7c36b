"""Run under torchrun on >= 2 GPUs: a training iteration sharded over W ranks (B/W envs each, NCCL gradient
all-reduce) equals the same iteration on one GPU with B envs, because env and sampling streams are keyed by the
GLOBAL env index.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d2d_ppo_b200 import presets  # noqa: E402
from d2d_ppo_b200.algorithms.d2d_ppo import D2DPPO  # noqa: E402
from d2d_ppo_b200.algorithms.ippo import iPPO  # noqa: E402
from d2d_ppo_b200.envs import CombinatorialEnv  # noqa: E402


def run(algo, B, offset, dev, epochs=2):
    kw = presets.combinatorial_kwargs("setup_8_channels", load=0.5, episode_length=30)
    env = CombinatorialEnv(n_envs=B, device=dev, seed=21, env_offset=offset, **kw)
    common = dict(hidden_size=32, gamma=0.6, policy_lr=3e-4, value_lr=1e-3, useRNN=True, combinatorial=True,
                  history_len=4, early_stopping=False, seed=5)
    ag = iPPO(env, **common) if algo == "ippo" else D2DPPO(env, **common)
    scores = ag.create_rollouts(B)[6]
    losses = []
    for e in range(epochs):
        losses.append(ag.update_epoch() if algo == "ippo" else ag.update_epoch(cycle=np.roll(np.arange(6), e)))
    return ag, scores, losses


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B = 512
    for algo in ("ippo", "d2dppo"):
        ag, scores, losses = run(algo, B // world, rank * (B // world), dev)
        gathered = [torch.empty_like(scores) for _ in range(world)]
        dist.all_gather(gathered, scores)
        params = ag.policies.params.clone()
        dist.barrier()
        if rank == 0:
            # single-GPU reference of the same global batch, with the process group hidden from the agents
            from d2d_ppo_b200.algorithms import _dist
            saved = _dist.active
            _dist.active = lambda: False
            ref, ref_scores, ref_losses = run(algo, B, 0, dev)
            _dist.active = saved
            assert torch.equal(torch.cat(gathered), ref_scores), "sharded rollouts differ from the single-GPU run"
            d = (params - ref.policies.params).abs().max().item()
            moved = (ref.policies.params - 0).abs().max().item()
            l0 = np.asarray(losses[-1][0], dtype=np.float64)
            l1 = np.asarray(ref_losses[-1][0], dtype=np.float64)
            assert np.allclose(l0, l1, rtol=1e-4, atol=1e-6), (l0, l1)
            assert d < 5e-5, d
            print(f"{algo}: {world} ranks x {B // world} envs == 1 rank x {B} envs "
                  f"(scores bit-equal, max |param diff| {d:.2e}, losses {l0.reshape(-1)[:2]} vs {l1.reshape(-1)[:2]})")
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
