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
from d2d_ppo_b200.algorithms.irdqn import iRDQN  # noqa: E402
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


def run_irdqn(B, offset, mb, pick, dev, K=4):
    """K iterations of iRDQN on B envs starting at global env index `offset`; pick(ep) -> (start, global env column)
    of the GLOBAL minibatch, of which this run takes the samples whose column it owns."""
    kw = presets.combinatorial_kwargs("setup_8_channels", load=0.5, episode_length=20)
    env = CombinatorialEnv(n_envs=B, device=dev, seed=23, env_offset=offset, **kw)
    ag = iRDQN(env, history_len=4, replay_start_size=1, replay_buffer_size=10 ** 6, gamma=0.6, update_target_frequency=2,
               minibatch_size=mb, learning_rate=1e-3, loss="huber", early_stopping=False, hidden_size=32, seed=6)
    ag.test = lambda *a, **k: (0.0, 0.0)

    def forced(ep):
        start, col = pick(ep)
        own = (col >= offset) & (col < offset + B)
        assert own.sum() == mb
        return start[own], col[own] - offset
    scores, _, _ = ag.train(K, early_stopping=False, forced_samples=forced)
    return ag, torch.tensor(scores, dtype=torch.float64, device=dev)


def check_irdqn(rank, world, dev):
    B, mb, T, L = 64, 32, 20, 4
    Bw = B // world

    def pick(ep):      # every rank draws the same global minibatch: mb / world samples per shard
        rng = np.random.default_rng(100 + ep)
        start = rng.integers(0, (ep + 1) * T - L, mb)
        col = np.concatenate([r * Bw + rng.integers(0, Bw, mb // world) for r in range(world)])
        return start, col
    ag, scores = run_irdqn(Bw, rank * Bw, mb // world, pick, dev)
    gathered = [torch.empty_like(scores) for _ in range(world)]
    dist.all_gather(gathered, scores)
    params = ag.network.params.clone()
    dist.barrier()
    if rank == 0:
        from d2d_ppo_b200.algorithms import _dist
        saved = _dist.active
        _dist.active = lambda: False
        ref, ref_scores = run_irdqn(B, 0, mb, pick, dev)
        _dist.active = saved
        K = ref_scores.numel() // B
        mine = torch.stack([g.view(K, Bw) for g in gathered], dim=1).reshape(-1)     # iteration-major, shard, env
        assert torch.equal(mine, ref_scores), "sharded iRDQN rollouts differ from the single-GPU run"
        d = (params - ref.network.params).abs().max().item()
        l0, l1 = torch.stack(ag.losses).cpu().numpy(), torch.stack(ref.losses).cpu().numpy()
        assert np.allclose(l0, l1, rtol=1e-4, atol=1e-6), (l0, l1)
        assert d < 5e-5, d
        print(f"irdqn: {world} ranks x {Bw} envs == 1 rank x {B} envs (scores bit-equal, max |param diff| {d:.2e}, "
              f"last losses {l0[-1][:2]} vs {l1[-1][:2]})")
    dist.barrier()


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
    check_irdqn(rank, world, dev)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
