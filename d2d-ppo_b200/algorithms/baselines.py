"""Baseline policies on the device.

* ``CombinatorialRandomAccess`` (algorithms/baselines.py:171-222): the policy of run_ma_baselines.py:71-74 and
  xp_n_agents.py:137-140, fused into the combinatorial env's step kernel.
* ``RandomAccess`` (algorithms/baselines.py:5-45): uniform channel pick on ``ChannelSelectionEnv``; the one other
  baseline of the reference that still runs against its envs.
``EarliestDeadlineFirstScheduler`` and ``GFAccess`` (baselines.py:48-168) are not carried over: in the reference
snapshot they unpack ``env.reset()``'s state as a (buffers, channel) pair, which ``D2DEnv`` no longer returns (flat
array, env.py:189-190), and ``GFAccess.run`` reads ``buffer_state`` before assigning it (:150-154).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _dist


class CombinatorialRandomAccess:
    def __init__(self, env, transmission_prob=0.5, transmission_prob_list=None, verbose=False):
        self.env = env
        self.transmission_prob = transmission_prob
        self.transmission_prob_list = np.arange(0, 1, 0.1) if transmission_prob_list is None else transmission_prob_list
        self.verbose = verbose

    def act(self, buffers=None):
        """Bernoulli(tp) per (device, channel) (baselines.py:181-183).  In ``run`` the draw is fused into the env
        step kernel; this host-visible variant exists for API parity."""
        env = self.env
        shape = (env.n_agents, env.n_channels) if env.compat else (env.n_envs, env.n_agents, env.n_channels)
        return np.random.binomial(1, self.transmission_prob, shape)

    def get_best_transmission_probs(self, n_episodes):
        cv = []
        for tp in self.transmission_prob_list:
            self.transmission_prob = tp
            score, _, _, _ = self.run(n_episodes)
            cv.append(np.mean(score))
        return cv

    def run(self, n_episodes):
        """ceil(n_episodes / B) lockstep batches with the fused random-access policy.  Returns
        (1 - sum discarded / sum received, mean Jain, mean channel score, mean per-episode reward sum), where the
        reward sum counts every agent's copy of the shared reward (np.sum(rewards_episode), baselines.py:211)."""
        env = self.env
        B = env.n_envs
        batches = max(1, -(-int(n_episodes) // B))
        tot = torch.zeros(6, dtype=torch.float64, device=env.device)
        for _ in range(batches):
            env.reset(with_state=False)
            # the whole episode in one library call: steps enqueued back to back, rewards summed on the device
            rew = torch.zeros(B, dtype=torch.int32, device=env.device)
            steps = env.run_random_access(self.transmission_prob, env.episode_length, out_reward=rew, accumulate=True)
            assert steps == env.episode_length
            rew = rew.to(torch.float64)
            disc = torch.as_tensor(env.discarded_packets, device=env.device).to(torch.float64).sum()
            recv = torch.as_tensor(env.received_packets, device=env.device).to(torch.float64).sum()
            jains = torch.as_tensor(env.compute_jains(), device=env.device, dtype=torch.float64).sum()
            chsc = torch.as_tensor(env.compute_channel_score(), device=env.device, dtype=torch.float64).sum()
            tot += torch.stack([disc, recv, jains, chsc, rew.sum() * env.n_agents,
                                torch.tensor(float(B), dtype=torch.float64, device=env.device)])
        _dist.all_reduce_sum_(tot)
        n = tot[5].item()
        if self.verbose:
            print(f"Number of received packets: {tot[1].item()}")
            print(f"Channel score: {tot[3].item() / n}")
        return 1 - tot[0].item() / tot[1].item(), tot[2].item() / n, tot[3].item() / n, tot[4].item() / n


class RandomAccess:
    """Uniform random channel pick for every device that has a packet (baselines.py:5-45), on ``ChannelSelectionEnv``.
    ``act`` takes the flattened buffers [..., N * Dmax] the reference passes (``state[0]``) and draws from torch's CUDA
    generator in batched mode (seed with ``torch.manual_seed``), from ``np.random`` in single-env mode as in the
    reference.  ``run`` in batched mode uses the policy fused into ``sel_step_kernel`` (Philox policy stream, one
    library call per episode) unless ``fused=False``."""

    def __init__(self, env, verbose=False, fused=True):
        self.env = env
        self.verbose = verbose
        self.fused = fused        # batched mode: draw the actions inside the step kernel (Philox) instead of torch.randint

    def act(self, buffers):
        env = self.env
        N, C = env.n_agents, env.n_channels
        if env.compat:
            n_packets = np.asarray(buffers).reshape((N, int(env.deadlines.max()))).sum(1)
            actions = np.random.choice(np.arange(0, C + 1), size=N)
            actions[n_packets == 0] = 0
            return actions
        b = torch.as_tensor(buffers, device=env.device)
        n_packets = b.reshape(env.n_envs, N, int(env.deadlines.max())).sum(2)
        actions = torch.randint(0, C + 1, (env.n_envs, N), device=env.device)
        return torch.where(n_packets == 0, torch.zeros_like(actions), actions)

    def run(self, n_episodes):
        """ceil(n_episodes / B) lockstep batches.  Returns (1 - sum discarded / sum received, mean Jain, mean channel
        score, mean per-episode reward sum over agents and steps), as the reference's ``run``."""
        env = self.env
        if len(set(int(d) for d in env.deadlines)) != 1:
            raise ValueError("RandomAccess reshapes the buffers to (n_agents, deadlines.max()): equal deadlines only")
        B = env.n_envs
        sd = int(env.deadlines.sum())
        batches = max(1, -(-int(n_episodes) // B))
        tot = torch.zeros(6, dtype=torch.float64, device=env.device)
        for _ in range(batches):
            _, state = env.reset()
            buffers = state[0] if env.compat else state[:, :sd]
            rew = torch.zeros(B, dtype=torch.float64, device=env.device)
            done = False
            if self.fused and not env.compat:
                # the policy is drawn inside the step kernel and the whole episode is one library call
                acc = torch.zeros(B, dtype=torch.int32, device=env.device)
                steps = env.run_random_access(0.0, env.episode_length, out_reward=acc, accumulate=True)
                assert steps == env.episode_length
                rew = acc.to(torch.float64) * env.n_agents        # np.sum(rewards_episode): every agent's copy counts
                done = True
            while not done:
                _, state, r, done, _ = env.step(self.act(buffers))
                buffers = state[0] if env.compat else state[:, :sd]
                r = torch.as_tensor(np.asarray(r) if env.compat else r, device=env.device).to(torch.float64)
                rew += r.reshape(B, -1).sum(1)                 # np.sum(rewards_episode): every agent's copy counts
            disc = torch.as_tensor(env.discarded_packets, device=env.device).to(torch.float64).sum()
            recv = torch.as_tensor(env.received_packets, device=env.device).to(torch.float64).sum()
            jains = torch.as_tensor(env.compute_jains(), device=env.device, dtype=torch.float64).sum()
            chsc = torch.as_tensor(env.compute_channel_score(), device=env.device, dtype=torch.float64).sum()
            tot += torch.stack([disc, recv, jains, chsc, rew.sum(),
                                torch.tensor(float(B), dtype=torch.float64, device=env.device)])
        _dist.all_reduce_sum_(tot)
        n = tot[5].item()
        if self.verbose:
            print(f"Number of received packets: {tot[1].item()}")
            print(f"Channel score: {tot[3].item() / n}")
        return 1 - tot[0].item() / tot[1].item(), tot[2].item() / n, tot[3].item() / n, tot[4].item() / n
