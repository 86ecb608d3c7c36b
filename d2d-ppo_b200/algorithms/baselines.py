"""Baseline policies on the device.  Only CombinatorialRandomAccess is in scope: it is
the policy of run_ma_baselines.py:71-74 and xp_n_agents.py:137-140; the other baselines of the reference are stale
against its current env return types.  Drop-in for algorithms/baselines.py:171-222.
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
            rew = torch.zeros(B, dtype=torch.float64, device=env.device)
            done = False
            while not done:
                _, _, r, done, _ = env.step_random_access(self.transmission_prob, with_obs=False, with_state=False)
                r = torch.as_tensor(r, device=env.device)
                rew += (r.reshape(B, -1)[:, 0] if r.dim() > 0 else r).to(torch.float64)
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
