"""Baseline policies on the device.

* ``CombinatorialRandomAccess`` (algorithms/baselines.py:171-222): the policy of run_ma_baselines.py:71-74 and
  xp_n_agents.py:137-140, fused into the combinatorial env's step kernel.
* ``RandomAccess`` (algorithms/baselines.py:5-45): uniform channel pick on ``ChannelSelectionEnv``; the one other
  baseline of the reference that still runs against its envs.
* ``EarliestDeadlineFirstScheduler`` / ``GFAccess`` (algorithms/baselines.py:48-168) on ``D2DEnv``.  In the reference
  snapshot their ``run`` loops no longer work (they unpack the env's state as a (buffers, channel) pair, which
  ``D2DEnv`` stopped returning: flat array, env.py:98-99; ``GFAccess.run`` reads ``buffer_state`` before assigning it,
  :153) while their ``act`` methods are intact.  Here ``run`` is the loop those lines evidently intend with
  ``buffer_state = env.current_buffers``: EDF's ``act`` is a kernel on the env's buffer records
  (``d2d_env_policy_edf``), GFAccess's ``act`` -- Bernoulli(tp) per device, zero for empty buffers -- is the random
  access draw fused into ``sc_step_kernel`` (the env masks attempts by has-a-packet itself, env.py:126).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L
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


def _episode_totals(env, rew_sum):
    """[discarded, received, sum Jain, sum channel errors, sum of per-episode reward sums, episodes] of one lockstep
    batch (baselines.py:101-111); the reward vector is (N,) copies of the ack, so np.sum counts it N times."""
    dev = env.device
    f64 = torch.float64
    ce = env.channel_errors
    ce = ce.to(f64).sum() if torch.is_tensor(ce) else torch.tensor(float(ce), dtype=f64, device=dev)
    return torch.stack([torch.as_tensor(env.discarded_packets, device=dev).to(f64).sum(),
                        torch.as_tensor(env.received_packets, device=dev).to(f64).sum(),
                        torch.as_tensor(env.compute_jains(), device=dev, dtype=f64).sum(), ce,
                        rew_sum.to(f64).sum() * env.n_agents,
                        torch.tensor(float(env.n_envs), dtype=f64, device=dev)])


def _finish_run(tot, verbose):
    _dist.all_reduce_sum_(tot)
    disc, recv, jains, errors, rew, n = tot.tolist()
    if verbose:
        print(f"Number of received packets: {recv}")
        print(f"Number of channel_losses: {errors}")
    return (1 - disc / recv) if recv else float("nan"), jains / n, int(errors), rew / n   # no packet: numpy's 0 / 0


class EarliestDeadlineFirstScheduler:
    """Centralised earliest-deadline-first grant of D2DEnv's shared channel (baselines.py:48-111)."""

    def __init__(self, env, use_channel=False, verbose=False):
        if env.KIND != L.ENV_SINGLE_CHANNEL:
            raise ValueError("EarliestDeadlineFirstScheduler grants one shared channel: D2DEnv only")
        self.env = env
        self.use_channel = use_channel
        self.verbose = verbose
        self.name = "EDF"

    def preprocess_state(self, state):
        """Slot index of the oldest packet per device, -1 for an empty buffer (baselines.py:55-63).  Host helper on a
        [N, D] array, kept for API parity; ``run`` uses the device kernel."""
        state = np.asarray(state)
        nz = state != 0
        return np.where(nz.any(1), nz.argmax(1), -1)

    def act(self, buffers=None):
        """One-hot grant over the devices.  With ``buffers`` = a host [N, D] array: the reference's host function
        (single-env API parity).  With ``buffers`` = None: the kernel on the env's CURRENT buffers of all B envs
        (``use_channel`` skips devices whose channel is bad) -> device tensor u8 [N, B]."""
        env = self.env
        if buffers is not None:
            agg = self.preprocess_state(buffers)
            holders = np.flatnonzero(agg >= 0)
            pick = holders[agg[holders].argmin()] if holders.size else np.random.randint(env.n_agents)
            actions = np.zeros(env.n_agents)
            actions[pick] = 1.0
            return actions
        out = torch.empty((env.n_agents, env.n_envs), dtype=torch.uint8, device=env.device)
        with torch.cuda.device(env.device):
            L.check(L.lib().d2d_env_policy_edf(env._h, int(bool(self.use_channel)), L.ptr(out), L.current_stream()))
        return out

    def run(self, n_episodes):
        """ceil(n_episodes / B) lockstep batches.  Returns (1 - sum discarded / sum received, mean Jain, total channel
        errors, mean per-episode reward sum) as baselines.py:111."""
        env = self.env
        B, T = env.n_envs, env.episode_length
        tot = torch.zeros(6, dtype=torch.float64, device=env.device)
        rew = torch.empty((T, B), dtype=torch.int32, device=env.device)
        for _ in range(max(1, -(-int(n_episodes) // B))):
            env._reset_device(False, False)
            for t in range(T):
                env._step_device(self.act(), False, False, out_reward=rew[t])
            tot += _episode_totals(env, rew.sum(0))
        return _finish_run(tot, self.verbose)


class GFAccess:
    """Grant-free access: every device with a packet transmits with probability ``transmission_prob``
    (baselines.py:113-168), on D2DEnv."""

    def __init__(self, env, transmission_prob=0.5,
                 transmission_prob_list=[0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1], use_channel=False,
                 verbose=False):
        if env.KIND != L.ENV_SINGLE_CHANNEL:
            raise ValueError("GFAccess draws one binary action per device: D2DEnv only")
        if use_channel:
            raise ValueError("GFAccess(use_channel=True) is not supported: the reference's branch reads an undefined "
                             "variable (baselines.py:148-151) and the fused draw does not see the channel state")
        self.env = env
        self.transmission_prob = transmission_prob
        self.transmission_prob_list = transmission_prob_list
        self.use_channel = use_channel
        self.verbose = verbose

    def act(self, buffers):
        """Host function of the reference on a [N, D] buffers array (baselines.py:121-125), for API parity; ``run``
        draws inside the step kernel."""
        n_packets = np.asarray(buffers).sum(1)
        actions = np.random.binomial(1, p=self.transmission_prob, size=self.env.n_agents)
        actions[n_packets == 0] = 0
        return actions

    def get_best_transmission_probs(self, n_episodes):
        cv = []
        for tp in self.transmission_prob_list:
            self.transmission_prob = tp
            score, _, _, _ = self.run(n_episodes)
            cv.append(np.mean(score))
        return cv

    def run(self, n_episodes):
        env = self.env
        B = env.n_envs
        tot = torch.zeros(6, dtype=torch.float64, device=env.device)
        for _ in range(max(1, -(-int(n_episodes) // B))):
            env._reset_device(False, False)
            rew = torch.zeros(B, dtype=torch.int32, device=env.device)
            steps = env.run_random_access(self.transmission_prob, env.episode_length, out_reward=rew, accumulate=True)
            assert steps == env.episode_length
            tot += _episode_totals(env, rew)
        return _finish_run(tot, self.verbose)
