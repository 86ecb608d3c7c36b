"""ChannelSelectionEnv: each of N devices picks one of C channels (0 = stay idle).

Drop-in for envs/channel_selection_env.py:4-235 of the reference; the step is ``sel_step_kernel`` in
csrc/env_kernels.cu.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L
from .. import spaces
from ._base import LockstepEnv


class ChannelSelectionEnv(LockstepEnv):
    KIND = L.ENV_CHANNEL_SELECTION

    def __init__(self, n_agents, n_channels, deadlines, lbdas, period=5, arrival_probs=None, offsets=None,
                 episode_length=100, traffic_model="aperiodic", periodic_devices=[], reward_type=0,
                 channel_switch=None, verbose=False,
                 *, n_envs=None, device=None, seed=0, rng="philox", env_offset=0):
        if channel_switch is None:
            # The reference default is zeros(n_agents) (channel_selection_env.py:35-36) although the vector is
            # indexed by channel 0..C (:105); all-zero switch probabilities of the right length mean the same.
            self.channel_switch = np.zeros(n_channels + 1)
        else:
            self.channel_switch = np.asarray(channel_switch, dtype=np.float64)
        if self.channel_switch.shape[0] < n_channels + 1:
            raise IndexError("channel_switch needs n_channels + 1 entries (channel 0 is the idle choice)")
        if n_channels + 1 > L.MAX_CHANNELS:
            raise ValueError(f"n_channels must be <= {L.MAX_CHANNELS - 1}")
        self._setup(n_agents=n_agents, n_channels=n_channels, deadlines=deadlines, lbdas=lbdas, period=period,
                    arrival_probs=arrival_probs, offsets=offsets, episode_length=episode_length,
                    traffic_model=traffic_model, periodic_devices=periodic_devices, reward_type=reward_type,
                    switch_probs=self.channel_switch[: n_channels + 1], verbose=verbose,
                    n_envs=n_envs, device=device, seed=seed, rng=rng, env_offset=env_offset)
        self.action_space = spaces.Tuple([spaces.Discrete(self.n_channels + 1) for _ in range(self.n_agents)])
        self.channel_errors = 0  # never incremented by the reference env (channel_selection_env.py:85)

    def _split_state(self, state):
        s = state[:, 0].cpu().numpy().astype(np.float64)
        sd = int(self.deadlines.sum())
        return [s[:sd], s[sd:]]

    def reset(self, *, with_state=True):
        obs, state = self._reset_device(True, with_state)
        if self.compat:
            return self._compat_obs(obs), self._split_state(state)
        return self._obs_views(obs), (state.t() if state is not None else None)

    def _new_ack(self):
        return torch.empty((self.n_channels + 1, self.n_envs), dtype=torch.float32, device=self.device)

    def step(self, actions, *, packed=False, with_obs=True, with_state=True, out_obs=None, out_state=None):
        if packed:
            a = actions
            assert a.shape == (self.n_agents, self.n_envs) and a.dtype == torch.uint8
        else:
            a = torch.as_tensor(np.asarray(actions) if self.compat else actions)
            if a.device != self.device:
                a = a.to(self.device, non_blocking=True)
            a = a.reshape(self.n_envs, self.n_agents)
            if a.numel() and (int(a.min()) < 0 or int(a.max()) > self.n_channels):
                raise IndexError("channel id out of range 0..n_channels")
            a = a.to(torch.uint8).t().contiguous()
        obs, state, reward, done = self._step_device(a, with_obs, with_state, out_obs, out_state)
        if self.compat:
            rewards = np.array([int(reward[0].item()) for _ in range(self.n_agents)])
            return self._compat_obs(obs), self._split_state(state), rewards, done, {}
        rewards = reward.unsqueeze(1).expand(self.n_envs, self.n_agents)
        return (self._obs_views(obs) if obs is not None else None,
                state.t() if state is not None else None, rewards, done, {})

    def step_random_access(self, *, with_obs=True, with_state=True, out_obs=None, out_state=None, return_actions=False):
        """One step with the fused ``RandomAccess`` policy (algorithms/baselines.py:10-14): every device with a packet
        picks a channel id uniformly from 0..C inside the step kernel (Philox policy stream)."""
        if self.compat:
            with_obs = with_state = True
        acts = torch.empty((self.n_agents, self.n_envs), dtype=torch.uint8, device=self.device) if return_actions else None
        obs, state, reward, done = self._step_device(None, with_obs, with_state, out_obs, out_state,
                                                     random_access_tp=0.0, actions_out=acts)
        if self.compat:
            rewards = np.array([int(reward[0].item()) for _ in range(self.n_agents)])
            out = (self._compat_obs(obs), self._split_state(state), rewards, done, {})
        else:
            rewards = reward.unsqueeze(1).expand(self.n_envs, self.n_agents)
            out = (self._obs_views(obs) if obs is not None else None, state.t() if state is not None else None,
                   rewards, done, {})
        return out + (acts,) if return_actions else out

    @property
    def channel_state(self):
        chan = self._export()[1]
        return self._maybe_squeeze(self._bits(chan, self.n_channels + 1))

    @property
    def selected_channel_qualities(self):
        return self._stat(0)

    @property
    def number_selected_channel(self):
        return self._stat(1)
