"""D2DEnv: N devices, one shared channel, binary transmit action, neighbourhood observations.

Drop-in for envs/env.py:4-233 of the reference; the step is ``sc_step_kernel`` in csrc/env_kernels.cu.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L
from .. import spaces
from ._base import LockstepEnv


class D2DEnv(LockstepEnv):
    KIND = L.ENV_SINGLE_CHANNEL

    def __init__(self, n_agents, deadlines, lbdas, period=5, arrival_probs=None, offsets=None, episode_length=100,
                 traffic_model="aperiodic", periodic_devices=[], reward_type=0, channel_switch=0.2,
                 channel_decoding=0.8, neighbourhoods=None, verbose=False,
                 *, n_envs=None, device=None, seed=0, rng="philox", env_offset=0):
        self.channel_switch = channel_switch
        self.channel_decoding = channel_decoding
        if neighbourhoods is None:  # env.py:38-41
            self.neighbourhoods = [[k] for k in range(n_agents)]
        else:
            self.neighbourhoods = [list(nb) for nb in neighbourhoods]
        self._setup(n_agents=n_agents, n_channels=1, deadlines=deadlines, lbdas=lbdas, period=period,
                    arrival_probs=arrival_probs, offsets=offsets, episode_length=episode_length,
                    traffic_model=traffic_model, periodic_devices=periodic_devices, reward_type=reward_type,
                    switch_probs=np.full(n_agents, float(channel_switch)),
                    neighbourhoods=None if neighbourhoods is None else self.neighbourhoods, verbose=verbose,
                    n_envs=n_envs, device=device, seed=seed, rng=rng, env_offset=env_offset)
        self.action_space = spaces.Tuple([spaces.Discrete(2) for _ in range(self.n_agents)])

    def reset(self, *, with_state=True):
        obs, state = self._reset_device(True, with_state)
        if self.compat:
            return self._compat_obs(obs), state[:, 0].cpu().numpy().astype(np.float64)
        return self._obs_views(obs), (state.t() if state is not None else None)

    def _device_actions(self, actions, packed):
        if packed:
            assert actions.shape == (self.n_agents, self.n_envs) and actions.dtype == torch.uint8
            return actions
        a = torch.as_tensor(np.asarray(actions) if self.compat else actions)
        if a.device != self.device:
            a = a.to(self.device, non_blocking=True)
        a = a.reshape(self.n_envs, self.n_agents)   # (N,) per env, not (N, 1): env.py:126
        return (a != 0).to(torch.uint8).t().contiguous()

    def step(self, actions, *, packed=False, with_obs=True, with_state=True, out_obs=None, out_state=None):
        obs, state, reward, done = self._step_device(self._device_actions(actions, packed), with_obs, with_state,
                                                     out_obs, out_state)
        return self._finish(obs, state, reward, done)

    def step_random_access(self, transmission_prob, *, with_obs=True, with_state=True, out_obs=None, out_state=None):
        obs, state, reward, done = self._step_device(None, with_obs, with_state, out_obs, out_state,
                                                     random_access_tp=transmission_prob)
        return self._finish(obs, state, reward, done)

    def _finish(self, obs, state, reward, done):
        if self.compat:
            rewards = np.zeros(self.n_agents) + int(reward[0].item())   # env.py:191
            return (self._compat_obs(obs), state[:, 0].cpu().numpy().astype(np.float64), rewards, done, {})
        rewards = reward.to(torch.float32).unsqueeze(1).expand(self.n_envs, self.n_agents)
        return (self._obs_views(obs) if obs is not None else None,
                state.t() if state is not None else None, rewards, done, {})

    @property
    def channel_state(self):
        chan = self._export()[1].view(self.n_agents, self.n_envs)
        return self._maybe_squeeze((chan & 1).t().contiguous())

    @property
    def channel_errors(self):
        return self._stat(0)

    @property
    def n_collisions(self):
        return self._stat(1)
