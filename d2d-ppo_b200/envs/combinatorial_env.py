"""CombinatorialEnv: N devices x C channels, MultiBinary(C) action per device.

Drop-in for envs/combinatorial_env.py:4-264 of the reference (same constructor arguments, reset/step
return types, observation/state layouts and reward), plus ``n_envs`` / ``device`` / ``seed`` / ``rng`` /
``env_offset`` keyword arguments for the batched device path.  The step itself is
``comb_step_kernel`` in csrc/env_kernels.cu.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L
from .. import spaces
from ._base import LockstepEnv


class CombinatorialEnv(LockstepEnv):
    KIND = L.ENV_COMBINATORIAL

    def __init__(self, n_agents, n_channels, deadlines, lbdas, period=5, arrival_probs=None, offsets=None,
                 episode_length=100, traffic_model="aperiodic", periodic_devices=[], reward_type=0,
                 collision_type="pessimistic", homogeneous_size=False, channel_switch=None, verbose=False,
                 *, n_envs=None, device=None, seed=0, rng="philox", env_offset=0):
        self.collision_type = collision_type
        if channel_switch is None:  # combinatorial_env.py:42-45
            self.channel_switch = np.zeros((n_agents, n_channels))
        else:
            self.channel_switch = np.asarray(channel_switch, dtype=np.float64)
        if self.channel_switch.shape != (n_agents, n_channels):
            raise ValueError("channel_switch must have shape (n_agents, n_channels)")
        if n_channels > L.MAX_CHANNELS:
            raise ValueError(f"n_channels must be <= {L.MAX_CHANNELS}")
        self._setup(n_agents=n_agents, n_channels=n_channels, deadlines=deadlines, lbdas=lbdas, period=period,
                    arrival_probs=arrival_probs, offsets=offsets, episode_length=episode_length,
                    traffic_model=traffic_model, periodic_devices=periodic_devices, reward_type=reward_type,
                    switch_probs=self.channel_switch, homogeneous_size=homogeneous_size, verbose=verbose,
                    n_envs=n_envs, device=device, seed=seed, rng=rng, env_offset=env_offset)
        self.action_space = spaces.Tuple([spaces.MultiBinary(self.n_channels) for _ in range(self.n_agents)])
        self.channel_errors = 0            # never incremented by the reference env (combinatorial_env.py:97)
        self.n_collisions = 0

    # ---------------------------------------------------------------- reference API
    def reset(self, *, with_state=True):
        obs, state = self._reset_device(True, with_state)
        if self.compat:
            s = state[:, 0].cpu().numpy().astype(np.float64)
            sd = int(self.deadlines.sum())
            nc = self.n_agents * self.n_channels
            return self._compat_obs(obs), [s[:sd], s[sd:sd + nc], s[sd + nc:]]
        return self._obs_views(obs), (state.t() if state is not None else None)

    def pack_actions(self, actions):
        """[B, N, C] 0/1 (any dtype, device or host) -> channel bitmasks [N, B]."""
        a = torch.as_tensor(actions)
        if a.device != self.device:
            a = a.to(self.device, non_blocking=True)
        a = a.reshape(self.n_envs, self.n_agents, self.n_channels)
        a = (a != 0).to(torch.uint8).contiguous()
        out = torch.empty((self.n_agents, self.n_envs), dtype=self._mask_dtype, device=self.device)
        with torch.cuda.device(self.device):
            L.check(self._lib.d2d_pack_actions(L.ptr(a), L.ptr(out), self.n_envs, self.n_agents, self.n_channels,
                                               L.current_stream()))
        return out

    def _new_ack(self):
        return torch.empty((self.n_channels, self.n_envs), dtype=torch.int8, device=self.device)

    def step(self, actions, *, packed=False, with_obs=True, with_state=True, out_obs=None, out_state=None):
        if packed:
            masks = actions
            assert masks.shape == (self.n_agents, self.n_envs) and masks.dtype == self._mask_dtype
        else:
            masks = self.pack_actions(np.asarray(actions) if self.compat else actions)
        obs, state, reward, done = self._step_device(masks, with_obs, with_state, out_obs, out_state)
        return self._finish(obs, state, reward, done)

    def step_random_access(self, transmission_prob, *, with_obs=True, with_state=True, out_obs=None,
                           out_state=None, return_actions=False):
        """One step with the fused CombinatorialRandomAccess policy (algorithms/baselines.py:181-183)."""
        if self.compat:
            with_obs = with_state = True
        acts = torch.empty((self.n_agents, self.n_envs), dtype=self._mask_dtype, device=self.device) \
            if return_actions else None
        obs, state, reward, done = self._step_device(None, with_obs, with_state, out_obs, out_state,
                                                     random_access_tp=transmission_prob, actions_out=acts)
        out = self._finish(obs, state, reward, done)
        return out + (acts,) if return_actions else out

    def _finish(self, obs, state, reward, done):
        if self.compat:
            s = state[:, 0].cpu().numpy().astype(np.float64)
            sd = int(self.deadlines.sum())
            nc = self.n_agents * self.n_channels
            rewards = np.array([int(reward[0].item()) for _ in range(self.n_agents)])
            return self._compat_obs(obs), [s[:sd], s[sd:sd + nc], s[sd + nc:]], rewards, done, {}
        rewards = reward.unsqueeze(1).expand(self.n_envs, self.n_agents)
        return (self._obs_views(obs) if obs is not None else None,
                state.t() if state is not None else None, rewards, done, {})

    @property
    def channel_state(self):
        chan = self._export()[1].view(self.n_agents, self.n_envs)
        bits = self._bits(chan, self.n_channels).permute(1, 0, 2).contiguous()
        return self._maybe_squeeze(bits)
