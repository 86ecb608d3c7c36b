"""numpy restatement of the three channel-access environments, batched over B lockstep envs.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity pinning: validated
against the unmodified reference through ``oracle/ref_harness.py`` and the
committed fixtures in ``tests/golden/env_*.npz`` (``tests/test_oracle_envs.py``).

What is restated (file:line relative to /root/reference):

* ``CombinatorialOracle``      <- envs/combinatorial_env.py:61-114 (reset), :127-242 (step),
                                  :116-118 (evolve_channel), :120-124 (evolve_buffer), :245-264 (metrics)
* ``D2DOracle``                <- envs/env.py:51-101 (reset), :118-217 (step), :103-109, :220-233
* ``ChannelSelectionOracle``   <- envs/channel_selection_env.py:49-98 (reset), :116-214 (step), :104-107

Differences from the reference that are NOT behavioural: state is integer
(the reference keeps small integers in float64 arrays), there is a leading
batch axis, and random draws come from a ``DrawSource`` instead of the global
``np.random`` so that streams can be replayed (parity), produced by numpy
(CPU baseline, same distributions as the reference) or by the Philox transforms
the CUDA kernels use (throughput-mode parity).
"""
from __future__ import annotations

import numpy as np

from . import philox_np as px


# --------------------------------------------------------------------------------------
# traffic description shared by the three envs
# --------------------------------------------------------------------------------------
class Traffic:
    """Who draws an arrival at timestep t, and from which distribution.

    Mirrors the three ``traffic_model`` branches (combinatorial_env.py:66-85 for
    reset, :178-196 for step; identical code in env.py and channel_selection_env.py).
    """

    def __init__(self, n_agents, traffic_model, lbdas, period, arrival_probs, offsets, periodic_devices):
        self.n = int(n_agents)
        self.model = traffic_model
        self.lbdas = None if lbdas is None else np.asarray(lbdas, dtype=np.float64)
        self.period = period
        self.arrival_probs = None if arrival_probs is None else np.asarray(arrival_probs, dtype=np.float64)
        self.offsets = None if offsets is None else np.asarray(offsets)
        self.periodic = [int(i) for i in periodic_devices]
        self.aperiodic = [i for i in range(self.n) if i not in self.periodic]
        if traffic_model not in ("aperiodic", "periodic", "heterogeneous"):
            raise ValueError("traffic model not supported")
        if traffic_model == "heterogeneous":
            assert self.periodic != [] and self.aperiodic != [], \
                "periodic_devices and aperiodic_devices must be non empty"

    def draws(self, t):
        """[(device, 'poisson'|'bernoulli')] in the reference's draw order for timestep t (0 = reset)."""
        if self.model == "aperiodic":
            return [(i, "poisson") for i in range(self.n)]
        if self.model == "periodic":
            if t == 0:
                act = np.where(self.offsets == 0)[0]
            else:
                act = np.where(t % self.period == self.offsets)[0]
            return [(int(i), "bernoulli") for i in act]
        out = [(i, "poisson") for i in self.aperiodic]
        for i in self.periodic:
            if (self.offsets[i] == 0) if t == 0 else (t % self.period[i] == self.offsets[i]):
                out.append((i, "bernoulli"))
        return out


# --------------------------------------------------------------------------------------
# draw sources
# --------------------------------------------------------------------------------------
class ReplaySource:
    """Pre-drawn streams: arrivals [T+1, B, N] ints; switches [T+1, B, ...] 0/1 (index 0 unused)."""

    def __init__(self, arrivals, switches):
        self.arrivals, self.switches = np.asarray(arrivals), np.asarray(switches)

    def arrival(self, t, dev, kind, traffic):
        return self.arrivals[t, :, dev].astype(np.int64)

    def switch(self, t, p):
        return self.switches[t].astype(np.int64)


class NumpySource:
    """Same distributions as the reference, from a numpy Generator (CPU-baseline timing)."""

    def __init__(self, n_envs, seed=0):
        self.rng = np.random.default_rng(seed)
        self.B = n_envs

    def arrival(self, t, dev, kind, traffic):
        if kind == "poisson":
            return self.rng.poisson(traffic.lbdas[dev], self.B)
        return self.rng.binomial(1, traffic.arrival_probs[dev], self.B)

    def switch(self, t, p):
        p = np.asarray(p, dtype=np.float64)
        return self.rng.binomial(1, np.broadcast_to(p, (self.B,) + p.shape))


class PhiloxSource:
    """The CUDA throughput-mode streams (see oracle/philox_np.py for the layout)."""

    def __init__(self, n_envs, seed, env_offset=0, env_level_switch=False):
        self.env = np.arange(n_envs, dtype=np.uint64) + np.uint64(env_offset)
        self.seed = seed
        self.env_level = env_level_switch
        self._cdf = {}
        self.episode, self.episode_length = -1, None

    def begin_episode(self, episode_length):
        """Every reset() starts a fresh stream: the counter's timestep field is t + episode * (T + 1)
        (csrc/env_kernels.cu: fill_args), so episode e of an env does not replay episode 0's traffic."""
        self.episode += 1
        self.episode_length = int(episode_length)

    def ctr_t(self, t):
        """Timestep field of the Philox counter for env timestep t of the current episode."""
        return (int(t) + max(self.episode, 0) * (self.episode_length + 1 if self.episode_length else 0)) & 0xFFFFFFFF

    def arrival(self, t, dev, kind, traffic):
        t = self.ctr_t(t)
        u = px.word32(self.seed, self.env, t, dev, px.PURPOSE_ARRIVAL)
        if kind == "poisson":
            lam = float(traffic.lbdas[dev])
            if lam not in self._cdf:
                self._cdf[lam] = px.poisson_cdf_thresholds(lam)
            return px.poisson_from_u32(u, self._cdf[lam])
        return (u.astype(np.uint64) < np.uint64(px.thr32(traffic.arrival_probs[dev]))).astype(np.int64)

    def switch(self, t, p):
        t = self.ctr_t(t)
        p = np.asarray(p, dtype=np.float64)
        if self.env_level:  # ChannelSelectionEnv: one vector of C+1 channels per env
            lanes = px.lanes16(self.seed, self.env, t, px.ENV_LEVEL_DEVICE, px.PURPOSE_SWITCH, p.shape[0])
            thr = np.array([px.thr16(x) for x in p], dtype=np.uint32)
            return (lanes < thr[None, :]).astype(np.int64)
        if p.ndim == 0:  # D2DEnv: scalar p, one channel per device -> caller passes shape via n
            raise ValueError("scalar switch prob must be broadcast by the caller")
        if p.ndim == 1:  # D2DEnv [N]
            out = np.empty((self.env.shape[0], p.shape[0]), dtype=np.int64)
            for k in range(p.shape[0]):
                out[:, k] = px.lane16_shared(self.seed, self.env, t, k, px.PURPOSE_SWITCH) < px.thr16(p[k])
            return out
        out = np.empty((self.env.shape[0],) + p.shape, dtype=np.int64)  # Combinatorial [N, C]
        for k in range(p.shape[0]):
            lanes = px.lanes16(self.seed, self.env, t, k, px.PURPOSE_SWITCH, p.shape[1])
            thr = np.array([px.thr16(x) for x in p[k]], dtype=np.uint32)
            out[:, k, :] = lanes < thr[None, :]
        return out


# --------------------------------------------------------------------------------------
# shared buffer mechanics
# --------------------------------------------------------------------------------------
class _BufferEnv:
    def _init_common(self, n_envs, n_agents, deadlines, episode_length, traffic, source):
        self.B, self.n_agents = int(n_envs), int(n_agents)
        self.deadlines = np.asarray(deadlines, dtype=np.int64)
        self.D = int(self.deadlines.max())
        self.episode_length = int(episode_length)
        self.traffic, self.source = traffic, source

    def _reset_buffers(self):
        self.timestep = 0
        if hasattr(self.source, "begin_episode"):
            self.source.begin_episode(self.episode_length)
        self.buffers = np.zeros((self.B, self.n_agents, self.D), dtype=np.int64)
        for dev, kind in self.traffic.draws(0):
            self.buffers[:, dev, self.deadlines[dev] - 1] = self.source.arrival(0, dev, kind, self.traffic)
        self.discarded = np.zeros((self.B, self.n_agents), dtype=np.int64)
        self.received = self.buffers.sum(2)

    def _has_packet(self):
        return self.buffers.sum(2) > 0

    def _serve_age_arrive(self, success):
        """success [B,N] bool -> pop earliest packet, age by one slot, count drops, draw arrivals."""
        nb = self.buffers.copy()
        earliest = (nb > 0).argmax(2)
        b, k = np.nonzero(success)
        nb[b, k, earliest[b, k]] -= 1
        self.discarded += nb[:, :, 0]
        nb = np.concatenate([nb[:, :, 1:], np.zeros((self.B, self.n_agents, 1), dtype=np.int64)], axis=2)
        return nb

    def _arrivals(self, nb):
        for dev, kind in self.traffic.draws(self.timestep):
            a = self.source.arrival(self.timestep, dev, kind, self.traffic)
            nb[:, dev, self.deadlines[dev] - 1] = a
            self.received[:, dev] += a
        return nb

    def _all_buffers(self, buf):
        return np.concatenate([buf[:, i, : self.deadlines[i]] for i in range(self.n_agents)], axis=1)

    # metrics, per env -------------------------------------------------------------
    def compute_urllc(self):
        return 1.0 - self.discarded.sum(1) / self.received.sum(1)

    def compute_jains(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.where(self.received > 0, 1.0 - self.discarded / np.maximum(self.received, 1), 1.0)
        return s.sum(1) ** 2 / self.n_agents / (s ** 2).sum(1)


# --------------------------------------------------------------------------------------
class CombinatorialOracle(_BufferEnv):
    def __init__(self, n_envs, n_agents, n_channels, deadlines, lbdas, period=5, arrival_probs=None,
                 offsets=None, episode_length=100, traffic_model="aperiodic", periodic_devices=(),
                 homogeneous_size=False, channel_switch=None, source=None):
        traffic = Traffic(n_agents, traffic_model, lbdas, period, arrival_probs, offsets, periodic_devices)
        self._init_common(n_envs, n_agents, deadlines, episode_length, traffic, source)
        self.n_channels = int(n_channels)
        self.homogeneous_size = bool(homogeneous_size)
        self.channel_switch = (np.zeros((n_agents, n_channels)) if channel_switch is None
                               else np.asarray(channel_switch, dtype=np.float64))

    def _obs(self, buf, chan_obs, ack):
        obs = []
        for k in range(self.n_agents):
            bk = buf[:, k] if self.homogeneous_size else buf[:, k, : self.deadlines[k]]
            obs.append(np.concatenate([bk, chan_obs[:, k], ack], axis=1).astype(np.float32))
        return obs

    def reset(self):
        self._reset_buffers()
        self.channel_state = np.ones((self.B, self.n_agents, self.n_channels), dtype=np.int64)
        ones = np.ones((self.B, self.n_channels), dtype=np.int64)
        obs = self._obs(self.buffers, np.ones_like(self.channel_state), ones)
        state = np.concatenate([self._all_buffers(self.buffers), self.channel_state.reshape(self.B, -1), ones],
                               axis=1).astype(np.float32)
        self.last_ack = ones
        return obs, state

    def step(self, actions):
        actions = np.asarray(actions).reshape(self.B, self.n_agents, self.n_channels)
        self.timestep += 1
        attempts = (actions != 0) & self._has_packet()[:, :, None]
        good = attempts & (self.channel_state != 0)
        users = attempts.sum(1)
        ack = np.full((self.B, self.n_channels), -1, dtype=np.int64)
        ack[(good.sum(1) == 1) & (users == 1)] = 1
        ack[users == 0] = 0
        success = (good & (ack[:, None, :] == 1)).any(2)
        chan_obs = self.channel_state.copy()
        nb = self._serve_age_arrive(success)
        self.channel_state = self.channel_state ^ self.source.switch(self.timestep, self.channel_switch)
        nb = self._arrivals(nb)
        obs = self._obs(nb, chan_obs, ack)
        state = np.concatenate([self._all_buffers(nb), self.channel_state.reshape(self.B, -1), ack],
                               axis=1).astype(np.float32)
        rewards = np.repeat(success.sum(1)[:, None], self.n_agents, axis=1)
        self.buffers, self.last_ack = nb, ack
        return obs, state, rewards, self.timestep >= self.episode_length, {}


# --------------------------------------------------------------------------------------
class D2DOracle(_BufferEnv):
    def __init__(self, n_envs, n_agents, deadlines, lbdas, period=5, arrival_probs=None, offsets=None,
                 episode_length=100, traffic_model="aperiodic", periodic_devices=(), channel_switch=0.2,
                 neighbourhoods=None, source=None):
        traffic = Traffic(n_agents, traffic_model, lbdas, period, arrival_probs, offsets, periodic_devices)
        self._init_common(n_envs, n_agents, deadlines, episode_length, traffic, source)
        self.channel_switch = float(channel_switch)
        self.neighbourhoods = [[k] for k in range(n_agents)] if neighbourhoods is None else \
            [list(map(int, nb)) for nb in neighbourhoods]

    def _obs_state(self, buf, ack):
        obs = []
        for k in range(self.n_agents):
            parts = [buf[:, i, : self.deadlines[i]] for i in self.neighbourhoods[k]]
            parts.append(self.channel_state[:, self.neighbourhoods[k]])
            parts.append(ack[:, None])
            obs.append(np.concatenate(parts, axis=1).astype(np.float32))
        state = np.concatenate([self._all_buffers(buf), self.channel_state, ack[:, None]], axis=1).astype(np.float32)
        return obs, state

    def reset(self):
        self._reset_buffers()
        self.channel_state = np.ones((self.B, self.n_agents), dtype=np.int64)
        self.channel_errors = np.zeros(self.B, dtype=np.int64)
        self.n_collisions = np.zeros(self.B, dtype=np.int64)
        return self._obs_state(self.buffers, np.zeros(self.B, dtype=np.int64))

    def step(self, actions):
        actions = np.asarray(actions).reshape(self.B, self.n_agents)
        self.timestep += 1
        attempts = (actions != 0) & self._has_packet()
        n_att = attempts.sum(1)
        who = attempts.argmax(1)
        lone = n_att == 1
        decoded = lone & (self.channel_state[np.arange(self.B), who] != 0)
        ack = np.where(n_att > 1, -1, np.where(decoded, 1, 0)).astype(np.int64)
        self.channel_errors += lone & ~decoded
        self.n_collisions += n_att > 1
        success = attempts & decoded[:, None]
        nb = self._serve_age_arrive(success)
        p = np.full(self.n_agents, self.channel_switch)
        self.channel_state = self.channel_state ^ self.source.switch(self.timestep, p)
        nb = self._arrivals(nb)
        obs, state = self._obs_state(nb, ack)
        rewards = np.repeat(ack[:, None], self.n_agents, axis=1).astype(np.float64)
        self.buffers = nb
        return obs, state, rewards, self.timestep >= self.episode_length, {}


# --------------------------------------------------------------------------------------
class ChannelSelectionOracle(_BufferEnv):
    def __init__(self, n_envs, n_agents, n_channels, deadlines, lbdas, period=5, arrival_probs=None,
                 offsets=None, episode_length=100, traffic_model="aperiodic", periodic_devices=(),
                 channel_switch=None, source=None):
        traffic = Traffic(n_agents, traffic_model, lbdas, period, arrival_probs, offsets, periodic_devices)
        self._init_common(n_envs, n_agents, deadlines, episode_length, traffic, source)
        self.n_channels = int(n_channels)
        # the reference's default is zeros(n_agents) (channel_selection_env.py:35-36) but it is indexed by
        # channel 0..C (:105); a usable default therefore needs C+1 entries.
        self.channel_switch = (np.zeros(n_channels + 1) if channel_switch is None
                               else np.asarray(channel_switch, dtype=np.float64))

    def _obs_state(self, buf, ack):
        obs = [np.concatenate([buf[:, k, : self.deadlines[k]].astype(np.float64), ack], axis=1).astype(np.float32)
               for k in range(self.n_agents)]
        state = np.concatenate([self._all_buffers(buf), self.channel_state], axis=1).astype(np.float32)
        return obs, state

    def reset(self):
        self._reset_buffers()
        self.channel_state = np.ones((self.B, self.n_channels + 1), dtype=np.int64)
        self.selected_channel_qualities = np.zeros(self.B, dtype=np.int64)
        self.number_selected_channel = np.zeros(self.B, dtype=np.int64)
        return self._obs_state(self.buffers, np.zeros((self.B, self.n_channels + 1)))

    def step(self, actions):
        actions = np.asarray(actions).reshape(self.B, self.n_agents).astype(np.int64)
        self.timestep += 1
        C1 = self.n_channels + 1
        attempts = actions * self._has_packet()
        onehot = (attempts[:, :, None] == np.arange(C1)[None, None, :]) & (attempts[:, :, None] != 0)
        counts = onehot.sum(1)                       # [B, C+1]; column 0 is always 0
        good = self.channel_state != 0
        ack = np.zeros((self.B, C1), dtype=np.float64)
        ack[(counts > 0) & ~good] = -1.0
        sel_good = (counts > 0) & good
        ack[sel_good] = 1.0 / counts[sel_good]       # float64 division, as in the reference (:137)
        self.selected_channel_qualities += sel_good.sum(1)
        self.number_selected_channel += (counts > 0).sum(1)
        win = (counts == 1) & good                   # good channels with exactly one attempt
        success = (onehot & win[:, None, :]).any(2)
        nb = self._serve_age_arrive(success)
        self.channel_state = self.channel_state ^ self.source.switch(self.timestep, self.channel_switch)
        nb = self._arrivals(nb)
        obs, state = self._obs_state(nb, ack)
        rewards = np.repeat(success.sum(1)[:, None], self.n_agents, axis=1)
        self.buffers = nb
        return obs, state, rewards, self.timestep >= self.episode_length, {}

    def compute_channel_score(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(self.number_selected_channel != 0,
                            self.selected_channel_qualities / np.maximum(self.number_selected_channel, 1), 1.0)
