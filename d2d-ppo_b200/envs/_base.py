"""Shared host logic of the three lockstep environments (plumbing around the C ABI).

The reference envs are single-instance numpy objects (envs/combinatorial_env.py:4, envs/env.py:4,
envs/channel_selection_env.py:4).  Here one Python object owns B lockstep instances whose state lives
in HBM as an env-minor structure of arrays inside ``libd2d_b200.so``; ``reset``/``step`` keep the
reference signatures and return types, with two modes:

* ``n_envs=None`` (default): reference-compatible single env.  ``reset()``/``step()`` return host numpy
  objects with the reference's exact shapes and dtypes (list of N float64 arrays, list-of-arrays state,
  ``(N,)`` rewards, Python ``bool`` done), so existing callers run unchanged.
* ``n_envs=B``: batched.  ``obs`` is a list of N device tensors ``[B, obs_dim_k]`` (views of one env-minor
  ``[rows, B]`` buffer), ``state`` a ``[B, S]`` tensor, ``rewards`` ``[B, N]``, ``done`` a Python bool (all
  envs finish together; ``done_tensor`` holds the per-env flags the kernel wrote).

There is no auto-reset, as in the reference: the caller resets after ``done``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib as L
from .. import _rng
from .. import spaces


def _as_int_list(x):
    return [int(v) for v in x]


class LockstepEnv:
    KIND = None  # L.ENV_*

    # ------------------------------------------------------------------ construction
    def _setup(self, *, n_agents, n_channels, deadlines, lbdas, period, arrival_probs, offsets,
               episode_length, traffic_model, periodic_devices, reward_type, switch_probs,
               homogeneous_size=False, neighbourhoods=None, verbose=False,
               n_envs=None, device=None, seed=0, rng="philox", env_offset=0):
        self.verbose = verbose
        self.n_agents = int(n_agents)
        self.n_channels = int(n_channels)
        self.lbdas = lbdas
        self.period = period
        self.deadlines = np.asarray(deadlines)
        self.arrival_probs = arrival_probs
        self.offsets = offsets
        self.episode_length = int(episode_length)
        self.traffic_model = traffic_model
        self.reward_type = reward_type
        self.periodic_devices = periodic_devices
        self.aperiodic_devices = [i for i in range(self.n_agents) if i not in _as_int_list(periodic_devices)]
        self.homogeneous_size = bool(homogeneous_size)
        if traffic_model not in ("aperiodic", "periodic", "heterogeneous"):
            raise ValueError("traffic model not supported")
        if traffic_model == "heterogeneous":
            assert len(_as_int_list(periodic_devices)) > 0 and len(self.aperiodic_devices) > 0, \
                "periodic_devices and aperiodic_devices must be non empty"
        if self.n_agents > L.MAX_AGENTS:
            raise ValueError(f"n_agents must be <= {L.MAX_AGENTS}")
        if rng not in ("philox", "replay"):
            raise ValueError("rng must be 'philox' or 'replay'")

        self.compat = n_envs is None
        self.n_envs = 1 if self.compat else int(n_envs)
        if not torch.cuda.is_available():
            raise RuntimeError("d2d-ppo_b200 environments need a CUDA device (B200); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("d2d-ppo_b200 environments run on CUDA devices only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.seed, self.rng, self.env_offset = int(seed), rng, int(env_offset)

        N, T = self.n_agents, self.episode_length
        dl = np.ascontiguousarray(self.deadlines, dtype=np.int32)
        kind = np.zeros(N, dtype=np.int32)
        cdf = np.zeros((N, _rng.POISSON_KMAX), dtype=np.uint32)
        bern = np.zeros(N, dtype=np.uint64)
        periodic = set(_as_int_list(periodic_devices))
        for k in range(N):
            is_bern = traffic_model == "periodic" or (traffic_model == "heterogeneous" and k in periodic)
            kind[k] = L.ARRIVAL_BERNOULLI if is_bern else L.ARRIVAL_POISSON
            if is_bern:
                bern[k] = _rng.bernoulli_thr32(arrival_probs[k])
            else:
                _rng.check_poisson_rate(lbdas[k], f"lbdas[{k}]")
                cdf[k] = _rng.poisson_cdf_table(lbdas[k])
        active = np.array([self._active_mask(t) for t in range(T + 1)], dtype=np.uint64)
        sw = np.ascontiguousarray([_rng.bernoulli_thr16(p) for p in np.asarray(switch_probs, dtype=np.float64).ravel()],
                                  dtype=np.uint32)
        nbr_off = nbr_idx = None
        if neighbourhoods is not None:
            off = [0]
            idx = []
            for nb in neighbourhoods:
                idx += _as_int_list(nb)
                off.append(len(idx))
            nbr_off = np.asarray(off, dtype=np.int32)
            nbr_idx = np.asarray(idx, dtype=np.int32)

        def p(a, ct):
            return None if a is None else a.ctypes.data_as(C.POINTER(ct))

        cfg = L.EnvConfig(
            kind=self.KIND, n_envs=self.n_envs, n_agents=N, n_channels=self.n_channels, episode_length=T,
            homogeneous_size=int(self.homogeneous_size), rng_mode=L.RNG_PHILOX if rng == "philox" else L.RNG_REPLAY,
            reserved0=0, seed=self.seed & 0xFFFFFFFFFFFFFFFF, env_offset=self.env_offset,
            deadlines=p(dl, C.c_int32), arrival_kind=p(kind, C.c_int32), arrival_active=p(active, C.c_uint64),
            poisson_cdf=p(cdf, C.c_uint32), bernoulli_thr=p(bern, C.c_uint64), switch_thr=p(sw, C.c_uint32),
            nbr_offset=p(nbr_off, C.c_int32), nbr_index=p(nbr_idx, C.c_int32))
        self._lib = L.lib()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self._lib.d2d_env_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._obs_rows = self._lib.d2d_env_obs_rows(h)
        self._state_rows = self._lib.d2d_env_state_rows(h)
        self._obs_off = [self._lib.d2d_env_obs_offset(h, k) for k in range(N)]
        self._obs_dim = [self._lib.d2d_env_obs_dim(h, k) for k in range(N)]
        self._rec = self._lib.d2d_env_record_bytes(h)
        self._mask_bytes = self._lib.d2d_env_mask_bytes(h)
        self._mask_dtype = {1: torch.uint8, 2: torch.int16, 4: torch.int32}[self._mask_bytes]
        self._replay_keepalive = None
        self.timestep = 0
        self.done_tensor = None
        self.last_ack = None
        self.observation_space = spaces.Tuple([spaces.Box(shape=(d,)) for d in self._obs_dim])
        self.state_space = spaces.Box(shape=(self._state_rows,))

    def _active_mask(self, t):
        """Which devices draw an arrival at timestep t (0 = reset); python/numpy `%` semantics of the
        reference (combinatorial_env.py:66-83 for reset, :178-196 for step)."""
        N = self.n_agents
        if self.traffic_model == "aperiodic":
            act = range(N)
        elif self.traffic_model == "periodic":
            off = np.asarray(self.offsets)
            act = np.where(off == 0)[0] if t == 0 else np.where(t % np.asarray(self.period) == off)[0]
        else:
            act = list(self.aperiodic_devices)
            for i in _as_int_list(self.periodic_devices):
                if (self.offsets[i] == 0) if t == 0 else (t % self.period[i] == self.offsets[i]):
                    act.append(i)
        m = 0
        for i in act:
            m |= 1 << int(i)
        return m

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.d2d_env_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ episode index of the Philox streams
    @property
    def episode(self):
        """Index of the current episode (-1 before the first reset).  Every reset starts a fresh Philox stream."""
        return int(self._lib.d2d_env_episode(self._h))

    def set_episode(self, next_episode):
        """The next reset() starts episode ``next_episode`` (resume a run, or reproduce a given episode)."""
        L.check(self._lib.d2d_env_set_episode(self._h, int(next_episode)))

    # ------------------------------------------------------------------ replay streams (parity runs)
    def set_replay(self, arrivals, switches):
        """arrivals [T+1, B, N] ints (value device k draws at timestep t; t = 0 is reset);
        switches [T+1, B, ...] 0/1 flip draws (index 0 unused): combinatorial [.., N, C],
        single-channel [.., N], selection [.., C+1]."""
        dev = self.device
        arr = torch.as_tensor(np.asarray(arrivals), device=dev)
        if arr.min() < 0 or arr.max() > 255:
            raise ValueError("replayed arrivals must be in 0..255")
        arr = arr.to(torch.uint8).permute(0, 2, 1).contiguous()        # [T+1][N][B]
        sw = torch.as_tensor(np.asarray(switches), device=dev).to(torch.int64)
        if self.KIND == L.ENV_COMBINATORIAL:
            w = (1 << torch.arange(self.n_channels, device=dev, dtype=torch.int64))
            m = (sw * w).sum(-1)                                         # [T+1, B, N]
            m = m.permute(0, 2, 1)
        elif self.KIND == L.ENV_SINGLE_CHANNEL:
            m = sw.permute(0, 2, 1)
        else:
            w = (1 << torch.arange(self.n_channels + 1, device=dev, dtype=torch.int64))
            m = (sw * w).sum(-1)                                         # [T+1, B]
        if self._mask_bytes == 4:
            m = torch.where(m >= (1 << 31), m - (1 << 32), m)
        elif self._mask_bytes == 2:
            m = torch.where(m >= (1 << 15), m - (1 << 16), m)
        m = m.to(self._mask_dtype).contiguous()
        t_len = arr.shape[0]
        L.check(self._lib.d2d_env_set_replay(self._h, L.ptr(arr), L.ptr(m), t_len))
        self._replay_keepalive = (arr, m)

    # ------------------------------------------------------------------ reset / step plumbing
    def _new_outputs(self, with_obs, with_state):
        B, dev = self.n_envs, self.device
        obs = torch.empty((self._obs_rows, B), dtype=torch.float32, device=dev) if with_obs else None
        state = torch.empty((self._state_rows, B), dtype=torch.float32, device=dev) if with_state else None
        return obs, state

    def _obs_views(self, obs):
        return [obs[o:o + d].t() for o, d in zip(self._obs_off, self._obs_dim)]

    def _reset_device(self, with_obs=True, with_state=True, out_obs=None, out_state=None):
        obs, state = self._new_outputs(with_obs and out_obs is None, with_state and out_state is None)
        obs = out_obs if out_obs is not None else obs
        state = out_state if out_state is not None else state
        with torch.cuda.device(self.device):
            L.check(self._lib.d2d_env_reset(self._h, L.ptr(obs), L.ptr(state), L.current_stream()))
        self.timestep = 0
        self.obs_rows_tensor, self.state_rows_tensor = obs, state
        return obs, state

    def _step_device(self, actions_dev, with_obs=True, with_state=True, out_obs=None, out_state=None,
                     random_access_tp=None, actions_out=None, out_reward=None):
        B, dev = self.n_envs, self.device
        obs, state = self._new_outputs(with_obs and out_obs is None, with_state and out_state is None)
        obs = out_obs if out_obs is not None else obs
        state = out_state if out_state is not None else state
        reward = out_reward if out_reward is not None else torch.empty(B, dtype=torch.int32, device=dev)
        done = torch.empty(B, dtype=torch.uint8, device=dev)
        ack = self._new_ack()
        with torch.cuda.device(dev):
            if random_access_tp is None:
                L.check(self._lib.d2d_env_step(self._h, L.ptr(actions_dev), L.ptr(obs), L.ptr(state), L.ptr(reward),
                                               L.ptr(done), L.ptr(ack), L.current_stream()))
            else:
                L.check(self._lib.d2d_env_step_random_access(
                    self._h, float(random_access_tp), L.ptr(actions_out), L.ptr(obs), L.ptr(state), L.ptr(reward),
                    L.ptr(done), L.ptr(ack), L.current_stream()))
        self.timestep += 1
        self.done_tensor, self.last_ack = done, ack
        self.obs_rows_tensor, self.state_rows_tensor = obs, state
        return obs, state, reward, self.timestep >= self.episode_length

    def _new_ack(self):
        return None

    def run_random_access(self, transmission_prob, n_steps, *, auto_reset=False, out_obs=None, obs_stride=0,
                          out_state=None, state_stride=0, out_reward=None, reward_stride=0, accumulate=False):
        """``n_steps`` fused random-access steps enqueued by the library in one call (d2d_env_run_random_access):
        the inner loop of ``CombinatorialRandomAccess.run`` (baselines.py:199-213) without a Python round trip per
        step.  out_obs / out_state: env-minor blocks, step i writes at element offset i * stride (stride 0: the same
        block every step; None: not emitted).  out_reward int32: [B] (stride 0; ``accumulate=True`` adds every
        step's reward into it) or [n_steps, B] with reward_stride = B.  auto_reset: reset whenever an episode ends
        (or before the first step); otherwise the run stops at the end of the episode.  Returns the steps run."""
        if out_reward is None:
            out_reward = torch.zeros(self.n_envs, dtype=torch.int32, device=self.device)
            self.last_run_reward = out_reward
        done = torch.empty(self.n_envs, dtype=torch.uint8, device=self.device)
        n_done = C.c_int(0)
        with torch.cuda.device(self.device):
            L.check(self._lib.d2d_env_run_random_access(
                self._h, float(transmission_prob), int(n_steps), int(bool(auto_reset)), L.ptr(out_obs),
                int(obs_stride), L.ptr(out_state), int(state_stride), L.ptr(out_reward), int(reward_stride),
                int(bool(accumulate)), L.ptr(done), L.current_stream(), C.byref(n_done)))
        self.timestep = int(self._lib.d2d_env_timestep(self._h))
        self.done_tensor = done
        return int(n_done.value)

    # ------------------------------------------------------------------ host-buffer step (pipelined copies)
    def step_host(self, host_actions, host_reward, *, layout="reference", host_done=None, with_obs=True,
                  with_state=False, out_obs=None, out_state=None):
        """step(actions) for callers whose actions and rewards live in HOST memory (d2d_env_step_host).

        host_actions: pinned CPU tensor; layout "reference" = u8 [B, N, C] 0/1 (the reference's (N, C) array per
        env, combinatorial env only), layout "device" = the [N, B] device layout (masks / flags / channel ids).
        host_reward: pinned int32 [B] CPU tensor that receives the per-env reward; host_done: optional pinned u8 [B].
        Asynchronous: returns a ticket; ``host_wait(ticket)`` blocks until host_reward of that call is valid.  The
        H2D copy, the kernels and the D2H copy of consecutive calls overlap (two calls in flight), so issue call
        k + 1 before waiting for call k.  Observations / state stay on the device (``obs_rows_tensor``)."""
        if host_actions.device.type != "cpu" or host_reward.device.type != "cpu":
            raise ValueError("step_host takes CPU (pinned) tensors; use step() for device tensors")
        if host_reward.dtype != torch.int32 or host_reward.numel() != self.n_envs or not host_reward.is_contiguous():
            raise ValueError("host_reward must be a contiguous int32 [B] tensor")
        if not host_actions.is_contiguous():
            raise ValueError("host_actions must be contiguous")
        N, B = self.n_agents, self.n_envs
        if layout == "reference":
            want, code = (B * N * self.n_channels, torch.uint8), L.ACT_HOST_REFERENCE
        elif layout == "device":
            comb = self.KIND == L.ENV_COMBINATORIAL
            want, code = (N * B, self._mask_dtype if comb else torch.uint8), L.ACT_HOST_DEVICE_LAYOUT
        else:
            raise ValueError("layout must be 'reference' or 'device'")
        if host_actions.numel() != want[0] or host_actions.dtype != want[1]:
            raise ValueError(f"host_actions must hold {want[0]} elements of {want[1]} for layout '{layout}'")
        obs, state = self._new_outputs(with_obs and out_obs is None, with_state and out_state is None)
        obs = out_obs if out_obs is not None else obs
        state = out_state if out_state is not None else state
        ack = self._new_ack()
        ticket = C.c_uint64()
        with torch.cuda.device(self.device):
            L.check(self._lib.d2d_env_step_host(self._h, L.ptr(host_actions), code, L.ptr(obs), L.ptr(state),
                                                L.ptr(host_reward), L.ptr(host_done), L.ptr(ack),
                                                L.current_stream(), C.byref(ticket)))
        self.timestep += 1
        self.last_ack = ack
        self.obs_rows_tensor, self.state_rows_tensor = obs, state
        return int(ticket.value)

    @property
    def host_pack_state(self):
        """-1: host-side action packing not used by step_host (so far); 1: active; 0: switched off after the timed
        first calls because this host packs slower than PCIe moves the unpacked bytes (d2d_env_host_pack_state)."""
        return int(self._lib.d2d_env_host_pack_state(self._h))

    def host_wait(self, ticket):
        """Block until the host buffers of the step_host call `ticket` are valid."""
        L.check(self._lib.d2d_env_host_wait(self._h, int(ticket)))

    # ------------------------------------------------------------------ zero-copy entry points for the learners
    def reset_into(self, out_obs, out_state=None):
        """reset() writing the env-minor observation block [obs_rows, B] (and state block) in place."""
        self._reset_device(True, out_state is not None, out_obs, out_state)

    def step_into(self, actions_nb, out_obs, out_state=None, out_reward=None):
        """step() on device-layout actions [N, B], writing observation / state / reward blocks in place.
        Returns the lockstep done flag."""
        return self._step_device(actions_nb, True, out_state is not None, out_obs, out_state,
                                 out_reward=out_reward)[3]

    @property
    def obs_layout(self):
        """(rows per time block, [first row of agent k], [obs_dim of agent k]) of the env-minor obs matrix."""
        return self._obs_rows, list(self._obs_off), list(self._obs_dim)

    @property
    def action_kind(self):
        """'bernoulli_mask' (combinatorial) | 'binary' (D2DEnv) | 'index' (channel selection)."""
        return {L.ENV_COMBINATORIAL: "bernoulli_mask", L.ENV_SINGLE_CHANNEL: "binary",
                L.ENV_CHANNEL_SELECTION: "index"}[self.KIND]

    # ------------------------------------------------------------------ raw state (reference attribute names)
    def _export(self):
        N, B, dev = self.n_agents, self.n_envs, self.device
        buf = torch.empty((N, B, self._rec), dtype=torch.uint8, device=dev)
        n_chan = B if self.KIND == L.ENV_CHANNEL_SELECTION else N * B
        chan = torch.empty(n_chan, dtype=self._mask_dtype, device=dev)
        disc = torch.empty((N, B), dtype=torch.int32, device=dev)
        recv = torch.empty((N, B), dtype=torch.int32, device=dev)
        stats = torch.empty((2, B), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            L.check(self._lib.d2d_env_export_state(self._h, L.ptr(buf), L.ptr(chan), L.ptr(disc), L.ptr(recv),
                                                   L.ptr(stats), L.current_stream()))
        return buf, chan, disc, recv, stats

    def import_state(self, *, channel_masks=None, buffers=None, discarded=None, received=None, stats=None,
                     timestep=None):
        """Overwrite raw device state (inverse of the export behind ``current_buffers`` etc.).
        channel_masks: integer bitmasks, combinatorial / single-channel [N, B], selection [B];
        buffers: uint8 [N, B, record_bytes]; discarded / received: int32 [N, B]; stats int32 [2, B]."""
        def dev(x, dtype):
            return None if x is None else torch.as_tensor(x, device=self.device).to(dtype).contiguous()
        chan = dev(channel_masks, self._mask_dtype)
        buf, disc, recv, st = dev(buffers, torch.uint8), dev(discarded, torch.int32), dev(received, torch.int32), \
            dev(stats, torch.int32)
        t = self.timestep if timestep is None else int(timestep)
        with torch.cuda.device(self.device):
            L.check(self._lib.d2d_env_import_state(self._h, L.ptr(buf), L.ptr(chan), L.ptr(disc), L.ptr(recv),
                                                   L.ptr(st), t, L.current_stream()))
            torch.cuda.current_stream().synchronize()   # sources may be temporaries
        self.timestep = t

    def _maybe_squeeze(self, t, as_float=True):
        if self.compat:
            a = t[0].cpu().numpy()
            return a.astype(np.float64) if as_float else a
        return t

    @property
    def current_buffers(self):
        buf = self._export()[0]
        D = int(self.deadlines.max())
        return self._maybe_squeeze(buf.permute(1, 0, 2)[:, :, :D].contiguous())

    @property
    def discarded_packets(self):
        return self._maybe_squeeze(self._export()[2].t().contiguous())

    @property
    def received_packets(self):
        return self._maybe_squeeze(self._export()[3].t().contiguous())

    def _stat(self, i):
        s = self._export()[4][i]
        return int(s[0].item()) if self.compat else s

    def _bits(self, masks, n_bits):
        m = masks.to(torch.int64) & ((1 << (8 * self._mask_bytes)) - 1)
        return ((m.unsqueeze(-1) >> torch.arange(n_bits, device=m.device)) & 1).to(torch.uint8)

    # ------------------------------------------------------------------ metrics (combinatorial_env.py:245-264)
    def _scores(self):
        B, dev = self.n_envs, self.device
        out = torch.empty((3, B), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            L.check(self._lib.d2d_env_scores(self._h, L.ptr(out[0]), L.ptr(out[1]), L.ptr(out[2]), L.current_stream()))
        return out

    def compute_urllc(self):
        s = self._scores()[0]
        return float(s[0].item()) if self.compat else s

    def compute_jains(self):
        s = self._scores()[1]
        return float(s[0].item()) if self.compat else s

    def compute_channel_score(self):
        s = self._scores()[2]
        return float(s[0].item()) if self.compat else s

    # ------------------------------------------------------------------ compat helpers
    def _compat_obs(self, obs):
        return [o[0].cpu().numpy().astype(np.float64) for o in self._obs_views(obs)]
