"""CUDA envs (through the C ABI) against the oracle and the reference's golden outputs -- bit exact."""
import numpy as np
import pytest

from _helpers import cat_obs, env_case_names, load_env_case, make_cuda_env, make_oracle, to_np

pytestmark = pytest.mark.gpu


def _check_state(env, orc, kind, tag):
    assert np.array_equal(to_np(env.current_buffers), orc.buffers), tag
    assert np.array_equal(to_np(env.discarded_packets), orc.discarded), tag
    assert np.array_equal(to_np(env.received_packets), orc.received), tag
    assert np.array_equal(to_np(env.channel_state), orc.channel_state), tag


def _rollout_against_oracle(env, orc, kind, actions, check_every=1):
    T = actions.shape[0]
    obs, state = env.reset()
    o_obs, o_state = orc.reset()
    assert np.array_equal(cat_obs(obs), cat_obs(o_obs))
    assert np.array_equal(to_np(state), o_state)
    _check_state(env, orc, kind, "reset")
    for t in range(T):
        obs, state, rew, done, _ = env.step(actions[t])
        o_obs, o_state, o_rew, o_done, _ = orc.step(actions[t])
        assert np.array_equal(cat_obs(obs), cat_obs(o_obs)), ("obs", t)
        assert np.array_equal(to_np(state), o_state), ("state", t)
        assert np.array_equal(to_np(rew).astype(np.float64), o_rew.astype(np.float64)), ("reward", t)
        assert done == o_done
        assert np.array_equal(to_np(env.done_tensor), np.full(env.n_envs, int(o_done), dtype=np.uint8))
        if t % check_every == 0 or t == T - 1:
            _check_state(env, orc, kind, t)
    assert np.allclose(to_np(env.compute_urllc()), orc.compute_urllc(), rtol=0, atol=1e-15, equal_nan=True)
    assert np.allclose(to_np(env.compute_jains()), orc.compute_jains(), rtol=1e-14, equal_nan=True)


@pytest.mark.parametrize("name", env_case_names())
def test_cuda_env_matches_reference_golden(name, cuda_device):
    import torch
    g = load_env_case(name)
    kind, kw = g["kind"], g["config"]
    T, B = g["actions"].shape[:2]
    env = make_cuda_env(kind, kw, B, rng="replay", device=cuda_device)
    env.set_replay(g["arrivals"], g["switches"])
    obs, state = env.reset()
    assert np.array_equal(cat_obs(obs), g["obs0"])
    assert np.array_equal(to_np(state), g["state0"])
    assert np.array_equal(to_np(env.current_buffers), g["buffers0"])
    for t in range(T):
        obs, state, rew, done, _ = env.step(torch.as_tensor(g["actions"][t]))
        assert np.array_equal(cat_obs(obs), g["obs"][t]), (name, t)
        assert np.array_equal(to_np(state), g["state"][t]), (name, t)
        assert np.array_equal(to_np(rew).astype(np.float32), g["rewards"][t]), (name, t)
        assert done == bool(g["done"][t])
        assert np.array_equal(to_np(env.current_buffers), g["buffers"][t])
        assert np.array_equal(to_np(env.discarded_packets), g["discarded"][t])
        assert np.array_equal(to_np(env.received_packets), g["received"][t])
        assert np.array_equal(to_np(env.channel_state), g["channel"][t])
    assert np.allclose(to_np(env.compute_urllc()), g["urllc"], rtol=0, atol=1e-15)
    assert np.allclose(to_np(env.compute_jains()), g["jains"], rtol=1e-14)
    if kind == "d2d":
        assert np.array_equal(to_np(env.channel_errors), g["channel_errors"])
        assert np.array_equal(to_np(env.n_collisions), g["n_collisions"])
    if kind == "channel_selection":
        assert np.allclose(to_np(env.compute_channel_score()), g["channel_score"], rtol=1e-15)


def test_survey_known_answer_step_cuda(cuda_device):
    g = load_env_case("kat_survey4")
    env = make_cuda_env("combinatorial", g["config"], 1, rng="replay", device=cuda_device)
    env.set_replay(g["arrivals"], g["switches"])
    env.reset()
    masks = (g["forced_channel"].astype(np.int64) * (1 << np.arange(3))).sum(1)[:, None]   # [N, B=1]
    env.import_state(channel_masks=masks)
    obs, state, rew, done, _ = env.step(g["actions"][0])
    assert to_np(env.last_ack)[:, 0].tolist() == [-1, 1, -1]
    assert to_np(rew)[0].tolist() == [1, 1, 1, 1]
    assert to_np(env.current_buffers)[0].tolist() == [[0, 2, 0], [0, 0, 1], [0, 0, 0], [0, 1, 0]]
    assert np.array_equal(cat_obs(obs)[0], g["obs"]) and np.array_equal(to_np(state)[0], g["state"])
    assert to_np(env.received_packets)[0].tolist() == [2, 2, 0, 1]
    assert np.array_equal(to_np(env.channel_state)[0], g["channel"])


@pytest.mark.parametrize("name,B,T", [("comb_c3_load0.33", 1000, 200), ("comb_c1_16ch", 333, 60),
                                      ("comb_c4_n12_aperiodic", 257, 50), ("comb_deadline20_c20", 65, 45),
                                      ("d2d_c2", 4096, 200), ("d2d_neighbourhoods", 100, 50),
                                      ("sel_xp_gamma", 500, 50), ("sel_heterogeneous", 129, 40)])
def test_cuda_env_matches_oracle_replay_large(name, B, T, cuda_device):
    from oracle.envs_np import ReplaySource
    from oracle.gen_golden import draw_actions, draw_streams
    g = load_env_case(name)
    kind, kw = g["kind"], dict(g["config"])
    kw["episode_length"] = T
    rng = np.random.default_rng(abs(hash(name)) % 1000)
    arr, sw = draw_streams(kind, kw, B, T, rng)
    act = draw_actions(kind, kw, B, T, 0.3, rng)
    env = make_cuda_env(kind, kw, B, rng="replay", device=cuda_device)
    env.set_replay(arr, sw)
    orc = make_oracle(kind, kw, B, ReplaySource(arr, sw))
    _rollout_against_oracle(env, orc, kind, act, check_every=7)


@pytest.mark.parametrize("name,offset", [("comb_c3_load0.33", 0), ("comb_c1_16ch", 12345), ("comb_c4_n12_aperiodic", 7),
                                         ("comb_periodic_offsets", 0), ("comb_deadline20_c20", 3), ("d2d_c2", 99),
                                         ("sel_xp_gamma", 5), ("sel_heterogeneous", 0)])
def test_cuda_philox_mode_matches_oracle(name, offset, cuda_device):
    """Throughput-mode streams: device Philox == numpy Philox restatement, env state bit exact."""
    from oracle.envs_np import PhiloxSource
    from oracle.gen_golden import draw_actions
    g = load_env_case(name)
    kind, kw = g["kind"], g["config"]
    B, T = 300, kw["episode_length"]
    act = draw_actions(kind, kw, B, T, 0.3, np.random.default_rng(3))
    env = make_cuda_env(kind, kw, B, rng="philox", seed=0x1234567890ABCDEF, env_offset=offset, device=cuda_device)
    orc = make_oracle(kind, kw, B, PhiloxSource(B, 0x1234567890ABCDEF, env_offset=offset,
                                                env_level_switch=(kind == "channel_selection")))
    _rollout_against_oracle(env, orc, kind, act, check_every=5)


@pytest.mark.parametrize("n_agents,load", [(4, 1 / 14), (16, 1 / 3), (32, 0.8), (64, 1.0)])
def test_n_agents_sweep_matches_oracle(n_agents, load, cuda_device):
    """Config c4 (xp_n_agents.py:62-83 with xp_load.py traffic levels): N = 4..64 devices on 4 channels,
    collision-heavy; device Philox streams against the numpy restatement, bit exact."""
    from d2d_ppo_b200.presets import n_agents_sweep_kwargs
    from oracle.envs_np import PhiloxSource
    from oracle.gen_golden import draw_actions
    kw = n_agents_sweep_kwargs(n_agents, load=load, episode_length=30)
    B, T = 97, 30
    act = draw_actions("combinatorial", kw, B, T, 0.4, np.random.default_rng(n_agents))
    env = make_cuda_env("combinatorial", kw, B, rng="philox", seed=n_agents, env_offset=5, device=cuda_device)
    orc = make_oracle("combinatorial", kw, B, PhiloxSource(B, n_agents, env_offset=5))
    _rollout_against_oracle(env, orc, "combinatorial", act, check_every=3)


def test_maximum_sizes_match_oracle(cuda_device):
    """The build's limits in one env: 32 channels, deadlines up to 32, ragged observations, Poisson load 1."""
    from oracle.envs_np import PhiloxSource
    from oracle.gen_golden import draw_actions
    N, C = 5, 32
    kw = dict(n_agents=N, n_channels=C, deadlines=np.array([32, 1, 17, 32, 8]), lbdas=np.array([1.0] * N), period=None,
              arrival_probs=None, offsets=None, episode_length=25, traffic_model="aperiodic", periodic_devices=[],
              homogeneous_size=False, channel_switch=np.linspace(0.0, 1.0, N * C).reshape(N, C))
    B, T = 70, 25
    act = draw_actions("combinatorial", kw, B, T, 0.1, np.random.default_rng(9))
    env = make_cuda_env("combinatorial", kw, B, rng="philox", seed=17, device=cuda_device)
    orc = make_oracle("combinatorial", kw, B, PhiloxSource(B, 17))
    _rollout_against_oracle(env, orc, "combinatorial", act, check_every=4)


def test_fused_random_access_policy(cuda_device):
    """step_random_access == oracle step fed with the Philox policy-stream action bits."""
    from oracle import philox_np as px
    from oracle.envs_np import PhiloxSource
    g = load_env_case("comb_c3_load0.33")
    kw = g["config"]
    B, T, seed, tp = 500, kw["episode_length"], 77, 0.3
    env = make_cuda_env("combinatorial", kw, B, rng="philox", seed=seed, device=cuda_device)
    orc = make_oracle("combinatorial", kw, B, PhiloxSource(B, seed))
    env.reset(), orc.reset()
    envs = np.arange(B)
    for t in range(1, T + 1):
        obs, state, rew, done, _, acts = env.step_random_access(tp, return_actions=True)
        a = np.stack([px.lanes16(seed, envs, t, k, px.PURPOSE_POLICY, 8) < px.thr16(tp) for k in range(6)], axis=1)
        packed = (a.astype(np.int64) * (1 << np.arange(8))).sum(-1)            # [B, N]
        assert np.array_equal(to_np(acts).astype(np.uint8).T, packed.astype(np.uint8))
        o_obs, o_state, o_rew, o_done, _ = orc.step(a)
        assert np.array_equal(cat_obs(obs), cat_obs(o_obs)) and np.array_equal(to_np(state), o_state)
        assert np.array_equal(to_np(rew), o_rew)
    assert np.array_equal(to_np(env.received_packets), orc.received)
    assert np.array_equal(to_np(env.discarded_packets), orc.discarded)


@pytest.mark.parametrize("layout", ["reference", "reference-device-pack", "device"])
def test_host_buffer_step_matches_oracle(layout, cuda_device):
    """step_host (d2d_env_step_host: pinned host actions in, pinned host rewards out, pipelined copies) against the
    oracle on Philox streams; the host reads step i - 1 after issuing step i, as bench.py's e2e loop does.
    "reference": the [B, N, C] bytes are packed on the host (12 pool threads) before the copy; "reference-device-pack":
    D2D_SWITCH_HOST_PACK = 0, they are copied as they are and packed by a kernel."""
    import torch
    from d2d_ppo_b200 import _lib as L
    from oracle.envs_np import PhiloxSource
    saved = L.lib().d2d_get_host_threads()
    L.check(L.lib().d2d_set_host_threads(12))
    L.set_kernel_switch(L.SWITCH_HOST_PACK, layout != "reference-device-pack")
    try:
        _host_buffer_step_case("reference" if layout.startswith("reference") else layout, cuda_device)
    finally:
        L.set_kernel_switch(L.SWITCH_HOST_PACK, True)
        L.check(L.lib().d2d_set_host_threads(saved))


def _host_buffer_step_case(layout, cuda_device):
    import torch
    from oracle.envs_np import PhiloxSource
    g = load_env_case("comb_c3_load0.33")
    kw = g["config"]
    B, T, seed = 777, 40, 5
    env = make_cuda_env("combinatorial", kw, B, rng="philox", seed=seed, device=cuda_device)
    orc = make_oracle("combinatorial", kw, B, PhiloxSource(B, seed))
    env.reset(), orc.reset()
    rng = np.random.default_rng(3)
    acts = rng.binomial(1, 0.3, (T, B, 6, 8)).astype(np.uint8)
    w = (1 << np.arange(8)).astype(np.uint8)
    host_a = [torch.from_numpy(acts[t] if layout == "reference"
                               else np.ascontiguousarray((acts[t] * w).sum(-1).astype(np.uint8).T)).pin_memory()
              for t in range(T)]
    host_r = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
    host_d = [torch.empty(B, dtype=torch.uint8).pin_memory() for _ in range(2)]
    expected, pending = [], None
    for t in range(T):
        tk = env.step_host(host_a[t], host_r[t % 2], layout=layout, host_done=host_d[t % 2], with_state=True)
        obs_rows, state_rows = env.obs_rows_tensor, env.state_rows_tensor
        o_obs, o_state, o_rew, o_done, _ = orc.step(acts[t])
        expected.append((o_rew[:, 0].copy(), o_done))
        if pending is not None:
            env.host_wait(pending)
            assert np.array_equal(host_r[(t - 1) % 2].numpy(), expected[t - 1][0]), ("reward", t - 1)
            assert np.array_equal(host_d[(t - 1) % 2].numpy(), np.full(B, int(expected[t - 1][1]), np.uint8))
        pending = tk
        assert np.array_equal(to_np(obs_rows.t()), np.concatenate(o_obs, axis=1)), ("obs", t)
        assert np.array_equal(to_np(state_rows.t()), o_state), ("state", t)
    env.host_wait(pending)
    assert np.array_equal(host_r[(T - 1) % 2].numpy(), expected[T - 1][0])
    assert np.array_equal(to_np(env.received_packets), orc.received)
    assert np.array_equal(to_np(env.discarded_packets), orc.discarded)
    with pytest.raises(ValueError):
        env.step_host(host_a[0][:1], host_r[0])


def test_host_buffer_step_single_channel_env(cuda_device):
    """step_host on D2DEnv (device action layout u8 [N, B]) against step() on an identically seeded env."""
    import torch
    g = load_env_case("d2d_c2")
    kw = g["config"]
    B, T = 333, 25
    a_env = make_cuda_env("d2d", kw, B, rng="philox", seed=9, device=cuda_device)
    b_env = make_cuda_env("d2d", kw, B, rng="philox", seed=9, device=cuda_device)
    a_env.reset(), b_env.reset()
    rng = np.random.default_rng(1)
    host_r = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
    pending = None
    want = []
    for t in range(T):
        act = rng.binomial(1, 0.3, (B, kw["n_agents"])).astype(np.uint8)
        host_a = torch.from_numpy(np.ascontiguousarray(act.T)).pin_memory()
        tk = a_env.step_host(host_a, host_r[t % 2], layout="device", with_state=True)
        obs, state, rew, done, _ = b_env.step(act)
        want.append(to_np(rew)[:, 0].astype(np.int32))
        assert np.array_equal(to_np(a_env.obs_rows_tensor.t()), cat_obs(obs)) and np.array_equal(
            to_np(a_env.state_rows_tensor.t()), to_np(state))
        if pending is not None:
            a_env.host_wait(pending)
            assert np.array_equal(host_r[(t - 1) % 2].numpy(), want[t - 1])
        pending = tk
    a_env.host_wait(pending)
    assert np.array_equal(host_r[(T - 1) % 2].numpy(), want[T - 1])
    with pytest.raises(Exception):
        a_env.step_host(torch.zeros((B, kw["n_agents"], 1), dtype=torch.uint8).pin_memory(), host_r[0], layout="reference")


def test_fused_random_access_policy_single_channel(cuda_device):
    """D2DEnv.step_random_access == oracle step fed with the shared-lane Philox policy bits (one call per eight
    devices), with neighbourhood observations."""
    from oracle import philox_np as px
    from oracle.envs_np import PhiloxSource
    g = load_env_case("d2d_neighbourhoods")
    kw = dict(g["config"])
    B, T, seed, tp = 301, 30, 13, 0.35
    kw["episode_length"] = T
    N = kw["n_agents"]
    env = make_cuda_env("d2d", kw, B, rng="philox", seed=seed, env_offset=17, device=cuda_device)
    orc = make_oracle("d2d", kw, B, PhiloxSource(B, seed, env_offset=17))
    env.reset(), orc.reset()
    envs = np.arange(B) + 17
    for t in range(1, T + 1):
        obs, state, rew, done, _ = env.step_random_access(tp)
        a = np.stack([px.lane16_shared(seed, envs, t, k, px.PURPOSE_POLICY) < px.thr16(tp) for k in range(N)], axis=1)
        o_obs, o_state, o_rew, o_done, _ = orc.step(a.astype(np.int64))
        assert np.array_equal(cat_obs(obs), cat_obs(o_obs)) and np.array_equal(to_np(state), o_state), t
        assert np.array_equal(to_np(rew).astype(np.float64), o_rew.astype(np.float64))
    assert np.array_equal(to_np(env.received_packets), orc.received)
    assert np.array_equal(to_np(env.discarded_packets), orc.discarded)


def test_reference_compatible_single_env_mode(cuda_device):
    """n_envs=None: host numpy outputs with the reference's shapes and dtypes."""
    g = load_env_case("comb_c3_load1_ragged_obs")
    kw = g["config"]
    env = make_cuda_env("combinatorial", kw, None, rng="replay", device=cuda_device)
    env.set_replay(g["arrivals"][:, :1], g["switches"][:, :1])
    obs, state = env.reset()
    assert isinstance(obs, list) and len(obs) == 6 and obs[0].dtype == np.float64 and obs[0].shape == (7 + 16,)
    assert isinstance(state, list) and len(state) == 3
    assert np.array_equal(np.concatenate(obs).astype(np.float32), g["obs0"][0])
    for t in range(5):
        obs, state, rew, done, info = env.step(g["actions"][t, 0].astype(np.float64))
        assert np.array_equal(np.concatenate(obs).astype(np.float32), g["obs"][t, 0])
        assert np.array_equal(np.concatenate(state).astype(np.float32), g["state"][t, 0])
        assert rew.shape == (6,) and np.array_equal(rew.astype(np.float32), g["rewards"][t, 0])
        assert done is False and info == {}
    assert env.current_buffers.shape == (6, 14) and env.discarded_packets.shape == (6,)
    assert isinstance(env.compute_urllc(), float)
    assert env.observation_space[1].shape == (14 + 16,) and env.action_space[0].n == 8
    assert env.state_space.shape == (63 + 8 * 7,)


def test_error_behaviour(cuda_device):
    from d2d_ppo_b200._lib import D2DError
    g = load_env_case("comb_periodic_offsets")
    kw = dict(g["config"])
    kw["episode_length"] = 3
    env = make_cuda_env("combinatorial", kw, 4, device=cuda_device)
    a = np.zeros((4, 4, 3), dtype=np.uint8)
    with pytest.raises(D2DError):
        env.step(a)                       # step before reset
    env.reset()
    for _ in range(3):
        _, _, _, done, _ = env.step(a)
    assert done is True
    with pytest.raises(D2DError):
        env.step(a)                       # the reference has no auto-reset either; stepping past T is an error
    env.reset()
    env.step(a)
    replay = make_cuda_env("combinatorial", kw, 4, rng="replay", device=cuda_device)
    with pytest.raises(D2DError):
        replay.reset()                    # replay mode without streams
    with pytest.raises(ValueError):
        make_cuda_env("combinatorial", dict(kw, traffic_model="bursty"), 4, device=cuda_device)


def test_full_size_conservation_and_determinism(cuda_device):
    """BASELINE config c3/c5 shape at 1M envs: packet conservation, a checksum of checksums, determinism."""
    import torch
    from d2d_ppo_b200.presets import combinatorial_kwargs
    B, T = 1 << 20, 40
    kw = combinatorial_kwargs("setup_8_channels", load=2 / 3, episode_length=T)
    sums = []
    for rep in range(2):
        env = make_cuda_env("combinatorial", kw, B, seed=2024, device=cuda_device)
        env.reset()
        delivered = torch.zeros(B, dtype=torch.int64, device=cuda_device)
        for t in range(T):
            _, _, rew, done, _ = env.step_random_access(0.2, with_state=False)
            delivered += rew[:, 0]
        assert done
        recv = env.received_packets.to(torch.int64).sum(1)
        disc = env.discarded_packets.to(torch.int64).sum(1)
        held = env.current_buffers.to(torch.int64).sum((1, 2))
        assert torch.equal(recv, delivered + disc + held)          # every packet is delivered, dropped or queued
        assert int(delivered.sum()) > 0 and int(disc.sum()) > 0
        sums.append((int(recv.sum()), int(disc.sum()), int(delivered.sum()),
                     int((env.current_buffers.to(torch.int64) * 31).sum())))
        u = env.compute_urllc()
        assert float(u.min()) >= 0.0 and float(u.max()) <= 1.0
    assert sums[0] == sums[1]


@pytest.mark.parametrize("name", ["comb_c3_load0.33", "d2d_c2", "sel_xp_gamma"])
def test_philox_episodes_are_independent_and_match_oracle(name, cuda_device):
    """Every reset() starts a fresh Philox stream (counter timestep field t + episode * (T + 1)): three consecutive
    episodes match the oracle bit for bit, differ from each other, and set_episode() reproduces a given episode."""
    from oracle.envs_np import PhiloxSource
    from oracle.gen_golden import draw_actions
    g = load_env_case(name)
    kind, kw = g["kind"], dict(g["config"])
    kw["episode_length"] = 12
    B, T = 64, 12
    act = draw_actions(kind, kw, B, T, 0.3, np.random.default_rng(1))
    env = make_cuda_env(kind, kw, B, rng="philox", seed=77, env_offset=3, device=cuda_device)
    orc = make_oracle(kind, kw, B, PhiloxSource(B, 77, env_offset=3, env_level_switch=(kind == "channel_selection")))
    received = []
    for episode in range(3):
        _rollout_against_oracle(env, orc, kind, act, check_every=4)
        assert env.episode == episode
        received.append(to_np(env.received_packets).copy())
    assert not np.array_equal(received[0], received[1]) and not np.array_equal(received[1], received[2])
    env.set_episode(1)
    env.reset()
    for t in range(T):
        env.step(act[t])
    assert env.episode == 1 and np.array_equal(to_np(env.received_packets), received[1])


@pytest.mark.parametrize("name", ["comb_c3_load0.33", "d2d_c2"])
def test_run_random_access_equals_stepwise(name, cuda_device):
    """d2d_env_run_random_access (steps enqueued by the library) == the same steps issued one call at a time:
    observations, states, per-step rewards, counters, across an automatic reset; reward accumulation."""
    import torch
    g = load_env_case(name)
    kind, kw = g["kind"], dict(g["config"])
    kw["episode_length"] = T = 9
    B, tp, n = 200, 0.3, 2 * T + 4
    a = make_cuda_env(kind, kw, B, rng="philox", seed=5, device=cuda_device)
    b = make_cuda_env(kind, kw, B, rng="philox", seed=5, device=cuda_device)
    R, S = a.obs_layout[0], a.state_space.shape[0]
    obs = torch.zeros((n, R, B), device=cuda_device)
    state = torch.zeros((n, S, B), device=cuda_device)
    rew = torch.zeros((n, B), dtype=torch.int32, device=cuda_device)
    assert a.run_random_access(tp, n, auto_reset=True, out_obs=obs, obs_stride=R * B, out_state=state,
                               state_stride=S * B, out_reward=rew, reward_stride=B) == n
    for i in range(n):
        if i % T == 0:
            b.reset()
        _, _, r, done, _ = b.step_random_access(tp)
        assert torch.equal(obs[i], b.obs_rows_tensor) and torch.equal(state[i], b.state_rows_tensor), i
        assert torch.equal(rew[i].to(r.dtype), r[:, 0]), i
        assert done == ((i + 1) % T == 0)
    assert a.timestep == b.timestep == n - 2 * T and a.episode == b.episode == 2
    assert np.array_equal(to_np(a.received_packets), to_np(b.received_packets))
    assert np.array_equal(to_np(a.discarded_packets), to_np(b.discarded_packets))
    # without auto_reset the run stops at the end of the episode; accumulate sums the rewards of the steps run
    a.reset()
    acc = torch.zeros(B, dtype=torch.int32, device=cuda_device)
    assert a.run_random_access(tp, T + 5, out_reward=acc, accumulate=True) == T
    b.reset()
    tot = sum(b.step_random_access(tp)[2][:, 0].to(torch.int32) for _ in range(T))
    assert torch.equal(acc, tot)
    assert a.run_random_access(tp, 3, out_reward=acc, accumulate=True) == 0


@pytest.mark.parametrize("n_agents,B", [(64, 96), (40, 64), (32, 32)])
def test_grouped_device_mapping_matches_oracle(n_agents, B, cuda_device):
    """Many devices per env and few envs (xp_n_agents.py sweep at N >= 32, B a multiple of 32): the step kernel maps a
    block to 32 envs x 8 device groups and combines the per-channel masks through shared memory.  Bit exact against
    the oracle with actions from memory and with the fused random-access policy."""
    from d2d_ppo_b200.presets import n_agents_sweep_kwargs
    from oracle import philox_np as px
    from oracle.envs_np import PhiloxSource
    from oracle.gen_golden import draw_actions
    kw = n_agents_sweep_kwargs(n_agents, load=0.5, episode_length=20)
    T, C = 20, kw["n_channels"]
    act = draw_actions("combinatorial", kw, B, T, 0.15, np.random.default_rng(n_agents))
    env = make_cuda_env("combinatorial", kw, B, rng="philox", seed=n_agents + 1, env_offset=11, device=cuda_device)
    orc = make_oracle("combinatorial", kw, B, PhiloxSource(B, n_agents + 1, env_offset=11))
    _rollout_against_oracle(env, orc, "combinatorial", act, check_every=3)
    # fused policy: the oracle is fed the action bits the kernel draws (Philox policy stream of episode 1)
    env.reset()
    orc.reset()
    envs = np.arange(B) + 11
    tp = 0.1
    for t in range(1, T + 1):
        obs, state, rew, done, _ = env.step_random_access(tp)
        tc = orc.source.ctr_t(t)
        a = np.stack([px.lanes16(n_agents + 1, envs, tc, k, px.PURPOSE_POLICY, C) < px.thr16(tp)
                      for k in range(n_agents)], axis=1)
        o_obs, o_state, o_rew, o_done, _ = orc.step(a)
        assert np.array_equal(cat_obs(obs), cat_obs(o_obs)), t
        assert np.array_equal(to_np(state), o_state) and np.array_equal(to_np(rew), o_rew), t
    _check_state(env, orc, "combinatorial", "fused policy")


def test_selection_env_fused_random_access_matches_oracle(cuda_device):
    """RandomAccess (algorithms/baselines.py:10-14) fused into sel_step_kernel: the channel ids the kernel draws from
    the Philox policy stream equal the numpy restatement (uniform over 0..C, 0 for devices without a packet), and the
    env driven by them matches the oracle bit for bit; ids are uniform over the channels."""
    from oracle import philox_np as px
    from oracle.envs_np import PhiloxSource
    g = load_env_case("sel_xp_gamma")
    kw = dict(g["config"])
    kw["episode_length"] = T = 25
    B, N, C1 = 192, kw["n_agents"], kw["n_channels"] + 1
    env = make_cuda_env("channel_selection", kw, B, rng="philox", seed=31, env_offset=4, device=cuda_device)
    orc = make_oracle("channel_selection", kw, B, PhiloxSource(B, 31, env_offset=4, env_level_switch=True))
    envs = np.arange(B) + 4
    counts = np.zeros(C1, dtype=np.int64)
    for episode in range(2):
        env.reset()
        orc.reset()
        for t in range(1, T + 1):
            has = orc.buffers.sum(2) > 0
            obs, state, rew, done, _, acts = env.step_random_access(return_actions=True)
            tc = orc.source.ctr_t(t)
            want = np.stack([px.uniform_choice(31, envs, tc, k, C1) for k in range(N)], axis=1) * has
            assert np.array_equal(to_np(acts).T.astype(np.int64), want), (episode, t)
            counts += np.bincount(want[has], minlength=C1)
            o_obs, o_state, o_rew, o_done, _ = orc.step(want)
            assert np.array_equal(cat_obs(obs), cat_obs(o_obs)) and np.array_equal(to_np(state), o_state), (episode, t)
            assert np.array_equal(to_np(rew), o_rew) and done == o_done
        _check_state(env, orc, "channel_selection", episode)
    freq = counts / counts.sum()
    assert np.all(np.abs(freq - 1.0 / C1) < 0.01), freq


@pytest.mark.gpu
@pytest.mark.parametrize("B", [333, 40000])      # four lanes per env (sc_run_lanes_kernel) / one thread per env
@pytest.mark.parametrize("deadline,traffic", [(7, "aperiodic"), (14, "aperiodic"), (5, "periodic"), (7, "mixed")])
def test_multistep_kernel_equals_step_kernels(deadline, traffic, B, cuda_device):
    """D2DEnv, N <= 4, no per-step observation rows: d2d_env_run_random_access runs the steps of an episode inside ONE
    kernel with the env state in registers (sc_run_kernel, or sc_run_lanes_kernel up to 32,768 envs).  Bit-identical to the same call with the multi-step kernel
    switched off (one sc_step_kernel launch per step): per-step rewards, done flags, buffers, channel bits, packet
    counters, error / collision counters, across two automatic resets, and for the accumulating reward form."""
    import torch
    from d2d_ppo_b200 import _lib as L
    from d2d_ppo_b200.envs import D2DEnv
    N, T, tp = (3 if traffic == "periodic" else 4), 11, 0.35
    kw = dict(n_agents=N, deadlines=np.array([deadline] * N), lbdas=np.array([1 / 4] * N), episode_length=T,
              traffic_model="aperiodic" if traffic == "aperiodic" else "heterogeneous" if traffic == "mixed" else "periodic",
              channel_switch=0.2)
    if traffic == "periodic":
        kw.update(period=3, arrival_probs=np.array([0.7] * N), offsets=np.array([0, 1, 2]))
    if traffic == "mixed":
        kw.update(period=np.array([4] * N), arrival_probs=np.array([0.6] * N), offsets=np.array([0, 1, 2, 3]),
                  periodic_devices=[1, 3])
    n = 2 * T + 5
    outs = []
    for multi in (1, 0):
        L.set_kernel_switch(L.SWITCH_ENV_MULTISTEP, multi)
        try:
            e = D2DEnv(n_envs=B, device=cuda_device, seed=21, env_offset=1000, **kw)
            rew = torch.full((n, B), -7, dtype=torch.int32, device=cuda_device)
            before = L.launch_count()
            assert e.run_random_access(tp, n, auto_reset=True, out_reward=rew, reward_stride=B) == n
            launches = L.launch_count() - before
            acc = torch.zeros(B, dtype=torch.int32, device=cuda_device)
            e.reset()
            assert e.run_random_access(tp, T + 3, out_reward=acc, accumulate=True) == T
            bufs = [to_np(t) for t in e._export()]
            outs.append((to_np(rew), to_np(acc), bufs, to_np(e.channel_errors), to_np(e.n_collisions), e.timestep,
                         e.episode, launches))
        finally:
            L.set_kernel_switch(L.SWITCH_ENV_MULTISTEP, 1)
    a, b = outs
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[0].min() >= -1 and a[0].max() <= 1 and (a[0] != 0).any()
    for x, y in zip(a[2], b[2]):
        assert np.array_equal(x, y)
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4]) and a[5:7] == b[5:7]
    # 3 episodes -> 3 resets + 3 multi-step launches instead of one launch per step
    assert a[7] == 6 and b[7] == 3 + n
