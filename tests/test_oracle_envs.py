"""The numpy oracle reproduces the reference's outputs stored in tests/golden (and the live reference, if mounted)."""
import numpy as np
import pytest

from _helpers import cat_obs, env_case_names, load_env_case, make_oracle
from oracle.envs_np import PhiloxSource, ReplaySource
from oracle.ref_harness import reference_available


@pytest.mark.parametrize("name", env_case_names())
def test_oracle_matches_golden(name):
    g = load_env_case(name)
    kind, kw = g["kind"], g["config"]
    T, B = g["actions"].shape[:2]
    env = make_oracle(kind, kw, B, ReplaySource(g["arrivals"], g["switches"]))
    obs, state = env.reset()
    assert np.array_equal(cat_obs(obs), g["obs0"])
    assert np.array_equal(state, g["state0"])
    assert np.array_equal(env.buffers, g["buffers0"])
    for t in range(T):
        obs, state, rew, done, _ = env.step(g["actions"][t])
        assert np.array_equal(cat_obs(obs), g["obs"][t]), (name, t)
        assert np.array_equal(state, g["state"][t]), (name, t)
        assert np.array_equal(rew.astype(np.float32), g["rewards"][t]), (name, t)
        assert done == bool(g["done"][t])
        assert np.array_equal(env.buffers, g["buffers"][t])
        assert np.array_equal(env.discarded, g["discarded"][t])
        assert np.array_equal(env.received, g["received"][t])
        assert np.array_equal(env.channel_state, g["channel"][t])
    assert np.allclose(env.compute_urllc(), g["urllc"], rtol=0, atol=1e-15)
    assert np.allclose(env.compute_jains(), g["jains"], rtol=1e-15)
    if kind == "d2d":
        assert np.array_equal(env.channel_errors, g["channel_errors"])
        assert np.array_equal(env.n_collisions, g["n_collisions"])
    if kind == "channel_selection":
        assert np.allclose(env.compute_channel_score(), g["channel_score"], rtol=1e-15)


def test_survey_known_answer_step():
    """hand-checked CombinatorialEnv step (N=4, C=3)."""
    g = load_env_case("kat_survey4")
    env = make_oracle("combinatorial", g["config"], 1, ReplaySource(g["arrivals"], g["switches"]))
    env.reset()
    assert env.buffers[0].tolist() == [[0, 0, 2], [0, 0, 1], [0, 0, 0], [0, 0, 1]]
    env.channel_state = g["forced_channel"][None].astype(np.int64)
    obs, state, rew, done, _ = env.step(g["actions"][0])
    assert env.last_ack[0].tolist() == [-1, 1, -1]
    assert rew[0].tolist() == [1, 1, 1, 1]
    assert env.buffers[0].tolist() == [[0, 2, 0], [0, 0, 1], [0, 0, 0], [0, 1, 0]]
    assert env.discarded[0].tolist() == [0, 0, 0, 0] and env.received[0].tolist() == [2, 2, 0, 1]
    assert obs[0][0].tolist() == [0, 2, 0, 1, 1, 0, -1, 1, -1]       # channel part is pre-switch
    assert env.channel_state[0].tolist() == [[1, 1, 1]] * 4
    # and the fixture written by the reference says the same
    assert np.array_equal(cat_obs(obs)[0], g["obs"]) and np.array_equal(state[0], g["state"])
    assert np.array_equal(env.buffers[0], g["buffers"]) and np.array_equal(env.channel_state[0], g["channel"])


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
def test_oracle_matches_live_reference_random_config():
    from oracle.gen_golden import draw_actions, draw_streams, run_reference_env
    rng = np.random.default_rng(99)
    kw = dict(n_agents=5, n_channels=6, deadlines=[4, 9, 2, 9, 6], lbdas=[0.7] * 5, period=[2] * 5,
              arrival_probs=[0.3, 0.6, 0.9, 1.0, 0.5], offsets=[0, 1, 0, 1, 0], episode_length=25,
              traffic_model="heterogeneous", homogeneous_size=False, periodic_devices=[1, 4],
              channel_switch=rng.uniform(0, 1, (5, 6)).tolist())
    arr, sw = draw_streams("combinatorial", kw, 3, 25, rng)
    act = draw_actions("combinatorial", kw, 3, 25, 0.35, rng)
    ref = run_reference_env("combinatorial", kw, arr, sw, act)
    env = make_oracle("combinatorial", kw, 3, ReplaySource(arr, sw))
    obs, state = env.reset()
    assert np.array_equal(cat_obs(obs), ref["obs0"])
    for t in range(25):
        obs, state, rew, done, _ = env.step(act[t])
        assert np.array_equal(cat_obs(obs), ref["obs"][t]) and np.array_equal(state, ref["state"][t])
        assert np.array_equal(env.discarded, ref["discarded"][t]) and np.array_equal(env.received, ref["received"][t])


def test_philox_source_is_deterministic_and_sharded():
    """Env streams depend on the global env index only: two shards reproduce the unsharded run."""
    g = load_env_case("comb_c3_load0.33")
    kw = g["config"]
    full = make_oracle("combinatorial", kw, 8, PhiloxSource(8, seed=5))
    lo = make_oracle("combinatorial", kw, 4, PhiloxSource(4, seed=5, env_offset=0))
    hi = make_oracle("combinatorial", kw, 4, PhiloxSource(4, seed=5, env_offset=4))
    for e in (full, lo, hi):
        e.reset()
    rng = np.random.default_rng(0)
    for t in range(20):
        a = rng.binomial(1, 0.3, (8, 6, 8))
        full.step(a), lo.step(a[:4]), hi.step(a[4:])
    assert np.array_equal(full.buffers[:4], lo.buffers) and np.array_equal(full.buffers[4:], hi.buffers)
    assert np.array_equal(full.channel_state[4:], hi.channel_state)
    assert full.received.sum() > 0
