"""End-to-end learner parity: the iPPO / D2DPPO classes on the CUDA env reproduce one full reference training
iteration (rollout + epochs) from the fixtures, with the env on replayed streams and actions teacher-forced."""
import os

import numpy as np
import pytest
import torch

from _helpers import assert_params_close, load_ppo_case, make_cuda_env, params_from, rel_err, report_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _setup(algo, tag, dev):
    from d2d_ppo_b200.algorithms.d2d_ppo import D2DPPO
    from d2d_ppo_b200.algorithms.ippo import iPPO
    g = load_ppo_case(f"{algo}_{tag}")
    m = g["meta"]
    kind, comb = m.get("kind", "combinatorial"), m.get("combinatorial", True)
    env = make_cuda_env(kind, g["config"], m["E"], rng="replay", device=dev)
    env.set_replay(g["arrivals"][:, :m["E"]], g["switches"][:, :m["E"]])
    kw = dict(hidden_size=m["hidden"], gamma=m["gamma"], policy_lr=m["policy_lr"], value_lr=m["value_lr"],
              useRNN=(m["arch"] == "gru"), combinatorial=comb, history_len=m["L"], early_stopping=False)
    agent = iPPO(env, **kw) if algo == "ippo" else D2DPPO(env, beta_entropy=0.01, **kw)
    for i in range(m["N"]):
        agent.policies.load_state_dict(i, params_from(g, f"init/policy{i}"))
        if algo == "ippo":
            agent.values.load_state_dict(i, params_from(g, f"init/value{i}"))
    if algo == "d2dppo":
        agent.critic.load_state_dict(0, params_from(g, "init/critic"))
    forced = _device_actions(g["actions"], g, m["E"], agent, dev)
    return g, m, env, agent, forced


def _device_actions(actions, g, E, agent, dev):
    """[E*T, N(, C)] reference actions -> the device layout [T, N, E] (channel bitmask / category index)."""
    m = g["meta"]
    if m.get("combinatorial", True):
        C = g["config"]["n_channels"]
        actions = (actions.astype(np.int64) * (1 << np.arange(C))).sum(-1)                        # [R, N]
    t = torch.tensor(actions.astype(np.int64)).reshape(E, m["T"], m["N"]).permute(1, 2, 0).contiguous()
    return t.to(agent.act_buf.dtype).to(dev)


def _rows(t):     # [T, N, B] -> [B*T, N] episode-major ; [T, B] -> [B*T]
    if t.dim() == 3:
        T, N, B = t.shape
        return t.permute(2, 0, 1).reshape(B * T, N).cpu().numpy()
    return t.t().reshape(-1).cpu().numpy()


def _check_greedy_test(g, m, agent, algo, tag, dev):
    """test() (d2d_ppo.py:341-383, ippo.py:345-388) with the reference's UPDATED policies on the fixture's extra
    replayed episodes: greedy actions bit exact, the 4-tuple to 1e-12."""
    if not m.get("E_test"):
        return
    from d2d_ppo_b200.algorithms.d2d_ppo import D2DPPO
    from d2d_ppo_b200.algorithms.ippo import iPPO
    E, Et = m["E"], m["E_test"]
    env = make_cuda_env(m.get("kind", "combinatorial"), g["config"], Et, rng="replay", device=dev)
    env.set_replay(g["arrivals"][:, E:E + Et], g["switches"][:, E:E + Et])
    kw = dict(hidden_size=m["hidden"], gamma=m["gamma"], policy_lr=m["policy_lr"], value_lr=m["value_lr"],
              useRNN=(m["arch"] == "gru"), combinatorial=m.get("combinatorial", True), history_len=m["L"],
              early_stopping=False)
    ev = iPPO(env, **kw) if algo == "ippo" else D2DPPO(env, beta_entropy=0.01, **kw)
    for i in range(m["N"]):
        ev.policies.load_state_dict(i, params_from(g, f"final/policy{i}"))
    res = ev.test(Et)
    mine = ev._eval_storage[1]                                                                    # [T, N, Et]
    assert torch.equal(mine, _device_actions(g["test_actions"], g, Et, ev, dev)), (algo, tag, "greedy actions")
    assert np.allclose(res, g["test_result"], rtol=1e-12, atol=1e-12), (res, g["test_result"])
    assert isinstance(res[2], int)


def _adam_tol(tag):
    return dict(min_tight=0.99, loose_frac=2e-3) if "e256" in tag else {}


IPPO_CASES = ["small_gru", "small_mlp", "c3_gru", "d2denv_gru", "selenv_gru", "c3_e256"]
D2DPPO_CASES = ["small_gru", "small_mlp", "c3_gru", "d2denv_gru", "d2denv_mlp", "selenv_mlp", "c3_e256"]


@pytest.mark.parametrize("tag", IPPO_CASES)
def test_ippo_training_iteration(tag, cuda_device):
    g, m, env, agent, forced = _setup("ippo", tag, cuda_device)
    E, T, N = m["E"], m["T"], m["N"]
    obs, actions, logp, ret, values, adv, scores, dones = agent.create_rollouts(E, forced_actions=forced)
    rows = torch.cat([obs[:T, o:o + d].permute(2, 0, 1).reshape(E * T, d)
                      for o, d in zip(agent.obs_off, agent.obs_dim)], dim=1).cpu().numpy()
    ref_obs = g["obs"].reshape(E * T, -1)
    # env + zero-copy rollout buffer: bit exact (the selection env's 1/count acks are the fp32 rounding of the
    # reference's float64 quotient, as torch.tensor(obs, dtype=torch.float) makes them: ippo.py:297)
    assert np.array_equal(rows, ref_obs)
    assert report_err(f"ippo_{tag}/logp", _rows(logp), g["logp_old"]) < TOL
    assert report_err(f"ippo_{tag}/values", _rows(values), g["values"]) < TOL
    assert report_err(f"ippo_{tag}/advantages", _rows(adv), g["advantages"]) < TOL
    assert report_err(f"ippo_{tag}/returns", _rows(ret), g["returns"]) < TOL
    assert np.allclose(scores.cpu().numpy(), g["scores"], rtol=0, atol=1e-12)
    assert dones == [bool(d) for d in g["dones"][:T]]
    for epoch in range(m["n_epoch"]):
        ploss, vloss = agent.update_epoch()
        assert abs(ploss[-1] - g["policy_loss"][epoch]) <= 1e-4 * max(1.0, abs(g["policy_loss"][epoch]))
        assert abs(vloss[-1] - g["value_loss"][epoch]) <= 1e-4 * max(1.0, abs(g["value_loss"][epoch]))
    for i in range(N):
        assert_params_close(agent.policies.state_dict(i), params_from(g, f"final/policy{i}"),
                            params_from(g, f"init/policy{i}"), f"policy{i}", **_adam_tol(tag))
        assert_params_close(agent.values.state_dict(i), params_from(g, f"final/value{i}"),
                            params_from(g, f"init/value{i}"), f"value{i}", **_adam_tol(tag))
    _check_greedy_test(g, m, agent, "ippo", tag, cuda_device)


@pytest.mark.parametrize("tag", D2DPPO_CASES)
def test_d2dppo_training_iteration(tag, cuda_device):
    g, m, env, agent, forced = _setup("d2dppo", tag, cuda_device)
    E, T, N = m["E"], m["T"], m["N"]
    obs, states, actions, logp, rewards, ret, scores, dones = agent.create_rollouts(E, forced_actions=forced)
    S = g["states"].shape[1]
    mine_states = states.reshape(T, S, E).permute(2, 0, 1).reshape(E * T, S).cpu().numpy()
    assert np.array_equal(mine_states, g["states"])
    assert report_err(f"d2dppo_{tag}/logp", _rows(logp), g["logp_old"]) < TOL
    assert np.array_equal(_rows(rewards).astype(np.float64), g["rewards_mean"])
    assert report_err(f"d2dppo_{tag}/returns", _rows(ret), g["returns"]) < TOL
    assert np.allclose(scores.cpu().numpy(), g["scores"], rtol=0, atol=1e-12)
    for epoch in range(m["n_epoch"]):
        ploss, vloss = agent.update_epoch(cycle=g["cycles"][epoch])
        ref = g["policy_loss"][epoch]
        assert np.allclose(ploss, ref, rtol=1e-4, atol=1e-5), (epoch, ploss, ref)
        assert abs(vloss - g["value_loss"][epoch]) <= 1e-4 * max(1.0, abs(g["value_loss"][epoch]))
    for i in range(N):
        assert_params_close(agent.policies.state_dict(i), params_from(g, f"final/policy{i}"),
                            params_from(g, f"init/policy{i}"), f"policy{i}", **_adam_tol(tag))
    assert_params_close(agent.critic.state_dict(0), params_from(g, "final/critic"), params_from(g, "init/critic"),
                        "critic", **_adam_tol(tag))
    _check_greedy_test(g, m, agent, "d2dppo", tag, cuda_device)


def test_train_test_save_load_api(tmp_path, cuda_device):
    """Reference call signatures: train(...) 4-tuple, test(n) 4-tuple, agent_{i}.pth with the reference's keys."""
    from d2d_ppo_b200 import presets
    from d2d_ppo_b200.algorithms.d2d_ppo import D2DPPO
    from d2d_ppo_b200.algorithms.ippo import iPPO
    from d2d_ppo_b200.envs import CombinatorialEnv
    kw = presets.combinatorial_kwargs("setup_8_channels", load=0.5, episode_length=25)
    env = CombinatorialEnv(n_envs=64, device=cuda_device, seed=3, **kw)
    ppo = D2DPPO(env, hidden_size=64, gamma=0.6, policy_lr=3e-4, value_lr=1e-3, useRNN=True,
                 save_path=str(tmp_path), combinatorial=True, history_len=6, early_stopping=True)
    res = ppo.train(num_iter=2, n_epoch=2, num_episodes=64, test_freq=1)       # keyword order of xp_load.py:106
    scores_episode, score_test_list, policy_loss_list, value_loss_list = res
    assert len(scores_episode) == 128 and len(score_test_list) == 4
    assert len(policy_loss_list) == 4 and len(policy_loss_list[0]) == 6 and len(value_loss_list) == 4
    assert all(np.isfinite(v) for v in value_loss_list)
    files = sorted(os.listdir(tmp_path))
    assert files == [f"agent_{i}.pth" for i in range(6)]
    sd = torch.load(tmp_path / "agent_0.pth")
    assert list(sd) == ["lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
                        "layers.0.weight", "layers.0.bias", "layers.2.weight", "layers.2.bias"]
    assert sd["lstm.weight_ih_l0"].shape == (192, 30) and sd["layers.2.weight"].shape == (8, 64)
    before = ppo.policies.params.clone()
    ppo.policies.params.zero_()
    ppo.load(str(tmp_path))
    assert ppo.policies.params.abs().sum() > 0 and before.shape == ppo.policies.params.shape
    score, jains, errors, avg_reward = ppo.test(100)
    assert 0.0 <= score <= 1.0 and 0.0 < jains <= 1.0 and errors == 0 and avg_reward >= 0
    with pytest.raises(ValueError):
        ppo.train(num_iter=1, num_episodes=10)                                   # lockstep batch is n_envs
    ip = iPPO(env, hidden_size=32, gamma=0.4, policy_lr=3e-4, value_lr=1e-3, useRNN=False, combinatorial=True)
    r = ip.train(1, 2, 64, 100)                                                  # positional order of ippo.py:406
    assert len(r[2]) == 2 and isinstance(r[2][0], float)
    sd = ip.policies.state_dict(0)
    assert list(sd) == ["linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias"]


def test_learning_improves_score(cuda_device):
    """Sanity: a few hundred IPPO updates on an easy config raise the URLLC score of greedy rollouts."""
    from d2d_ppo_b200.algorithms.d2d_ppo import D2DPPO
    from d2d_ppo_b200.envs import CombinatorialEnv
    kw = dict(n_agents=3, n_channels=2, deadlines=np.array([4, 4, 4]), lbdas=np.array([0.3] * 3), period=None,
              arrival_probs=None, offsets=None, episode_length=50, traffic_model="aperiodic", periodic_devices=[],
              homogeneous_size=True, channel_switch=np.zeros((3, 2)))
    env = CombinatorialEnv(n_envs=256, device=cuda_device, seed=1, **kw)
    ppo = D2DPPO(env, hidden_size=32, gamma=0.6, policy_lr=3e-3, value_lr=3e-3, useRNN=False, combinatorial=True,
                 early_stopping=False, seed=4)
    s0 = ppo.test(256)[0]
    ppo.train(num_iter=60, num_episodes=256, n_epoch=4, test_freq=10 ** 9)
    s1 = ppo.test(256)[0]
    assert s1 > s0 + 0.02, (s0, s1)


def test_random_access_baseline(cuda_device):
    from d2d_ppo_b200 import presets
    from d2d_ppo_b200.algorithms.baselines import CombinatorialRandomAccess
    from d2d_ppo_b200.envs import CombinatorialEnv
    kw = presets.combinatorial_kwargs("setup", load=1 / 3, homogeneous_size=False)      # config c1 (16 channels)
    env = CombinatorialEnv(n_envs=512, device=cuda_device, seed=9, **kw)
    gf = CombinatorialRandomAccess(env)
    cv = gf.get_best_transmission_probs(512)
    assert len(cv) == 10 and np.isnan(cv[0]) or cv[0] <= max(cv)      # tp = 0 never transmits
    gf.transmission_prob = gf.transmission_prob_list[int(np.nanargmax(cv))]
    score, jains, chsc, rew = gf.run(1024)
    assert 0.3 < score <= 1.0 and 0 < jains <= 1 and chsc == 1.0 and rew > 0
    # statistical agreement with the CPU oracle driven by numpy draws (same distributions, different streams)
    from _helpers import make_oracle
    from oracle.envs_np import NumpySource
    orc = make_oracle("combinatorial", kw, 512, NumpySource(512, 0))
    rng = np.random.default_rng(1)
    orc.reset()
    done = False
    while not done:
        _, _, _, done, _ = orc.step(rng.binomial(1, gf.transmission_prob, (512, 6, 16)))
    ref = 1 - orc.discarded.sum() / orc.received.sum()
    assert abs(score - ref) < 0.02, (score, ref)


def test_random_access_channel_selection_baseline(cuda_device):
    """RandomAccess (baselines.py:5-45) on ChannelSelectionEnv: idle devices never pick a channel, metrics in range,
    reproducible under torch.manual_seed, and more channels help (fewer collisions)."""
    import torch
    from d2d_ppo_b200.algorithms.baselines import RandomAccess
    from d2d_ppo_b200.envs import ChannelSelectionEnv

    def make(C, seed=3):
        N = 5
        return ChannelSelectionEnv(n_agents=N, n_channels=C, deadlines=np.array([7] * N), lbdas=np.array([0.3] * N),
                                   period=None, arrival_probs=None, offsets=None, episode_length=60,
                                   traffic_model="aperiodic", periodic_devices=[],
                                   channel_switch=np.array([0.1] * (C + 1)), n_envs=256, device=cuda_device, seed=seed)
    env = make(8)
    ra = RandomAccess(env)
    _, state = env.reset()
    a = ra.act(state[:, :35])
    empty = state[:, :35].reshape(256, 5, 7).sum(2) == 0
    assert a.shape == (256, 5) and int(a.max()) <= 8 and bool((a[empty] == 0).all()) and bool((a[~empty] >= 0).all())
    torch.manual_seed(11)
    env.set_episode(0)            # every reset starts a fresh Philox stream: pin the episode to replay it
    r1 = ra.run(256)
    torch.manual_seed(11)
    env.set_episode(0)
    r2 = ra.run(256)
    assert r1 == r2
    assert ra.run(256) != r1      # the next episode draws new traffic
    score, jain, chsc, rew = r1
    assert 0.0 <= score <= 1.0 and 0.0 < jain <= 1.0 + 1e-12 and rew > 0
    torch.manual_seed(11)
    few = RandomAccess(make(2)).run(256)
    assert score > few[0]


def test_reference_default_hidden_size(cuda_device):
    """The reference constructors default to hidden_size=128 (ippo.py:223, d2d_ppo.py:220): MLP and GRU learners run
    a training iteration with it, the rollout log-probs match the oracle, and checkpoints round-trip."""
    import torch
    from d2d_ppo_b200 import presets
    from d2d_ppo_b200.algorithms.d2d_ppo import D2DPPO
    from d2d_ppo_b200.algorithms.ippo import iPPO
    from d2d_ppo_b200.envs import CombinatorialEnv
    from oracle import ppo_torch as P
    kw = presets.combinatorial_kwargs("setup_8_channels", load=0.5, episode_length=12)
    B, T, N, C = 24, 12, 6, 8
    for cls, rnn in ((iPPO, False), (D2DPPO, True)):
        env = CombinatorialEnv(n_envs=B, device=cuda_device, seed=3, **kw)
        ag = cls(env, combinatorial=True, useRNN=rnn, history_len=4, early_stopping=False, seed=1)   # hidden_size=128
        assert ag.hidden_size == 128
        out = ag.create_rollouts(B)
        obs, actions, logp = out[0], out[2] if cls is D2DPPO else out[1], out[3] if cls is D2DPPO else out[2]
        I = 14 + 2 * C
        rows = obs[:T].reshape(T, N, I, B).permute(3, 0, 1, 2).reshape(B * T, N, I).cpu()
        acts = ((actions.long().unsqueeze(-1) >> torch.arange(C, device=cuda_device)) & 1).permute(2, 0, 1, 3)
        acts = acts.reshape(B * T, N, C).float().cpu()
        for i in (0, N - 1):
            sd = ag.policies.state_dict(i)
            if rnn:
                x, valid = P.windows(rows[:, i], T, 4, pad=False)
                probs = P.net_forward(sd, x, "sigmoid", valid)
            else:
                probs = P.net_forward(sd, rows[:, i], "softmax")
            ref, _ = P.logp_entropy(probs, acts[:, i], True)
            assert torch.allclose(logp[:, i].t().reshape(-1).cpu(), ref, rtol=1e-5, atol=2e-6)
        before = ag.policies.params.clone()
        res = ag.train(num_iter=1, n_epoch=2, num_episodes=B, test_freq=10 ** 9)
        losses = np.concatenate([np.ravel(np.asarray(v, dtype=np.float64)) for v in list(res[2]) + list(res[3])])
        assert np.isfinite(losses).all() and not torch.equal(before, ag.policies.params)


@pytest.mark.parametrize("algo", ["ippo", "d2dppo"])
def test_mid_training_test_leaves_rollout_untouched(algo, cuda_device):
    """train() runs test(50) between the epochs of an iteration (d2d_ppo.py:450, ippo.py:428; iteration 0 always
    tests, 0 % test_freq == 0); the reference's test() keeps local lists, so the remaining epochs must still update on
    the TRAINING rollout: every rollout buffer an epoch reads is bit-identical before and after test()."""
    from d2d_ppo_b200 import presets
    from d2d_ppo_b200.algorithms.d2d_ppo import D2DPPO
    from d2d_ppo_b200.algorithms.ippo import iPPO
    from d2d_ppo_b200.envs import CombinatorialEnv
    kw = presets.combinatorial_kwargs("setup_8_channels", load=0.5, episode_length=15)
    env = CombinatorialEnv(n_envs=32, device=cuda_device, seed=3, **kw)
    common = dict(hidden_size=32, gamma=0.6, policy_lr=3e-4, value_lr=1e-3, useRNN=True, combinatorial=True,
                  history_len=4, early_stopping=False, seed=2)
    agent = iPPO(env, **common) if algo == "ippo" else D2DPPO(env, **common)
    agent.create_rollouts(32)
    names = ["obs_buf", "act_buf", "logp_buf", "reward_buf", "ret_buf"] + \
        (["value_buf", "adv_buf"] if algo == "ippo" else ["state_buf"])
    before = {n: getattr(agent, n).clone() for n in names}
    it = agent._iter
    score, jains, errors, rew = agent.test(50)          # 50 > B: two lockstep batches, 50 episodes counted
    assert 0 <= score <= 1 and 0 < jains <= 1 and errors == 0 and rew >= 0
    for n in names:
        assert torch.equal(before[n], getattr(agent, n)), n
    assert agent._iter == it                             # the sampling stream of the next rollout is unaffected
    agent.update_epoch()
    # test() counts exactly num_episodes episodes (the first ones of the lockstep batch)
    s7 = agent.test(7)
    assert 0 <= s7[0] <= 1 and isinstance(s7[2], int)
    # and train() with a test every epoch runs end to end
    if algo == "ippo":
        res = agent.train(num_iter=1, n_epoch=2, num_episodes=32, test_freq=1)
    else:
        res = agent.train(num_iter=1, num_episodes=32, n_epoch=2, test_freq=1)
    assert len(res[1]) == 2
