"""iRDQN on the CUDA path against the UNMODIFIED reference (tests/golden/dqn_*.npz, written by
oracle/gen_golden_dqn.py) and against the torch restatement: replayed env streams, teacher-forced actions and replay
samples (Python's / numpy's / torch's random streams cannot be shared with a CUDA kernel)."""
import numpy as np
import pytest
import torch

from _helpers import assert_params_close, load_dqn_case, make_cuda_env, params_from, report_err

pytestmark = pytest.mark.gpu
CASES = ["small_huber", "small_mse", "c3_h64", "c3_h100"]


def _agent(g, env, **extra):
    from d2d_ppo_b200.algorithms.irdqn import iRDQN
    m = g["meta"]
    kw = dict(history_len=m["L"], replay_start_size=m["replay_start_size"], replay_buffer_size=10 ** 5,
              gamma=m["gamma"], update_target_frequency=m["update_target_frequency"], minibatch_size=m["minibatch"],
              learning_rate=m["lr"], update_frequency=1, loss=m["loss"], early_stopping=False,
              hidden_size=m["hidden"])
    kw.update(extra)
    return iRDQN(env, **kw)


@pytest.mark.parametrize("tag", CASES)
def test_irdqn_training_matches_reference(tag, cuda_device):
    g = load_dqn_case(tag)
    m = g["meta"]
    K, T, N, L = m["K"], m["T"], m["N"], m["L"]
    env = make_cuda_env("combinatorial", g["config"], 1, rng="replay", device=cuda_device)
    agent = _agent(g, env)
    for i in range(N):
        agent.network.load_state_dict(i, params_from(g, f"init/net{i}"))
    agent.sync_target()
    agent.test = lambda *a, **k: (0.0, 0.0)          # the fixture kept test() out of train() as well
    acts = torch.tensor(g["actions"]).to(cuda_device)                       # [K, T, N]
    steps = iter(range(len(g["start_idx"])))

    def forced_actions(ep):
        env.set_replay(g["arrivals"][:, ep:ep + 1], g["switches"][:, ep:ep + 1])
        return acts[ep].unsqueeze(-1).contiguous()                            # [T, N, 1]

    def forced_samples(ep):
        return g["start_idx"][next(steps)], np.zeros(m["minibatch"], dtype=np.int32)

    scores, _, _ = agent.train(K, early_stopping=False, forced_actions=forced_actions, forced_samples=forced_samples)
    assert np.allclose(scores, g["train_scores"], rtol=0, atol=1e-12)
    # the replay ring holds the reference's transitions bit for bit (env + zero-copy rollout + ring copy)
    rb = agent.replay_buffer
    assert len(rb) == K * T and rb.oldest_slot == 0
    I = agent.obs_dim[0]
    ring = rb.obs[:K, :, :, 0].cpu().numpy().reshape(K, T + 1, N, I)
    assert np.array_equal(ring[:, :T].reshape(K * T, N, I), g["buf_states"])
    assert np.array_equal(ring[:, 1:].reshape(K * T, N, I), g["buf_next"])
    assert np.array_equal(rb.rew[:K, :, 0].cpu().numpy().reshape(-1), g["buf_rewards"][:, 0].astype(np.int32))
    # every train_step's loss, all agents
    mine = torch.stack(agent.losses).cpu().numpy()
    assert mine.shape == g["losses"].shape
    assert report_err(f"dqn_{tag}/losses", mine, g["losses"]) < 1e-5
    assert abs(agent.epsilon - float(g["epsilon"][0])) < 1e-12
    for i in range(N):
        init = params_from(g, f"init/net{i}")
        assert_params_close(agent.network.state_dict(i), params_from(g, f"final/net{i}"), init, f"{tag}/net{i}",
                            min_tight=0.99, loose_frac=2e-3)
    tgt = agent.network.params
    agent.network.params = agent.target_params      # read the target parameters through the same layout
    try:
        for i in range(N):
            assert_params_close(agent.network.state_dict(i), params_from(g, f"final/target{i}"),
                                params_from(g, f"init/net{i}"), f"{tag}/target{i}", min_tight=0.99, loose_frac=2e-3)
    finally:
        agent.network.params = tgt


@pytest.mark.parametrize("tag", CASES)
def test_irdqn_greedy_test_matches_reference(tag, cuda_device):
    g = load_dqn_case(tag)
    m = g["meta"]
    K, Kt, N = m["K"], m["K_test"], m["N"]
    env = make_cuda_env("combinatorial", g["config"], Kt, rng="replay", device=cuda_device)
    env.set_replay(g["arrivals"][:, K:K + Kt], g["switches"][:, K:K + Kt])
    agent = _agent(g, env)
    for i in range(N):
        agent.network.load_state_dict(i, params_from(g, f"final/net{i}"))
    res = agent.test(Kt)
    mine = agent.act_buf.permute(2, 0, 1).cpu().numpy()                       # [Kt, T, N]
    assert np.array_equal(mine, g["test_actions"]), (tag, "greedy actions")
    assert np.allclose(res, g["test_result"], rtol=1e-12, atol=1e-12), (res, g["test_result"])
    # the env received the one-hot channel vectors of irdqn.py:326-327
    assert torch.equal(agent.mask_buf.to(torch.int64), torch.ones_like(agent.mask_buf, dtype=torch.int64)
                       << agent.act_buf.to(torch.int64))


@pytest.mark.parametrize("hidden,B", [(64, 256), (100, 8), (32, 12), (100, 6), (128, 16)])
def test_irdqn_lockstep_training_matches_restatement(hidden, B, cuda_device):
    """B lockstep envs on Philox streams, epsilon-greedy actions drawn by the kernel: every train_step (replay gather
    from several env columns and episodes, target network, TD target, loss, BPTT, Adam) against the torch
    restatement evaluated on the agent's own ring."""
    from oracle import irdqn_torch as Q
    g = load_dqn_case("c3_h64")
    m = g["meta"]
    T, N, L, mb = m["T"], m["N"], 4, 16
    env = make_cuda_env("combinatorial", g["config"], B, rng="philox", seed=5, device=cuda_device)
    agent = _agent(g, env, history_len=L, replay_start_size=1, minibatch_size=mb, update_target_frequency=2,
                   hidden_size=hidden, loss="huber", gamma=0.9, seed=3)
    agent.test = lambda *a, **k: (0.0, 0.0)
    rng = np.random.default_rng(7)
    samples = {}

    def forced_samples(ep):
        n = (ep + 1) * T
        samples[ep] = (rng.integers(0, n - L, mb), rng.integers(0, B, mb))
        return samples[ep]

    params = [{k: v.clone() for k, v in agent.network.state_dict(i).items()} for i in range(N)]
    init = [dict(p) for p in params]            # Q.Adam rebinds the entries of the dict it is given
    K = 4
    agent.train(K, early_stopping=False, forced_samples=forced_samples)
    # exploration: episode 0 is not training-ready -> only channels {0, 1}
    rb = agent.replay_buffer
    assert int(rb.act[0].max()) <= 1 and 0.4 < float(rb.act[0].float().mean()) < 0.6
    assert int(rb.act[1:K].max()) <= agent.n_actions - 1
    I = agent.obs_dim[0]
    obs = rb.obs[:K].cpu().numpy()                                            # [K, T + 1, rows, B]
    act = rb.act[:K].cpu().numpy().astype(np.int64)                           # [K, T, N, B]
    rew = rb.rew[:K].cpu().numpy().astype(np.float32)                         # [K, T, B]
    target = [dict(p) for p in params]
    opts = [Q.Adam(p, m["lr"]) for p in params]
    mine = torch.stack(agent.losses).cpu().numpy()
    j = 0
    for ep in range(1, K):
        start, col = samples[ep]
        k = start[:, None] + np.arange(L)[None, :]                            # [mb, L] deque indices
        e, t = k // T, k % T
        c = col[:, None]
        s = obs[e, t, :, c].reshape(mb, L, N, I)
        sn = obs[e, t + 1, :, c].reshape(mb, L, N, I)
        a_last = act[e[:, -1], t[:, -1], :, col]                              # [mb, N]
        r_last = rew[e[:, -1], t[:, -1], col]
        d_last = (t[:, -1] == T - 1).astype(np.float32)
        for i in range(N):
            loss, _ = Q.train_step(opts[i].p, target[i], opts[i], torch.tensor(s[:, :, i]), torch.tensor(a_last[:, i]),
                                   torch.tensor(r_last), torch.tensor(sn[:, :, i]), torch.tensor(d_last), 0.9, "huber")
            assert abs(loss - mine[j, i]) <= 2e-5 * max(1.0, abs(loss)), (hidden, B, ep, i, loss, mine[j, i])
            if ep % 2 == 0:
                target[i] = {k_: v.clone() for k_, v in opts[i].p.items()}
        j += 1
    for i in range(N):
        assert_params_close(agent.network.state_dict(i), opts[i].p, init[i], f"lockstep/net{i}", min_tight=0.99,
                            loose_frac=2e-3)


def test_q_select_semantics(cuda_device):
    from d2d_ppo_b200 import _lib as L
    from d2d_ppo_b200.algorithms._nets import q_select
    N, B, O = 3, 4096, 8
    torch.manual_seed(0)
    q = torch.randn(1, N, O, B, device=cuda_device)
    idx = torch.zeros((N, B), dtype=torch.uint8, device=cuda_device)
    mask = torch.zeros((N, B), dtype=torch.uint8, device=cuda_device)
    q_select(q, N, B, O, L.ACT_GREEDY, 0.5, True, 2, idx, mask)
    assert torch.equal(idx.long(), q[0].argmax(1))
    assert torch.equal(mask.long(), 1 << idx.long())
    q_select(q, N, B, O, L.ACT_SAMPLE, 0.0, True, 2, idx, mask, seed=9)           # epsilon 0: always greedy
    assert torch.equal(idx.long(), q[0].argmax(1))
    q_select(q, N, B, O, L.ACT_SAMPLE, 0.0, False, 2, idx, mask, seed=9)          # not ready: always random in {0, 1}
    assert int(idx.max()) == 1 and abs(float(idx.float().mean()) - 0.5) < 0.02
    q_select(q, N, B, O, L.ACT_SAMPLE, 0.3, True, 2, idx, mask, seed=9, t_abs=4)
    greedy = (idx.long() == q[0].argmax(1)).float().mean().item()                 # 0.7 + 0.3 * P(random == argmax)
    assert 0.7 < greedy < 0.8
    keep = idx.clone()
    mask.zero_()
    q_select(None, N, B, O, L.ACT_GIVEN, 0.0, True, 2, idx, mask)
    assert torch.equal(idx, keep) and torch.equal(mask.long(), 1 << idx.long())


def test_irdqn_checkpoint_roundtrip_and_reference_keys(tmp_path, cuda_device):
    g = load_dqn_case("small_huber")
    env = make_cuda_env("combinatorial", g["config"], 4, rng="philox", device=cuda_device)
    a, b = _agent(g, env, seed=1), _agent(g, env, seed=2)
    assert list(a.network.state_dict(0)) == list(params_from(g, "init/net0"))      # irdqn.RNN's state_dict keys
    for k, v in a.network.state_dict(0).items():
        assert tuple(v.shape) == tuple(params_from(g, "init/net0")[k].shape), k
    a.save(str(tmp_path))
    b.load(str(tmp_path))
    assert torch.equal(a.network.params, b.network.params) and torch.equal(b.target_params, a.network.params)


def test_replay_ring_wraps_like_a_deque(cuda_device):
    """A ring of 3 episode slots fed 5 episodes: deque index 0 is the oldest SURVIVING transition, chunks straddle
    episode ends and the physical end of the ring, done marks the last step of an episode."""
    from d2d_ppo_b200.algorithms.irdqn import ReplayBuffer
    B, N, T, rows, L, mb = 4, 3, 6, 5, 4, 64
    rb = ReplayBuffer(3 * B * T, cuda_device, n_envs=B, n_agents=N, episode_length=T, obs_rows=rows, seed=1)
    assert rb.n_slots == 3
    rng = np.random.default_rng(0)
    eps = []
    for e in range(5):
        obs = rng.standard_normal((T + 1, rows, B)).astype(np.float32)
        act = rng.integers(0, 4, (T, N, B)).astype(np.uint8)
        rew = rng.integers(0, 3, (T, B)).astype(np.int32)
        eps.append((obs, act, rew))
        rb.add_episode(torch.tensor(obs).to(cuda_device), torch.tensor(act).to(cuda_device),
                       torch.tensor(rew).to(cuda_device))
    assert len(rb) == 3 * T and rb.oldest_slot == 2          # episodes 2, 3, 4 survive; episode 2 sits in slot 2
    kept = eps[2:]
    s_flat = np.concatenate([o[:T] for o, _, _ in kept])     # [3T, rows, B] deque order
    n_flat = np.concatenate([o[1:] for o, _, _ in kept])
    a_flat = np.concatenate([a for _, a, _ in kept])
    r_flat = np.concatenate([r for _, _, r in kept])
    start = rng.integers(0, 3 * T - L, mb)
    start[:4] = [T - 2, 2 * T - 1, 3 * T - L - 1, 0]         # straddle both episode ends; the last legal start; the first
    col = rng.integers(0, B, mb)
    xs, act, rew, xn, done = rb.sample_chunk(mb, L, start, col)
    k = start[:, None] + np.arange(L)[None, :]
    assert np.array_equal(xs.cpu().numpy(), s_flat[k, :, col[:, None]].transpose(1, 2, 0))
    assert np.array_equal(xn.cpu().numpy(), n_flat[k, :, col[:, None]].transpose(1, 2, 0))
    last = k[:, -1]
    assert np.array_equal(act.cpu().numpy(), a_flat[last, :, col].T)
    assert np.array_equal(rew.cpu().numpy(), r_flat[last, col])
    assert np.array_equal(done.cpu().numpy(), (last % T == T - 1).astype(np.uint8))
    xs2 = rb.sample_chunk(mb, L)[0]                          # default draws stay in range
    assert xs2.shape == (L, rows, mb)
    with pytest.raises(ValueError):
        rb.sample_chunk(mb, L, start + 3 * T, col)
    rb.reset()
    assert len(rb) == 0
    with pytest.raises(ValueError):
        rb.sample_chunk(mb, L)


@pytest.mark.parametrize("hidden", [16, 32])
def test_irdqn_learning_improves_score(hidden, cuda_device):
    """Sanity beyond parity: two devices on two channels, busy traffic.  The untrained greedy policies of this seed send
    on the same channel (score ~0.4 - 0.5); 300 iterations of independent Q-learning on the shared reward separate them
    (hidden 16: FP32 kernels, hidden 32: tcgen05 window + BPTT)."""
    import contextlib
    import io
    from d2d_ppo_b200.algorithms.irdqn import iRDQN
    from d2d_ppo_b200.envs import CombinatorialEnv
    kw = dict(n_agents=2, n_channels=2, deadlines=np.array([3, 3]), lbdas=np.array([0.6] * 2), period=None,
              arrival_probs=None, offsets=None, episode_length=20, traffic_model="aperiodic", periodic_devices=[],
              homogeneous_size=True, channel_switch=np.zeros((2, 2)))
    env = CombinatorialEnv(n_envs=64, device=cuda_device, seed=2, **kw)
    ag = iRDQN(env, history_len=3, replay_start_size=5, replay_buffer_size=200000, gamma=0.5, update_target_frequency=20,
               minibatch_size=64, learning_rate=1e-3, loss="huber", early_stopping=False, hidden_size=hidden, seed=2)
    s0 = ag.test(256)[0]
    real_test = ag.test
    ag.test = lambda *a, **k: (0.0, 0.0)               # keep the periodic evaluation out of the training run
    with contextlib.redirect_stdout(io.StringIO()):
        scores, _, _ = ag.train(300, early_stopping=False)
    ag.test = real_test
    s1, r1 = ag.test(256)
    assert len(scores) == 300 * 64 and len(ag.losses) == 295
    assert s0 < 0.6 and s1 > s0 + 0.2 and s1 > 0.8, (s0, s1)
    assert r1 > 15.0
