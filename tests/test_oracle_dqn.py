"""The torch / numpy restatement of the reference's independent recurrent DQN (oracle/irdqn_torch.py) reproduces the
fixtures written from the UNMODIFIED reference (oracle/gen_golden_dqn.py): every train_step loss, the parameters after
training, the target-network sync points and the greedy test() trajectory."""
import numpy as np
import pytest
import torch

from _helpers import assert_params_close, load_dqn_case, make_oracle, params_from
from oracle import envs_np, irdqn_torch as Q

CASES = ["small_huber", "small_mse", "c3_h64", "c3_h100"]


@pytest.mark.parametrize("tag", CASES)
def test_restated_training_matches_reference(tag):
    g = load_dqn_case(tag)
    m = g["meta"]
    K, T, N, L = m["K"], m["T"], m["N"], m["L"]
    params = [params_from(g, f"init/net{i}") for i in range(N)]
    target = [dict(p) for p in params]
    opts = [Q.Adam(p, m["lr"]) for p in params]
    acts = g["actions"].reshape(K * T, N).astype(np.int64)
    j = 0
    for ep in range(K):
        if ep < m["replay_start_size"]:
            continue
        n = (ep + 1) * T                                    # transitions in the deque when this step samples
        assert g["start_idx"][j].max() + L <= n
        s, a, r, sn, d = Q.sample_chunk(g["buf_states"][:n], g["buf_next"][:n], acts[:n], g["buf_rewards"][:n],
                                        g["buf_dones"][:n], g["start_idx"][j], L)
        for i in range(N):
            loss, _ = Q.train_step(opts[i].p, target[i], opts[i], torch.tensor(s[:, :, i]), torch.tensor(a[:, i]),
                                   torch.tensor(r[:, i]), torch.tensor(sn[:, :, i]),
                                   torch.tensor(d.astype(np.float32)), m["gamma"], m["loss"])
            assert abs(loss - g["losses"][j, i]) <= 2e-5 * max(1.0, abs(g["losses"][j, i])), (tag, ep, i)
            if ep % m["update_target_frequency"] == 0:
                target[i] = {k: v.clone() for k, v in opts[i].p.items()}
        j += 1
    assert j == len(g["start_idx"])
    for i in range(N):
        init = params_from(g, f"init/net{i}")
        # Adam amplifies rounding noise on entries whose gradient is ~0 (ReLU-dead units): noise-aware criterion
        assert_params_close(opts[i].p, params_from(g, f"final/net{i}"), init, f"{tag}/net{i}", min_tight=0.99,
                            loose_frac=2e-3)
        assert_params_close(target[i], params_from(g, f"final/target{i}"), init, f"{tag}/target{i}", min_tight=0.99,
                            loose_frac=2e-3)
    assert abs(Q.epsilon_at(K - 1) - float(g["epsilon"][0])) < 1e-12


@pytest.mark.parametrize("tag", CASES)
def test_restated_greedy_test_matches_reference(tag):
    g = load_dqn_case(tag)
    m = g["meta"]
    K, Kt, N = m["K"], m["K_test"], m["N"]
    src = envs_np.ReplaySource(g["arrivals"][:, K:K + Kt], g["switches"][:, K:K + Kt])
    env = make_oracle("combinatorial", g["config"], Kt, src)
    params = [params_from(g, f"final/net{i}") for i in range(N)]
    res, acts = Q.greedy_test(env, params, m["L"])
    assert np.array_equal(acts, g["test_actions"])
    assert np.allclose(res, g["test_result"], rtol=1e-12, atol=1e-12)


def test_rollout_matches_oracle_env():
    """The recorded training rollouts are what the numpy env restatement produces under the recorded actions (the
    replay buffer of the fixture holds the observations of every transition)."""
    g = load_dqn_case("small_huber")
    m = g["meta"]
    K, T, N, C = m["K"], m["T"], m["N"], m["C"]
    env = make_oracle("combinatorial", g["config"], K, envs_np.ReplaySource(g["arrivals"][:, :K], g["switches"][:, :K]))
    obs, _ = env.reset()
    cur = np.stack(obs, axis=1)
    for t in range(T):
        a = g["actions"][:, t]                                                   # [K, N]
        onehot = (a[:, :, None] == np.arange(C)[None, None, :]).astype(np.uint8)
        obs, _, reward, done, _ = env.step(onehot)
        nxt = np.stack(obs, axis=1)
        rows = np.arange(K) * T + t
        assert np.array_equal(g["buf_states"][rows], cur)
        assert np.array_equal(g["buf_next"][rows], nxt)
        assert np.array_equal(g["buf_rewards"][rows], np.asarray(reward, dtype=np.float32))
        cur = nxt
    assert np.allclose(env.compute_urllc(), g["train_scores"], atol=1e-12)
