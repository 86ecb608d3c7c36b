"""The torch restatement of the learner maths reproduces the reference's outputs stored in tests/golden/ppo_*.npz."""
import numpy as np
import pytest
import torch

from _helpers import assert_params_close, load_ppo_case, params_from, rel_err
from oracle import ppo_torch as P

TOL = 1e-5   # north_star: log-probs, values, advantages and losses within 1e-5 relative in fp32


@pytest.mark.parametrize("tag", ["small", "c3", "categorical"])
def test_nets_and_distributions(tag):
    g = load_ppo_case(f"nets_{tag}")
    comb = g["meta"]["combinatorial"]
    for arch in ("mlp", "gru"):
        pol, val = params_from(g, f"{arch}/policy"), params_from(g, f"{arch}/value")
        x = torch.tensor(g["x_gru" if arch == "gru" else "x_mlp"])
        probs = P.net_forward(pol, x, P.policy_out_kind(arch, comb))
        assert rel_err(probs, g[f"{arch}/probs"]) < TOL
        assert rel_err(P.net_forward(val, x, "identity"), g[f"{arch}/value_out"]) < TOL
        logp, ent = P.logp_entropy(torch.tensor(g[f"{arch}/probs"]), torch.tensor(g[f"{arch}/actions"]), comb)
        assert rel_err(logp, g[f"{arch}/logp"]) < TOL and rel_err(ent, g[f"{arch}/entropy"]) < TOL
        # greedy select_action on single rows: probs > .5 (Bernoulli) / argmax (Categorical)
        p8 = probs[:8]
        greedy = (p8 > 0.5).float() if comb else p8.argmax(1)
        assert np.array_equal(greedy.numpy().reshape(8, -1), g[f"{arch}/greedy_actions"].reshape(8, -1))
        lp, en = P.logp_entropy(p8, greedy, comb)
        assert rel_err(lp, g[f"{arch}/greedy_logp"]) < TOL and rel_err(en, g[f"{arch}/greedy_entropy"]) < TOL


def test_gru_window_equals_torch_gru():
    torch.manual_seed(0)
    gru = torch.nn.GRU(7, 12, 1)
    p = {"lstm." + k: v.detach() for k, v in gru.state_dict().items()}
    x = torch.randn(9, 5, 7)
    out, _ = gru(x.permute(1, 0, 2))
    assert torch.allclose(P.gru_window(p, x), out[-1], atol=1e-6)


def test_returns():
    g = load_ppo_case("returns")
    for tag in ("a", "b", "c"):
        T = int(g[f"{tag}/T"])
        rewards, values = g[f"{tag}/rewards"], g[f"{tag}/values"]
        dones = [(i % T) == T - 1 for i in range(len(rewards))]
        for gamma in (0.6, 0.99):
            assert rel_err(P.lambda_returns(rewards, dones, values, gamma, 0.97), g[f"{tag}/g{gamma}/adv"]) < TOL
            assert rel_err(P.discounted_returns(rewards, gamma, dones), g[f"{tag}/g{gamma}/ret"]) < TOL
        assert rel_err(P.lambda_returns(rewards.mean(1), dones, values[:, 0], 0.6, 0.97), g[f"{tag}/adv1d"]) < TOL
    dones = [(i % 4) == 3 for i in range(12)]
    assert np.array_equal(P.lambda_returns(np.ones((12, 2)), dones, np.zeros((12, 2)), 0.0, 0.97).numpy(), g["deg/adv"])
    assert np.array_equal(P.discounted_returns(np.ones((12, 2)), 0.0, dones).numpy(), g["deg/ret"])


def _inputs(g, arch, i, pad):
    m = g["meta"]
    obs = torch.tensor(g["obs"][:, i])
    if arch == "mlp":
        return obs, None
    return P.windows(obs, m["T"], m["L"], pad)


def _adam_tol(tag):
    return dict(min_tight=0.99, loose_frac=2e-3) if "e256" in tag else {}


IPPO_CASES = ["small_gru", "small_mlp", "c3_gru", "d2denv_gru", "selenv_gru", "c3_e256"]
D2DPPO_CASES = ["small_gru", "small_mlp", "c3_gru", "d2denv_gru", "d2denv_mlp", "selenv_mlp", "c3_e256"]


def _replay_oracle(g, first, count):
    """The numpy oracle env replaying episodes [first, first + count) of a learner fixture's streams."""
    from _helpers import make_oracle
    from oracle.envs_np import ReplaySource
    kind = g["meta"].get("kind", "combinatorial")
    src = ReplaySource(g["arrivals"][:, first:first + count], g["switches"][:, first:first + count])
    return make_oracle(kind, g["config"], count, src)


@pytest.mark.parametrize("tag", IPPO_CASES)
def test_ippo_iteration(tag):
    g = load_ppo_case(f"ippo_{tag}")
    m = g["meta"]
    N, arch, comb = m["N"], m["arch"], m.get("combinatorial", True)
    actions = torch.tensor(g["actions"]).float()
    dones = list(g["dones"])
    # rollout quantities: log-probs from UNPADDED windows, values, then lambda-returns / returns per agent column
    values = []
    for i in range(N):
        pol, val = params_from(g, f"init/policy{i}"), params_from(g, f"init/value{i}")
        x, valid = _inputs(g, arch, i, pad=False)
        probs = P.net_forward(pol, x, P.policy_out_kind(arch, comb), valid)
        logp, _ = P.logp_entropy(probs, actions[:, i], comb)
        assert rel_err(logp, g["logp_old"][:, i]) < TOL
        values.append(P.net_forward(val, x, "identity", valid).squeeze(-1))
    values = torch.stack(values, 1)
    assert rel_err(values, g["values"]) < TOL
    # reward = number of successful devices, identical for all agents; recover it from the env replay
    env = _replay_oracle(g, 0, m["E"])
    env.reset()
    acts = g["actions"].reshape((m["E"], m["T"]) + g["actions"].shape[1:])
    rew = np.stack([env.step(acts[:, t])[2] for t in range(m["T"])], axis=1).reshape(m["E"] * m["T"], N)
    adv = P.lambda_returns(rew, dones, g["values"], m["gamma"], 0.97)
    ret = P.discounted_returns(rew, m["gamma"], dones)
    assert rel_err(adv, g["advantages"]) < TOL and rel_err(ret, g["returns"]) < TOL
    # the update: n_epoch x N train steps on PADDED windows
    pols = [params_from(g, f"init/policy{i}") for i in range(N)]
    vals = [params_from(g, f"init/value{i}") for i in range(N)]
    opt_p = [P.Adam(p, m["policy_lr"]) for p in pols]
    opt_v = [P.Adam(p, m["value_lr"]) for p in vals]
    logp_old = torch.tensor(g["logp_old"])
    for epoch in range(m["n_epoch"]):
        for i in range(N):
            x, valid = _inputs(g, arch, i, pad=True)
            pl, vl = P.ippo_train_step(pols[i], vals[i], opt_p[i], opt_v[i], x, valid, actions[:, i], logp_old[:, i],
                                       ret[:, i], adv[:, i], arch, comb)
        assert rel_err(pl, g["policy_loss"][epoch]) < 1e-4 and rel_err(vl, g["value_loss"][epoch]) < 1e-4
    for i in range(N):
        assert_params_close(pols[i], params_from(g, f"final/policy{i}"), params_from(g, f"init/policy{i}"), f"policy{i}", **_adam_tol(tag))
        assert_params_close(vals[i], params_from(g, f"final/value{i}"), params_from(g, f"init/value{i}"), f"value{i}", **_adam_tol(tag))
    _check_greedy_test(g, "ippo")


def _check_greedy_test(g, algo):
    """The reference's test() with the updated policies on the fixture's extra replayed episodes."""
    m = g["meta"]
    if not m.get("E_test"):
        return
    pols = [params_from(g, f"final/policy{i}") for i in range(m["N"])]
    env = _replay_oracle(g, m["E"], m["E_test"])
    acts, res = P.greedy_test(pols, env, m["arch"], m.get("combinatorial", True), m["L"])
    assert np.array_equal(acts.astype(np.uint8), g["test_actions"])
    assert np.allclose(res, g["test_result"], rtol=1e-12, atol=1e-12), (res, g["test_result"])


@pytest.mark.parametrize("tag", D2DPPO_CASES)
def test_d2dppo_iteration(tag):
    g = load_ppo_case(f"d2dppo_{tag}")
    m = g["meta"]
    N, arch, comb = m["N"], m["arch"], m.get("combinatorial", True)
    actions = torch.tensor(g["actions"]).float()
    dones = list(g["dones"])
    logp_old = torch.tensor(g["logp_old"])
    pols = [params_from(g, f"init/policy{i}") for i in range(N)]
    critic = params_from(g, "init/critic")
    for i in range(N):
        x, valid = _inputs(g, arch, i, pad=False)
        logp, _ = P.logp_entropy(P.net_forward(pols[i], x, P.policy_out_kind(arch, comb), valid), actions[:, i], comb)
        assert rel_err(logp, g["logp_old"][:, i]) < TOL
    rew = g["rewards_mean"]
    ret = P.discounted_returns(np.repeat(rew[:, None], N, 1), m["gamma"], dones).mean(1)
    assert rel_err(ret, g["returns"]) < TOL
    opts = [P.Adam(p, m["policy_lr"]) for p in pols]
    opt_c = P.Adam(critic, m["value_lr"])
    xs, valids = zip(*[_inputs(g, arch, i, pad=True) for i in range(N)])
    states = torch.tensor(g["states"])
    for epoch in range(m["n_epoch"]):
        losses, vl = P.d2dppo_epoch(pols, opts, critic, opt_c, xs, valids, states, actions, logp_old, rew, dones,
                                    torch.tensor(g["returns"]), list(g["cycles"][epoch]), arch, comb, m["gamma"], 0.01)
        assert rel_err(losses, g["policy_loss"][epoch]) < 1e-4 and rel_err(vl, g["value_loss"][epoch]) < 1e-4
    for i in range(N):
        assert_params_close(pols[i], params_from(g, f"final/policy{i}"), params_from(g, f"init/policy{i}"), f"policy{i}", **_adam_tol(tag))
    assert_params_close(critic, params_from(g, "final/critic"), params_from(g, "init/critic"), "critic", **_adam_tol(tag))
    _check_greedy_test(g, "d2dppo")
