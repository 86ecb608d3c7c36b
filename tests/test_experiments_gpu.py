"""The experiment drivers (d2d_ppo_b200/experiments.py) at smoke size: result-dict layout of the reference scripts
(xp_load.py:148-158, xp_n_agents.py:152-165, run_ma_baselines.py:88-95, xp_gamma.py:92-104) and sane metric ranges."""
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check_metrics(scores, jains):
    s, j = np.asarray(scores, dtype=np.float64), np.asarray(jains, dtype=np.float64)
    assert np.all((s >= 0) & (s <= 1)) and np.all((j > 0) & (j <= 1 + 1e-12))


def test_xp_load_smoke(tmp_path, cuda_device):
    from d2d_ppo_b200 import experiments as X
    out = str(tmp_path / "combinatorial_load")
    res = X.xp_load(out_dir=out, loads=[1 / 3, 1], num_iter=1, n_epoch=1, n_envs=8, test_freq=1, test_episodes=8,
                    device=cuda_device)
    assert set(res) == {"scores", "jains", "channel_errors", "average_rewards", "training"}
    assert res["scores"][0].shape == (2,) and len(res["training"][0]) == 2 and len(res["training"][0][0]) == 4
    _check_metrics(res["scores"], res["jains"])
    assert pickle.load(open(os.path.join(out, "results/mcappo_8_channels.p"), "rb"))["scores"][0].shape == (2,)
    assert os.path.exists(os.path.join(out, f"models_mcappo8_seed_0_load_{1 / 3}", "agent_5.pth"))
    res = X.xp_load(out_dir=out, learner="ippo", loads=[0.5], num_iter=1, n_epoch=1, n_envs=8, test_freq=1,
                    test_episodes=8, device=cuda_device)
    _check_metrics(res["scores"], res["jains"])
    # the iRDQN block the script keeps commented out (xp_load.py:111-131)
    res = X.xp_load(out_dir=out, learner="irdqn", loads=[0.5], num_iter=2, n_envs=8, test_episodes=8,
                    device=cuda_device)
    assert 0 <= res["scores"][0][0] <= 1 and res["channel_errors"][0][0] == "" and len(res["training"][0][0]) == 3


def test_baseline_sweeps_smoke(tmp_path, cuda_device):
    from d2d_ppo_b200 import experiments as X
    res = X.xp_n_agents(out_dir=str(tmp_path / "xp_n_agents"), n_agents_list=(4, 8, 64), n_envs=32, cv_episodes=32,
                        test_episodes=32, device=cuda_device)
    assert res["xp_params"] == {"n_agents": [4, 8, 64], "deadlines": 7} and res["scores"][0].shape == (3,)
    _check_metrics(res["scores"], res["jains"])
    assert res["scores"][0][0] > res["scores"][0][2]            # 64 devices on 4 channels collide far more than 4
    out = str(tmp_path / "combinatorial_load")
    res = X.run_ma_baselines(out_dir=out, n_envs=32, cv_episodes=32, test_episodes=32, device=cuda_device)
    assert set(res) == {"gf_scores", "gf_jains", "gf_channel_errors", "gf_average_rewards"}
    assert len(res["gf_scores"][0]) == 5
    _check_metrics(res["gf_scores"], res["gf_jains"])
    assert res["gf_scores"][0][0] > res["gf_scores"][0][-1]      # URLLC score drops as the load grows
    setup = pickle.load(open(os.path.join(out, "setup.p"), "rb"))        # run_ma_baselines.py:34 writes the load file
    assert setup["n_channels"] == 16 and setup["channel_switch"].shape == (6, 16)


def test_ippo_sweeps_smoke(tmp_path, cuda_device):
    from d2d_ppo_b200 import experiments as X
    res = X.xp_gamma(out_dir=str(tmp_path / "xp_gamma"), gammas=(0.5, 0.9), num_iter=1, n_epoch=1, n_envs=8,
                     test_freq=1, test_episodes=8, device=cuda_device)
    assert res["xp_params"]["gammas"] == [0.5, 0.9] and len(res["scores"]) == 2 and len(res["training"]) == 2
    _check_metrics(res["scores"], res["jains"])
    res = X.run_ippo_combinatorial(out_dir=str(tmp_path / "c16"), num_iter=1, n_epoch=1, n_envs=8, test_freq=1,
                                   test_episodes=8, device=cuda_device)
    _check_metrics(res["scores"], res["jains"])


def test_random_access_baseline_matches_reference_results(cuda_device):
    """The baseline experiments of run_ma_baselines.py:58-74 and xp_n_agents.py:62-140, statistically: URLLC score of
    CombinatorialRandomAccess for every transmission probability of the cross-validation sweep and at the reference's
    picked probability, against results of the UNMODIFIED reference (tests/golden/exp_baselines.json, written by
    oracle/gen_golden_experiments.py).  Different random streams, same distributions: the tolerance is 4 standard errors
    of the reference's own estimate (from its per-episode spread) plus 0.01."""
    import json
    from d2d_ppo_b200 import presets
    from d2d_ppo_b200.algorithms.baselines import CombinatorialRandomAccess
    from d2d_ppo_b200.envs import CombinatorialEnv
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exp_baselines.json")
    ref = json.load(open(path))
    B = 2048

    def check(kw, r, what):
        env = CombinatorialEnv(n_envs=B, device=cuda_device, seed=123, **kw)
        gf = CombinatorialRandomAccess(env)
        cv = gf.get_best_transmission_probs(B)
        se_cv = r["score_episode_std"] / np.sqrt(r["cv_episodes"])
        for tp, mine, theirs in zip(gf.transmission_prob_list, cv, r["cv_scores"]):
            assert abs(mine - theirs) <= 4 * se_cv + 0.01, (what, "cv", float(tp), mine, theirs)
        # the reference's pick is (near-)optimal here too
        assert cv[r["best_index"]] >= max(cv) - (4 * se_cv + 0.01), (what, cv, r["best_index"])
        gf.transmission_prob = r["best_tp"]
        score = gf.run(B)[0]
        se = r["score_episode_std"] / np.sqrt(r["episodes"])
        assert abs(score - r["score"]) <= 4 * se + 0.01, (what, score, r["score"])

    for n, r in ref["n_agents_sweep"].items():
        check(presets.n_agents_sweep_kwargs(int(n), load=1 / 14), r, f"xp_n_agents N={n}")
    for load, r in ref["ma_baselines"].items():
        check(presets.combinatorial_kwargs("setup", load=float(load), homogeneous_size=False), r,
              f"run_ma_baselines load={load}")
