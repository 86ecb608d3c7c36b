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
