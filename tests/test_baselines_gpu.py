"""EDF / GFAccess on the CUDA D2DEnv against fixtures produced by the reference's own ``act`` methods
(oracle/gen_golden_baselines.py)."""
import json
import os

import numpy as np
import pytest
import torch

from _helpers import GOLDEN, make_cuda_env

pytestmark = pytest.mark.gpu


def _load():
    z = np.load(os.path.join(GOLDEN, "baselines_edf.npz"))
    d = {k: z[k] for k in z.files}
    d["config"] = json.loads(str(d["config"]))
    return d


def test_edf_run_matches_reference(cuda_device):
    from d2d_ppo_b200.algorithms.baselines import EarliestDeadlineFirstScheduler
    g = _load()
    E = g["plain/arrivals"].shape[1]
    env = make_cuda_env("d2d", g["config"], E, rng="replay", device=cuda_device)
    env.set_replay(g["plain/arrivals"], g["plain/switches"])
    edf = EarliestDeadlineFirstScheduler(env)
    res = edf.run(E)
    assert edf.name == "EDF" and isinstance(res[2], int)
    assert np.allclose(res, g["plain/result"], rtol=0, atol=1e-12), (res, g["plain/result"])
    ref_rew, ref_recv, ref_disc, ref_jains, ref_errs = g["plain/per_episode"]
    assert np.array_equal(env.discarded_packets.sum(1).cpu().numpy(), ref_disc)
    assert np.array_equal(env.channel_errors.cpu().numpy(), ref_errs)


@pytest.mark.parametrize("tag,use_channel", [("plain", False), ("channel", True)])
def test_edf_kernel_matches_reference_step_by_step(tag, use_channel, cuda_device):
    """Teacher-forced with the recorded actions (which include the reference's random picks for all-empty envs):
    wherever a device holds a packet (on a good channel, with use_channel) the kernel's grant is the reference's."""
    from d2d_ppo_b200.algorithms.baselines import EarliestDeadlineFirstScheduler
    g = _load()
    E = g[f"{tag}/arrivals"].shape[1]
    env = make_cuda_env("d2d", g["config"], E, rng="replay", device=cuda_device)
    env.set_replay(g[f"{tag}/arrivals"], g[f"{tag}/switches"])
    edf = EarliestDeadlineFirstScheduler(env, use_channel=use_channel)
    env.reset()
    ref = torch.tensor(g[f"{tag}/actions"]).to(cuda_device)               # [E, T, N]
    anyp = torch.tensor(g[f"{tag}/any_packet"]).to(cuda_device)           # [E, T]
    n_random = 0
    for t in range(env.episode_length):
        a = edf.act()                                                     # [N, E]
        assert torch.equal(a.sum(0), torch.ones(E, dtype=a.dtype, device=a.device).to(a.sum(0).dtype))
        sel = anyp[:, t]
        assert torch.equal(a.t()[sel], ref[:, t][sel]), (tag, t)
        n_random += int((~sel).sum())
        env.step(ref[:, t])
    assert n_random > 0
    ref_rew, ref_recv, ref_disc, ref_jains, ref_errs = g[f"{tag}/per_episode"]
    assert np.array_equal(env.channel_errors.cpu().numpy(), ref_errs)


def test_edf_host_act_is_the_reference_function(cuda_device):
    from d2d_ppo_b200.algorithms.baselines import EarliestDeadlineFirstScheduler
    from oracle import baselines_np
    g = _load()
    env = make_cuda_env("d2d", g["config"], 2, device=cuda_device)
    edf = EarliestDeadlineFirstScheduler(env)
    rng = np.random.default_rng(0)
    for _ in range(200):
        buf = rng.integers(0, 3, (4, 7)) * (rng.random((4, 7)) < 0.2)
        if buf.sum() == 0:
            continue
        assert np.array_equal(edf.act(buf), baselines_np.edf_act(buf[None])[0][0])
        assert np.array_equal(edf.preprocess_state(buf) >= 0, buf.sum(1) > 0)


def test_gfaccess_matches_reference_statistics(cuda_device):
    from d2d_ppo_b200.algorithms.baselines import GFAccess
    ref = json.load(open(os.path.join(GOLDEN, "baselines_gf.json")))
    cfg = ref["config"]
    B = 8192
    for tp, r in ref["tp"].items():
        env = make_cuda_env("d2d", cfg, B, rng="philox", seed=3, device=cuda_device)
        score, jains, errors, rewards = GFAccess(env, transmission_prob=float(tp)).run(B)
        assert abs(score - r["score"]) <= 4 * r["score_se"] + 2e-3, (tp, score, r)
        assert abs(jains - r["jains"]) <= 4 * r["jains_se"] + 2e-3, (tp, jains, r)
        assert abs(rewards - r["rewards"]) <= 4 * r["rewards_se"] + 0.05, (tp, rewards, r)
        assert abs(errors / B - r["errors_per_episode"]) <= 4 * r["errors_se"] + 0.02, (tp, errors / B, r)
    with pytest.raises(ValueError):
        GFAccess(env, use_channel=True)


def test_baselines_single_env_mode(cuda_device):
    """Without n_envs the env is the reference's single host-facing env: the schedulers' run() keeps working and returns
    the reference's tuple types."""
    from d2d_ppo_b200.algorithms.baselines import EarliestDeadlineFirstScheduler, GFAccess
    g = _load()
    env = make_cuda_env("d2d", g["config"], None, device=cuda_device, seed=11)
    for pol in (EarliestDeadlineFirstScheduler(env), EarliestDeadlineFirstScheduler(env, use_channel=True),
                GFAccess(env, transmission_prob=0.3)):
        score, jains, errors, rewards = pol.run(3)
        assert 0.0 <= score <= 1.0 and 0.0 < jains <= 1.0 + 1e-12 and isinstance(errors, int) and errors >= 0
        assert np.isfinite(rewards)
    cv = GFAccess(env, transmission_prob_list=[0.1, 0.9]).get_best_transmission_probs(2)
    assert len(cv) == 2 and all(0.0 <= c <= 1.0 for c in cv)
    # EDF beats grant-free access on this load by a wide margin (it never collides)
    envb = make_cuda_env("d2d", g["config"], 2048, device=cuda_device, seed=12)
    assert EarliestDeadlineFirstScheduler(envb).run(2048)[0] > GFAccess(envb, transmission_prob=0.5).run(2048)[0] + 0.1
