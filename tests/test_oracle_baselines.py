"""The numpy restatement of the EDF scheduler (oracle/baselines_np.py) against the fixture written from the reference's
own ``EarliestDeadlineFirstScheduler.act`` on the reference ``D2DEnv`` (oracle/gen_golden_baselines.py)."""
import json
import os

import numpy as np

from _helpers import GOLDEN, make_oracle
from oracle import baselines_np, envs_np


def _load():
    z = np.load(os.path.join(GOLDEN, "baselines_edf.npz"))
    d = {k: z[k] for k in z.files}
    d["config"] = json.loads(str(d["config"]))
    return d


def test_edf_run_reproduces_reference():
    g = _load()
    E = g["plain/arrivals"].shape[1]
    env = make_oracle("d2d", g["config"], E, envs_np.ReplaySource(g["plain/arrivals"], g["plain/switches"]))
    disc, recv, jains, errs, rew, acts, anyp = baselines_np.edf_run(env)
    ref_rew, ref_recv, ref_disc, ref_jains, ref_errs = g["plain/per_episode"]
    assert np.array_equal(anyp.T, g["plain/any_packet"])
    ref_acts = g["plain/actions"].transpose(1, 0, 2)                      # [T, E, N]
    assert np.array_equal(acts[anyp], ref_acts[anyp])
    assert np.array_equal(rew, ref_rew) and np.array_equal(errs, ref_errs)
    assert np.array_equal(env.received.sum(1), ref_recv) and np.array_equal(env.discarded.sum(1), ref_disc)
    assert np.allclose(jains, ref_jains, atol=1e-12)
    res = [1 - disc / recv, jains.mean(), errs.sum(), rew.mean()]
    assert np.allclose(res, g["plain/result"], atol=1e-12)


def test_edf_act_with_channel_mask_matches_reference_step_by_step():
    g = _load()
    E = g["channel/arrivals"].shape[1]
    env = make_oracle("d2d", g["config"], E, envs_np.ReplaySource(g["channel/arrivals"], g["channel/switches"]))
    env.reset()
    ref_acts = g["channel/actions"].transpose(1, 0, 2)
    for t in range(env.episode_length):
        a, anyp = baselines_np.edf_act(env.buffers, env.channel_state)
        assert np.array_equal(anyp, g["channel/any_packet"][:, t])
        assert np.array_equal(a[anyp], ref_acts[t][anyp])
        env.step(ref_acts[t])                                             # teacher-forced: random picks included
    ref_rew, ref_recv, ref_disc, ref_jains, ref_errs = g["channel/per_episode"]
    assert np.array_equal(env.channel_errors, ref_errs) and np.array_equal(env.discarded.sum(1), ref_disc)


def test_edf_act_against_live_reference():
    """Where the reference tree is mounted: its own EarliestDeadlineFirstScheduler.act / preprocess_state and
    GFAccess.act on random buffers against the restatement (GFAccess: the empty-buffer mask; its draws are random)."""
    import pytest
    from oracle import ref_harness
    if not ref_harness.reference_available():
        pytest.skip("reference tree not mounted (GPU box / driver container)")
    base = ref_harness.import_reference("algorithms.baselines")

    class _Env:
        n_agents = 5
    edf = base.EarliestDeadlineFirstScheduler(_Env())
    gf = base.GFAccess(_Env(), transmission_prob=1.0)
    rng = np.random.default_rng(1)
    n_checked = 0
    for _ in range(300):
        buf = (rng.integers(0, 3, (5, 6)) * (rng.random((5, 6)) < 0.25)).astype(np.float64)
        mine, anyp = baselines_np.edf_act(buf[None])
        agg = edf.preprocess_state(buf)
        assert np.array_equal(agg >= 0, buf.sum(1) > 0)
        if anyp[0]:
            assert np.array_equal(edf.act(buf), mine[0])
            n_checked += 1
        assert np.array_equal(gf.act(buf) != 0, buf.sum(1) > 0)       # tp = 1: every device with a packet transmits
    assert n_checked > 200
