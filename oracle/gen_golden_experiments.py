"""Reference results of the baseline experiment scripts, for a STATISTICAL comparison with the batched drivers.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Runs the unmodified reference (`CombinatorialEnv` +
`CombinatorialRandomAccess`, global np.random seeded with 42 as xp_n_agents.py:13-15 does) through the loops of
run_ma_baselines.py:58-74 and xp_n_agents.py:62-140 with fewer episodes than the scripts' defaults, and stores, per
sweep point, the URLLC score of every transmission probability of the cross-validation sweep, the picked probability,
and the final run's 4-tuple plus the per-episode spread (to size the tolerance of the comparison).

    python -m oracle.gen_golden_experiments        # ~3 minutes on one core; writes tests/golden/exp_baselines.json
"""
from __future__ import annotations

import json
import os

import numpy as np

from .gen_golden import GOLDEN
from .ref_harness import import_reference

CV_EPISODES, TEST_EPISODES = 30, 300


def _run_point(env_mod, base_mod, kw):
    env = env_mod.CombinatorialEnv(**kw)
    gf = base_mod.CombinatorialRandomAccess(env)
    cv = gf.get_best_transmission_probs(CV_EPISODES)
    best = int(np.argmax(cv))
    gf.transmission_prob = gf.transmission_prob_list[best]
    # per-episode scores of the final run, for the spread
    per_episode = []
    disc = recv = 0.0
    for _ in range(TEST_EPISODES):
        out = gf.run(1)
        per_episode.append(float(out[0]))
        disc += float(env.discarded_packets.sum())
        recv += float(env.received_packets.sum())
    return {"cv_scores": [float(c) for c in cv], "best_index": best, "best_tp": float(gf.transmission_prob),
            "score": 1.0 - disc / recv, "score_episode_std": float(np.std(per_episode)),
            "episodes": TEST_EPISODES, "cv_episodes": CV_EPISODES}


def main():
    import pickle
    from .ref_harness import REFERENCE_ROOT
    env_mod = import_reference("envs.combinatorial_env")
    base_mod = import_reference("algorithms.baselines")
    np.random.seed(42)
    out = {"n_agents_sweep": {}, "ma_baselines": {},
           "note": "unmodified reference, np.random.seed(42); see oracle/gen_golden_experiments.py"}
    for n in (4, 8, 16):                                   # xp_n_agents.py:62-83 (aperiodic, C = 4, deadlines 7)
        kw = dict(n_agents=n, n_channels=4, deadlines=np.array([7] * n), lbdas=np.array([1 / 14] * n), period=None,
                  arrival_probs=None, offsets=None, episode_length=200, traffic_model="aperiodic",
                  periodic_devices=[], channel_switch=np.ones((n, 4)) * 0.8, verbose=False)
        out["n_agents_sweep"][str(n)] = _run_point(env_mod, base_mod, kw)
        print("n_agents", n, out["n_agents_sweep"][str(n)]["best_tp"], out["n_agents_sweep"][str(n)]["score"], flush=True)
    setup = pickle.load(open(os.path.join(REFERENCE_ROOT, "combinatorial_load", "setup.p"), "rb"))
    for load in (setup["loads_list"][0], setup["loads_list"][2], setup["loads_list"][4]):   # run_ma_baselines.py:53-69
        n = setup["n_agents"]
        kw = dict(n_agents=n, n_channels=setup["n_channels"], deadlines=setup["deadlines"], lbdas=np.array([load] * n),
                  period=np.array([int(1 / load)] * n), arrival_probs=setup["arrival_probs"], offsets=setup["offsets"],
                  episode_length=setup["episode_length"], traffic_model="heterogeneous",
                  periodic_devices=[int(i) for i in setup["periodic_devices"]],      # numpy 2: ndarray != [] (shim 2)
                  channel_switch=setup["channel_switch"], verbose=False)
        out["ma_baselines"][repr(float(load))] = _run_point(env_mod, base_mod, kw)
        print("load", load, out["ma_baselines"][repr(float(load))]["best_tp"],
              out["ma_baselines"][repr(float(load))]["score"], flush=True)
    path = os.path.join(GOLDEN, "exp_baselines.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
