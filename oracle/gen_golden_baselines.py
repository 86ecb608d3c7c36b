"""Fixtures for the scheduling baselines (run via ``python -m oracle.gen_golden_baselines``).

TEST INFRASTRUCTURE ONLY.  ``EarliestDeadlineFirstScheduler`` and ``GFAccess`` (algorithms/baselines.py:48-168) no
longer run end to end in the reference snapshot: ``run`` unpacks the env's state as a (buffers, channel) pair (:87, :98)
but ``D2DEnv`` returns one flat array (envs/env.py:98-99), and ``GFAccess.run`` reads ``buffer_state`` before assigning
it (:153).  Their ``act`` methods are intact.  The harness therefore drives the UNMODIFIED ``act`` of each class
against the UNMODIFIED ``D2DEnv`` with the loop ``run`` evidently intends -- ``buffer_state = env.current_buffers`` --
and accumulates the episode statistics exactly as :101-111 does.

* baselines_edf.npz: replayed env streams, the action of every step, the 4-tuple.  Without ``use_channel`` EDF is
  deterministic given the streams (the random pick of :73 only happens when no device holds a packet, where the action
  has no effect), so the 4-tuple is exact; with ``use_channel`` the random pick can land on a device with a packet and
  a bad channel, so the tests teacher-force the recorded actions and compare the policy step by step wherever it is
  not the random pick.
* baselines_gf.json: GFAccess score / Jain / reward statistics over many episodes on the global ``np.random`` (compared
  statistically).
"""
from __future__ import annotations

import json
import os

import numpy as np

from .gen_golden import GOLDEN, draw_streams, to_ref_kwargs
from .ref_harness import RefEnv, import_reference

D2D = dict(n_agents=4, deadlines=[7, 5, 7, 3], lbdas=[0.2, 0.3, 0.15, 0.25], episode_length=40,
           traffic_model="aperiodic", channel_switch=0.2)


def _episode(env, policy, use_channel, record):
    _, _ = env.reset()
    e = env.env
    done, rewards, acts, anyp = False, [], [], []
    while not done:
        buffers = np.array(e.current_buffers, dtype=np.float64)
        if use_channel:
            buffers[e.channel_state == 0] = 0                     # baselines.py:92-94 on the CURRENT channel state
        action = policy.act(buffers)
        anyp.append(bool(buffers.sum() > 0))
        acts.append(np.asarray(action).astype(np.uint8))
        _, _, reward, done, _ = env.step(action)
        rewards.append(reward)
    if record is not None:
        record["actions"].append(np.stack(acts)), record["any_packet"].append(np.asarray(anyp))
    return (np.sum(rewards), e.received_packets.sum(), e.discarded_packets.sum(), e.compute_jains(), e.channel_errors)


def gen_edf():
    base = import_reference("algorithms.baselines")
    E, T = 24, D2D["episode_length"]
    np.random.seed(7)
    out = {"config": json.dumps(D2D)}
    for tag, use_channel in (("plain", False), ("channel", True)):
        rng = np.random.default_rng(41 + use_channel)
        arr, sw = draw_streams("d2d", D2D, E, T, rng)
        rec = {"actions": [], "any_packet": []}
        stats = []
        for ep in range(E):
            env = RefEnv("d2d", arr[:, ep], sw[:, ep], **to_ref_kwargs("d2d", D2D))
            pol = base.EarliestDeadlineFirstScheduler(env.env, use_channel=use_channel)
            stats.append(_episode(env, pol, use_channel, rec))
        rew, recv, disc, jains, errs = map(np.asarray, zip(*stats))
        out.update({f"{tag}/arrivals": arr, f"{tag}/switches": sw, f"{tag}/actions": np.stack(rec["actions"]),
                    f"{tag}/any_packet": np.stack(rec["any_packet"]),
                    f"{tag}/result": np.asarray([1 - disc.sum() / recv.sum(), jains.mean(), errs.sum(), rew.mean()],
                                                dtype=np.float64),
                    f"{tag}/per_episode": np.stack([rew, recv, disc, jains, errs]).astype(np.float64)})
        print(tag, out[f"{tag}/result"])
    path = os.path.join(GOLDEN, "baselines_edf.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


def gen_gf(n_episodes=400):
    base = import_reference("algorithms.baselines")
    mod = import_reference("envs.env")
    np.random.seed(5)
    res = {"config": D2D, "n_episodes": n_episodes, "tp": {}}
    for tp in (0.2, 0.5):
        env = mod.D2DEnv(**{**D2D, "deadlines": np.asarray(D2D["deadlines"]), "lbdas": np.asarray(D2D["lbdas"])})
        pol = base.GFAccess(env, transmission_prob=tp)

        class _E:                                    # _episode() expects the harness wrapper's attribute layout
            pass
        w = _E()
        w.env, w.reset, w.step = env, env.reset, env.step
        stats = np.asarray([_episode(w, pol, False, None) for _ in range(n_episodes)], dtype=np.float64)
        rew, recv, disc, jains, errs = stats.T
        per_ep_score = 1 - disc / np.maximum(recv, 1)
        res["tp"][str(tp)] = {"score": float(1 - disc.sum() / recv.sum()),
                              "score_se": float(per_ep_score.std(ddof=1) / np.sqrt(n_episodes)),
                              "jains": float(jains.mean()), "jains_se": float(jains.std(ddof=1) / np.sqrt(n_episodes)),
                              "rewards": float(rew.mean()), "rewards_se": float(rew.std(ddof=1) / np.sqrt(n_episodes)),
                              "errors_per_episode": float(errs.mean()),
                              "errors_se": float(errs.std(ddof=1) / np.sqrt(n_episodes))}
        print(tp, res["tp"][str(tp)])
    path = os.path.join(GOLDEN, "baselines_gf.json")
    json.dump(res, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    gen_edf()
    gen_gf()
