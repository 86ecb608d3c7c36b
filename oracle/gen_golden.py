"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Usage (needs /root/reference):

    python -m oracle.gen_golden            # writes every fixture
    python -m oracle.gen_golden envs       # only the env trajectories

Every fixture stores the inputs (config as JSON, replayed arrival / switch streams, actions) next to the
reference's outputs, so tests can re-run the oracle and the CUDA path on identical inputs without the
reference being present.  The reference has no golden vectors of its own.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

from .ref_harness import RefEnv, reference_available

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# combinatorial_load/setup_8_channels.p and setup.p, as literals; the same
# values live in d2d-ppo_b200/presets.py and are compared with the pickles in tests/test_presets.py.
SETUP8 = dict(
    n_agents=6, n_channels=8, episode_length=200, deadlines=[7, 14, 7, 14, 7, 14],
    arrival_probs=[0.2, 0.4, 0.8, 1.0, 1.0, 1.0], offsets=[0.0] * 6, periodic_devices=[0, 1, 2],
    channel_switch=[[0.4, 0.8, 0.2, 0.4, 0.4, 0.2, 0.4, 0.2], [0.8, 0.2, 0.6, 0.6, 0.6, 0.2, 0.4, 0.2],
                    [0.8, 0.2, 0.4, 0.8, 0.2, 0.2, 0.2, 0.8], [0.4, 0.4, 0.4, 0.4, 0.4, 0.6, 0.2, 0.4],
                    [0.4, 0.4, 0.2, 0.2, 0.2, 0.2, 0.8, 0.6], [0.2, 0.4, 0.4, 0.2, 0.6, 0.6, 0.4, 0.4]])


def env_cases():
    """name -> (kind, constructor kwargs (JSON-able), B, T, action law, arrival scale)."""
    c = {}
    load = 1 / 3
    c["comb_c3_load0.33"] = ("combinatorial", dict(
        n_agents=6, n_channels=8, deadlines=SETUP8["deadlines"], lbdas=[load] * 6, period=[int(1 / load)] * 6,
        arrival_probs=SETUP8["arrival_probs"], offsets=SETUP8["offsets"], episode_length=60,
        traffic_model="heterogeneous", homogeneous_size=True, periodic_devices=SETUP8["periodic_devices"],
        channel_switch=SETUP8["channel_switch"]), 6, 60, 0.3)
    c["comb_c3_load1_ragged_obs"] = ("combinatorial", dict(
        n_agents=6, n_channels=8, deadlines=SETUP8["deadlines"], lbdas=[1.0] * 6, period=[1] * 6,
        arrival_probs=SETUP8["arrival_probs"], offsets=SETUP8["offsets"], episode_length=40,
        traffic_model="heterogeneous", homogeneous_size=False, periodic_devices=SETUP8["periodic_devices"],
        channel_switch=SETUP8["channel_switch"]), 4, 40, 0.5)
    rs = np.random.RandomState(7)
    c["comb_c1_16ch"] = ("combinatorial", dict(
        n_agents=6, n_channels=16, deadlines=SETUP8["deadlines"], lbdas=[0.5] * 6, period=[2] * 6,
        arrival_probs=SETUP8["arrival_probs"], offsets=SETUP8["offsets"], episode_length=40,
        traffic_model="heterogeneous", homogeneous_size=False, periodic_devices=SETUP8["periodic_devices"],
        channel_switch=rs.choice([0.2, 0.4, 0.6, 0.8], size=(6, 16)).tolist()), 4, 40, 0.2)
    c["comb_c4_n12_aperiodic"] = ("combinatorial", dict(
        n_agents=12, n_channels=4, deadlines=[7] * 12, lbdas=[1 / 14] * 12, period=None, arrival_probs=None,
        offsets=None, episode_length=50, traffic_model="aperiodic", periodic_devices=[],
        channel_switch=(np.ones((12, 4)) * 0.8).tolist()), 4, 50, 0.4)
    c["comb_periodic_offsets"] = ("combinatorial", dict(
        n_agents=4, n_channels=3, deadlines=[3, 5, 4, 5], lbdas=[0.5] * 4, period=3,
        arrival_probs=[0.5, 0.7, 1.0, 0.2], offsets=[0, 1, 2, 0], episode_length=30, traffic_model="periodic",
        periodic_devices=[], channel_switch=(np.ones((4, 3)) * 0.3).tolist()), 4, 30, 0.5)
    c["comb_deadline20_c20"] = ("combinatorial", dict(
        n_agents=3, n_channels=20, deadlines=[20, 9, 17], lbdas=[0.9] * 3, period=None, arrival_probs=None,
        offsets=None, episode_length=45, traffic_model="aperiodic", periodic_devices=[],
        channel_switch=rs.uniform(0, 1, size=(3, 20)).round(3).tolist()), 3, 45, 0.15)
    c["d2d_c2"] = ("d2d", dict(
        n_agents=4, deadlines=[7] * 4, lbdas=[1 / 14] * 4, episode_length=80, traffic_model="aperiodic",
        channel_switch=0.2), 6, 80, 0.4)
    c["d2d_neighbourhoods"] = ("d2d", dict(
        n_agents=4, deadlines=[3, 5, 4, 6], lbdas=[0.3] * 4, episode_length=50, traffic_model="aperiodic",
        channel_switch=0.5, neighbourhoods=[[0, 1], [1, 2, 0], [2], [3, 0]]), 4, 50, 0.3)
    c["sel_xp_gamma"] = ("channel_selection", dict(
        n_agents=5, n_channels=16, deadlines=[7] * 5, lbdas=[0.4] * 5, period=[7] * 5, arrival_probs=[1] * 5,
        offsets=[0, 2, 4, 0, 2], episode_length=50, traffic_model="aperiodic", periodic_devices=[2, 4],
        channel_switch=[0.8] * 17), 4, 50, None)
    c["sel_heterogeneous"] = ("channel_selection", dict(
        n_agents=5, n_channels=4, deadlines=[7, 5, 7, 3, 7], lbdas=[0.6] * 5, period=[3] * 5,
        arrival_probs=[0.9] * 5, offsets=[0, 2, 1, 0, 2], episode_length=40, traffic_model="heterogeneous",
        periodic_devices=[2, 4], channel_switch=[0.8, 0.3, 0.5, 0.2, 0.6]), 4, 40, None)
    return c


def to_ref_kwargs(kind, kw):
    out = dict(kw)
    for key in ("deadlines", "lbdas", "arrival_probs", "offsets", "channel_switch"):
        if out.get(key) is not None and not np.isscalar(out[key]):
            out[key] = np.asarray(out[key])
    if isinstance(out.get("period"), list):
        out["period"] = np.asarray(out["period"])
    return out


def draw_streams(kind, kw, B, T, rng):
    N = kw["n_agents"]
    lam = np.asarray(kw["lbdas"], dtype=np.float64)
    arr = rng.poisson(lam[None, None, :], (T + 1, B, N))
    model = kw["traffic_model"]
    if model == "periodic":
        arr = rng.binomial(1, np.asarray(kw["arrival_probs"])[None, None, :], (T + 1, B, N))
    elif model == "heterogeneous":
        for i in kw["periodic_devices"]:
            arr[:, :, i] = rng.binomial(1, kw["arrival_probs"][i], (T + 1, B))
    if kind == "combinatorial":
        p = np.asarray(kw["channel_switch"], dtype=np.float64)
        sw = rng.binomial(1, np.broadcast_to(p, (T + 1, B) + p.shape))
    elif kind == "d2d":
        sw = rng.binomial(1, kw["channel_switch"], (T + 1, B, N))
    else:
        p = np.asarray(kw["channel_switch"], dtype=np.float64)[: kw["n_channels"] + 1]
        sw = rng.binomial(1, np.broadcast_to(p, (T + 1, B) + p.shape))
    return arr.astype(np.uint8), sw.astype(np.uint8)


def draw_actions(kind, kw, B, T, tp, rng):
    N = kw["n_agents"]
    if kind == "combinatorial":
        return rng.binomial(1, tp, (T, B, N, kw["n_channels"])).astype(np.uint8)
    if kind == "d2d":
        return rng.binomial(1, tp, (T, B, N)).astype(np.uint8)
    return rng.integers(0, kw["n_channels"] + 1, (T, B, N)).astype(np.uint8)


def run_reference_env(kind, kw, arr, sw, actions):
    """Returns dict of per-step reference outputs, batched over the B replayed instances."""
    T, B = actions.shape[0], actions.shape[1]
    N = kw["n_agents"]
    refs = [RefEnv(kind, arr[:, b], sw[:, b], **to_ref_kwargs(kind, kw)) for b in range(B)]

    def flat_state(s):
        return np.concatenate(s) if isinstance(s, list) else s

    obs0, state0 = zip(*[r.reset() for r in refs])
    out = {"obs0": np.stack([np.concatenate(o) for o in obs0]).astype(np.float32),
           "state0": np.stack([flat_state(s) for s in state0]).astype(np.float32),
           "buffers0": np.stack([r.env.current_buffers for r in refs]).astype(np.uint8)}
    obs, state, rew, buf, disc, recv, chan, done = [], [], [], [], [], [], [], []
    for t in range(T):
        res = [r.step(actions[t, b].astype(np.float64) if kind != "channel_selection" else actions[t, b])
               for b, r in enumerate(refs)]
        obs.append(np.stack([np.concatenate(x[0]) for x in res]))
        state.append(np.stack([flat_state(x[1]) for x in res]))
        rew.append(np.stack([x[2] for x in res]))
        done.append(res[0][3])
        buf.append(np.stack([r.env.current_buffers for r in refs]))
        disc.append(np.stack([r.env.discarded_packets for r in refs]))
        recv.append(np.stack([r.env.received_packets for r in refs]))
        chan.append(np.stack([np.asarray(r.env.channel_state) for r in refs]))
    out.update(obs=np.stack(obs).astype(np.float32), state=np.stack(state).astype(np.float32),
               rewards=np.stack(rew).astype(np.float32), done=np.array(done), buffers=np.stack(buf).astype(np.uint8),
               discarded=np.stack(disc).astype(np.int32), received=np.stack(recv).astype(np.int32),
               channel=np.stack(chan).astype(np.uint8),
               urllc=np.array([r.env.compute_urllc() for r in refs]),
               jains=np.array([r.env.compute_jains() for r in refs]))
    if kind == "d2d":
        out["channel_errors"] = np.array([r.env.channel_errors for r in refs])
        out["n_collisions"] = np.array([r.env.n_collisions for r in refs])
    if kind != "d2d":
        out["channel_score"] = np.array([r.env.compute_channel_score() for r in refs], dtype=np.float64)
    return out


def gen_envs():
    for i, (name, (kind, kw, B, T, tp)) in enumerate(env_cases().items()):
        rng = np.random.default_rng(1000 + i)
        arr, sw = draw_streams(kind, kw, B, T, rng)
        actions = draw_actions(kind, kw, B, T, tp, rng)
        out = run_reference_env(kind, kw, arr, sw, actions)
        path = os.path.join(GOLDEN, f"env_{name}.npz")
        np.savez_compressed(path, kind=kind, config=json.dumps(kw), arrivals=arr, switches=sw, actions=actions, **out)
        print(f"wrote {path} ({os.path.getsize(path)} bytes)")


def gen_kat():
    """The hand-checked known-answer step of SURVEY.md section 4 (CombinatorialEnv, N=4, C=3)."""
    kw = dict(n_agents=4, n_channels=3, deadlines=[3, 3, 3, 3], lbdas=[0.5] * 4, episode_length=10,
              traffic_model="aperiodic", channel_switch=(np.ones((4, 3)) * 0.5).tolist())
    arr = np.array([[2, 1, 0, 1], [0, 1, 0, 0]], dtype=np.uint8)[:, None, :]          # [T+1=2, B=1, N]
    sw = np.zeros((2, 1, 4, 3), dtype=np.uint8)
    sw[1, 0] = [[0, 0, 1], [0, 0, 0], [0, 0, 0], [1, 0, 0]]
    actions = np.array([[1, 0, 1], [0, 1, 0], [1, 1, 1], [1, 0, 0]], dtype=np.uint8)[None, None]
    ref = RefEnv("combinatorial", arr[:, 0], sw[:, 0], **to_ref_kwargs("combinatorial", kw))
    ref.reset()
    forced = np.array([[1, 1, 0], [1, 1, 1], [1, 1, 1], [0, 1, 1]], dtype=np.float64)
    ref.env.channel_state = forced.copy()
    obs, state, rew, done, _ = ref.step(actions[0, 0].astype(np.float64))
    path = os.path.join(GOLDEN, "env_kat_survey4.npz")
    np.savez_compressed(path, config=json.dumps(kw), arrivals=arr, switches=sw, actions=actions,
                        forced_channel=forced.astype(np.uint8), obs=np.concatenate(obs).astype(np.float32),
                        state=np.concatenate(state).astype(np.float32), rewards=rew.astype(np.float32),
                        buffers=ref.env.current_buffers.astype(np.uint8),
                        discarded=ref.env.discarded_packets.astype(np.int32),
                        received=ref.env.received_packets.astype(np.int32),
                        channel=ref.env.channel_state.astype(np.uint8))
    print(f"wrote {path}")


def main(argv):
    if not reference_available():
        raise SystemExit("the reference tree is not mounted; fixtures can only be generated in the build container")
    os.makedirs(GOLDEN, exist_ok=True)
    what = set(argv) or {"envs", "kat", "ppo"}
    if "envs" in what:
        gen_envs()
    if "kat" in what:
        gen_kat()
    if "ppo" in what:
        from . import gen_golden_ppo
        gen_golden_ppo.main()
    if "ppo_round2" in what:          # only the fixtures added in round 2 (the earlier files stay byte-identical)
        from . import gen_golden_ppo
        gen_golden_ppo.main(("round2",))
    if "ppo_e256" in what:
        from . import gen_golden_ppo
        gen_golden_ppo.main(("e256",))


if __name__ == "__main__":
    main(sys.argv[1:])
