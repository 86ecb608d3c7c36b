"""Shared helpers for the parity tests: load a golden env case, build the oracle / CUDA env for it."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def env_case_names():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "env_*.npz"))
                  if "kat" not in os.path.basename(p))


def load_env_case(name):
    z = np.load(os.path.join(GOLDEN, f"env_{name}.npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["kind"] = str(d["kind"]) if "kind" in d else "combinatorial"
    d["config"] = json.loads(str(d["config"]))
    return d


def np_kwargs(kw):
    out = dict(kw)
    for key in ("deadlines", "lbdas", "arrival_probs", "offsets", "channel_switch"):
        if out.get(key) is not None and not np.isscalar(out[key]):
            out[key] = np.asarray(out[key])
    if isinstance(out.get("period"), list):
        out["period"] = np.asarray(out["period"])
    return out


def make_oracle(kind, kw, n_envs, source):
    from oracle import envs_np
    cls = {"combinatorial": envs_np.CombinatorialOracle, "d2d": envs_np.D2DOracle,
           "channel_selection": envs_np.ChannelSelectionOracle}[kind]
    return cls(n_envs=n_envs, source=source, **np_kwargs(kw))


def make_cuda_env(kind, kw, n_envs, **extra):
    from d2d_ppo_b200.envs import ChannelSelectionEnv, CombinatorialEnv, D2DEnv
    cls = {"combinatorial": CombinatorialEnv, "d2d": D2DEnv, "channel_selection": ChannelSelectionEnv}[kind]
    return cls(n_envs=n_envs, **np_kwargs(kw), **extra)


def cat_obs(obs):
    """list of N [B, I_k] arrays/tensors -> [B, sum I_k] float32 numpy."""
    parts = [o.detach().cpu().numpy() if hasattr(o, "detach") else np.asarray(o) for o in obs]
    return np.concatenate(parts, axis=1).astype(np.float32)


def to_np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)
