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


# ---------------------------------------------------------------------------------------------
# PPO fixtures
# ---------------------------------------------------------------------------------------------
def load_ppo_case(name):
    z = np.load(os.path.join(GOLDEN, f"ppo_{name}.npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    for key in ("meta", "config"):
        if key in d:
            d[key] = json.loads(str(d[key]))
    return d


def load_dqn_case(name):
    z = np.load(os.path.join(GOLDEN, f"dqn_{name}.npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    for key in ("meta", "config"):
        d[key] = json.loads(str(d[key]))
    return d


def params_from(d, prefix):
    """{'lstm.weight_ih_l0': tensor, ...} for keys 'prefix/...' of a loaded fixture."""
    import torch
    pre = prefix + "/"
    return {k[len(pre):]: torch.tensor(v) for k, v in d.items() if k.startswith(pre)}


def rel_err(a, b):
    """Norm-wise relative error max|a - b| / max|b|: the fp32 tolerance of the north_star (1e-5) is read relative to
    the scale of the tensor, because an fp32 dot product of O(1) terms that cancels to ~0 carries ~1e-7 ABSOLUTE
    noise in any implementation (torch CPU vs numpy already differ that much)."""
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else a
    b = b.detach().cpu().numpy() if hasattr(b, "detach") else b
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-6))


def assert_params_close(mine, ref, init, what="", min_tight=0.999, loose_frac=2e-4):
    """Parameters after a few Adam steps.  Adam's update is lr * m / (sqrt(v) + 1e-8): for an entry whose gradient is
    ~1e-8 (rounding noise around zero) the step is noise-amplified up to +-lr, in ANY two fp32 implementations.
    (The gradients themselves are compared entry by entry in test_learner_gpu.py: 2e-5 norm-wise bound, 1.3e-6
    measured on the tensor-core path, 4e-7 on the FP32 path.)
    So: >= 99.9 % of the entries must agree to 1e-4 relative / 2e-6 absolute; at most 0.02 % of the entries (and at
    least one) may deviate by more than 5 % of the largest parameter movement -- those are entries whose gradient is
    below the rounding noise in one of the steps, so Adam stepped them in opposite directions -- and none by more
    than the largest movement itself.
    The E = 256 fixtures pass min_tight=0.99 / loose_frac=2e-3: with R = 5120 rows the per-entry gradients are ~1/R
    smaller, so more entries sit at |g| ~ Adam's eps = 1e-8 where the step is proportional to g / eps and absolute
    rounding noise of 1e-9 moves it visibly (the torch restatement in oracle/ppo_torch.py deviates from the reference
    by the same amount there: tests/test_oracle_ppo.py)."""
    import torch
    for k, r in ref.items():
        m = mine[k].detach().cpu() if hasattr(mine[k], "detach") else torch.as_tensor(mine[k])
        r = torch.as_tensor(r)
        diff = (m - r).abs()
        tight = diff <= (2e-6 + 1e-4 * r.abs())
        moved = (r - torch.as_tensor(init[k])).abs().max().item()
        n_off = int((~tight).sum().item())        # small tensors: one noise-stepped entry is always allowed
        assert n_off <= max(1, int((1.0 - min_tight) * diff.numel())), (what, k, tight.float().mean().item())
        loose = int((diff > 0.05 * moved + 2e-6).sum().item())
        assert loose <= max(1, int(loose_frac * diff.numel())), (what, k, loose, diff.max().item(), moved)
        assert diff.max().item() <= 1.05 * moved + 2e-6, (what, k, diff.max().item(), moved)


# ---------------------------------------------------------------------------------------------
# element-wise error report ("print the element-wise relative-error percentiles")
# ---------------------------------------------------------------------------------------------
_REPORT = {}


def report_err(what, mine, ref, strict=1e-5):
    """Records (and prints) the ELEMENT-WISE relative error distribution |a - b| / max(|b|, floor) of one compared
    quantity, floor = 1e-3 * max|b| (entries far below the tensor's scale carry only absolute rounding noise), plus
    the fraction of entries that would fail a strict element-wise bound.  Returns the norm-wise error the asserts use.
    The collected report is written to gpurun_out/parity_error_report.json at the end of the session."""
    a = to_np(mine).astype(np.float64).ravel()
    b = to_np(ref).astype(np.float64).ravel()
    scale = max(float(np.max(np.abs(b))), 1e-30) if b.size else 1.0
    el = np.abs(a - b) / np.maximum(np.abs(b), 1e-3 * scale)
    pure = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    rec = {"n": int(b.size), "normwise": float(np.max(np.abs(a - b)) / max(scale, 1e-6)) if b.size else 0.0,
           "p50": float(np.percentile(el, 50)) if b.size else 0.0, "p99": float(np.percentile(el, 99)) if b.size else 0.0,
           "p999": float(np.percentile(el, 99.9)) if b.size else 0.0, "max": float(el.max()) if b.size else 0.0,
           "frac_over_strict_floored": float((el > strict).mean()) if b.size else 0.0,
           "frac_over_strict_pure": float((pure > strict).mean()) if b.size else 0.0}
    _REPORT[what] = rec
    print(f"[parity] {what}: n={rec['n']} normwise={rec['normwise']:.2e} p50={rec['p50']:.2e} p99={rec['p99']:.2e} "
          f"p99.9={rec['p999']:.2e} max={rec['max']:.2e} frac>1e-5 (floored / pure) = "
          f"{rec['frac_over_strict_floored']:.2e} / {rec['frac_over_strict_pure']:.2e}")
    return rec["normwise"]


def dump_report():
    if not _REPORT:
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_error_report.json"), "w") as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass
