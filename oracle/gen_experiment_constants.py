"""Extract the literal settings of the reference's experiment scripts (paths, sweeps, learner hyper-parameters, train /
test call arguments) with ``ast`` -- the scripts are NOT executed -- into tests/golden/experiment_constants.json.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  ``python -m oracle.gen_experiment_constants`` (needs /root/reference).
The drivers in d2d-ppo_b200/experiments.py are compared with this file in tests/test_presets.py.
"""
from __future__ import annotations

import ast
import json
import os

from .gen_golden import GOLDEN
from .ref_harness import REFERENCE_ROOT

SCRIPTS = ["xp_load", "xp_n_agents", "run_ma_baselines", "xp_gamma", "run_ippo_combinatorial"]


def _value(node, env):
    """Evaluate a literal expression (numbers, lists, f-strings over already known names, + - * /)."""
    if isinstance(node, ast.JoinedStr):
        return "".join(str(_value(v.value, env)) if isinstance(v, ast.FormattedValue) else v.value for v in node.values)
    return eval(compile(ast.Expression(node), "<const>", "eval"), {"__builtins__": {}}, dict(env))   # noqa: S307


def extract(script):
    tree = ast.parse(open(os.path.join(REFERENCE_ROOT, script + ".py")).read())
    env, out = {}, {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            if name in ("xp_name", "output_path", "path", "n_seeds", "n_channels", "n_agents", "load", "gammas",
                        "n_agents_list") and name not in env:
                try:
                    env[name] = _value(node.value, env)
                except Exception:
                    pass
    for k in ("n_seeds", "n_channels", "n_agents", "load", "gammas", "n_agents_list"):
        if k in env:
            out[k] = env[k]
    out["output_path"] = env.get("output_path", env.get("path"))
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute):
            kws = {k.arg: _value(k.value, env) for k in node.keywords if isinstance(k.value, ast.Constant)}
            attr = node.func.attr
            if attr == "train" and "train" not in out:
                out["train"] = kws
            elif attr == "test" and node.args and "test_episodes" not in out and isinstance(node.args[0], ast.Constant):
                out["test_episodes"] = node.args[0].value
            elif attr == "run" and node.args and isinstance(node.args[0], ast.Constant):
                out["test_episodes"] = node.args[0].value
            elif attr == "get_best_transmission_probs" and node.args:
                out["cv_episodes"] = node.args[0].value
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in ("iPPO", "D2DPPO") \
                and "hidden_size" not in out:
            for k in node.keywords:
                if k.arg in ("hidden_size", "gamma", "policy_lr", "value_lr", "history_len") and isinstance(k.value, ast.Constant):
                    out[k.arg] = k.value.value
    return out


def main():
    consts = {s: extract(s) for s in SCRIPTS}
    path = os.path.join(GOLDEN, "experiment_constants.json")
    with open(path, "w") as f:
        json.dump(consts, f, indent=1, sort_keys=True)
    print("wrote", path)
    print(json.dumps(consts, indent=1))


if __name__ == "__main__":
    main()
