"""Load-file presets: values equal the reference's pickles (when mounted) and round-trip in the reference's format."""
import os
import pickle

import numpy as np
import pytest

from d2d_ppo_b200 import presets

REF = "/root/reference/combinatorial_load"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("name", ["setup", "setup_8_channels"])
def test_presets_equal_reference_load_files(name):
    ref = presets.load_setup(os.path.join(REF, name + ".p"))
    mine = presets.SETUPS[name]
    assert set(ref) == set(mine)
    for k in ref:
        assert np.array_equal(np.asarray(ref[k]), np.asarray(mine[k])), (name, k)
    assert np.array_equal(presets.load_setup(os.path.join(REF, "channel_switch_8.p")), presets.CHANNEL_SWITCH_8)


def test_load_file_round_trip(tmp_path):
    presets.write_load_files(tmp_path)
    s = presets.load_setup(tmp_path / "setup_8_channels.p")
    assert s["n_agents"] == 6 and s["n_channels"] == 8 and s["deadlines"].tolist() == [7, 14, 7, 14, 7, 14]
    assert pickle.load(open(tmp_path / "channel_switch_8.p", "rb")).shape == (6, 8)
    kw = presets.combinatorial_kwargs(s, load=0.5)
    assert kw["period"].tolist() == [2] * 6 and kw["traffic_model"] == "heterogeneous"


def test_experiment_drivers_reproduce_the_reference_scripts_settings():
    """d2d_ppo_b200/experiments.py against the literal settings of the reference's experiment scripts
    (tests/golden/experiment_constants.json, parsed from xp_load.py / xp_n_agents.py / run_ma_baselines.py / xp_gamma.py /
    run_ippo_combinatorial.py by oracle/gen_experiment_constants.py without executing them): output paths, sweeps,
    learner hyper-parameters, train() / test() / run() arguments -- both the table in experiments.DRIVER_CONSTANTS and
    what the drivers actually default to."""
    import json
    import os
    from d2d_ppo_b200 import experiments as X
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "experiment_constants.json")
    ref = json.load(open(path))
    assert set(ref) == set(X.DRIVER_CONSTANTS)

    def close(a, b):
        if isinstance(a, dict):
            return set(a) == set(b) and all(close(a[k], b[k]) for k in a)
        if isinstance(a, (list, tuple)):
            return len(a) == len(b) and all(close(x, y) for x, y in zip(a, b))
        if isinstance(a, float) or isinstance(b, float):
            return abs(float(a) - float(b)) <= 1e-12 * max(1.0, abs(float(b)))
        return a == b
    for name, consts in ref.items():
        mine = X.DRIVER_CONSTANTS[name]
        for key, val in consts.items():
            if key == "n_agents" and key not in mine:
                continue                                    # taken from the load file (setup*.p), checked above
            assert key in mine and close(mine[key], val), (name, key, mine.get(key), val)
        live = X.driver_defaults(name)
        for key, val in live.items():
            if key in consts:
                assert close(val, consts[key]), (name, "default", key, val, consts[key])
    # when the reference tree is present, the committed constants are re-derived from it
    from oracle import gen_experiment_constants as G
    from oracle.ref_harness import reference_available
    if reference_available():
        for name in ref:
            assert close(G.extract(name), ref[name]), name
