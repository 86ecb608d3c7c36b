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
