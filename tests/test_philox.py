"""Philox4x32-10 restatement against the Random123 known-answer vectors, and product tables == oracle tables."""
import numpy as np

from oracle import philox_np as px


def _hex(t):
    return [int(x) for x in t]


def test_philox_known_answer_vectors():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert _hex(px.philox4x32_10(0, 0, 0, 0, 0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert _hex(px.philox4x32_10(f, f, f, f, f, f)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert _hex(px.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_vectorised_matches_scalar():
    env = np.arange(100, dtype=np.uint64)
    v = px.philox4x32_10(env, 3, 5, 1, 42, 0)
    for i in (0, 17, 99):
        s = px.philox4x32_10(i, 3, 5, 1, 42, 0)
        assert [int(x[i]) for x in v] == _hex(s)


def test_product_tables_equal_oracle_tables():
    from d2d_ppo_b200 import _rng
    for p in [0.0, 1e-9, 0.2, 0.4, 0.5, 0.6, 0.8, 0.999999, 1.0]:
        assert _rng.bernoulli_thr16(p) == px.thr16(p)
        assert _rng.bernoulli_thr32(p) == px.thr32(p)
    for lam in [0.0, 1 / 14, 1 / 3, 0.5, 2 / 3, 0.8, 1.0, 2.5]:
        assert np.array_equal(_rng.poisson_cdf_table(lam), px.poisson_cdf_thresholds(lam))


def test_poisson_transform_statistics():
    lam = 0.8
    thr = px.poisson_cdf_thresholds(lam)
    u = px.word32(7, np.arange(200000), 1, 0, px.PURPOSE_ARRIVAL)
    x = px.poisson_from_u32(u, thr)
    assert abs(x.mean() - lam) < 0.01 and abs(x.var() - lam) < 0.02


def test_lane_statistics():
    lanes = px.lanes16(3, np.arange(100000), 2, 1, px.PURPOSE_SWITCH, 16)
    assert lanes.max() < 65536
    for p in (0.2, 0.8):
        frac = (lanes < px.thr16(p)).mean(0)
        assert np.all(np.abs(frac - p) < 0.01)
