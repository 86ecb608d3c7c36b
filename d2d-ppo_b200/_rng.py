"""Host-side integer tables of the Philox RNG mode (thresholds the step kernels compare against).

Counter layout and lane extraction are documented in csrc/philox.cuh.  All tables are computed in
float64 on the host so that the device does integer compares only.
"""
from __future__ import annotations

import math

import numpy as np

POISSON_KMAX = 16  # D2D_POISSON_KMAX


def bernoulli_thr16(p: float) -> int:
    """Bernoulli(p) on a 16-bit lane u: event iff u < thr."""
    return int(min(65536, max(0, math.floor(float(p) * 65536.0 + 0.5))))


def bernoulli_thr32(p: float) -> int:
    """Bernoulli(p) on a 32-bit word u: event iff u < thr (64-bit compare, thr may be 2^32)."""
    return int(min(1 << 32, max(0, math.floor(float(p) * 4294967296.0 + 0.5))))


def poisson_cdf_table(lam: float) -> np.ndarray:
    """count = #{m < KMAX: u >= thr[m]} with thr[m] = floor(CDF_lam(m) * 2^32) clamped to 2^32 - 1."""
    lam = float(lam)
    out = np.zeros(POISSON_KMAX, dtype=np.uint32)
    pmf = math.exp(-lam)
    cdf = pmf
    for m in range(POISSON_KMAX):
        out[m] = min((1 << 32) - 1, int(math.floor(cdf * 4294967296.0)))
        pmf = pmf * lam / (m + 1)
        cdf = min(1.0, cdf + pmf)
    return out
