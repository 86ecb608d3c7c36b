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


def poisson_tail_mass(lam: float) -> float:
    """P(X >= POISSON_KMAX) for X ~ Poisson(lam): the probability mass the 16-entry table truncates."""
    lam = float(lam)
    pmf = math.exp(-lam)
    cdf = 0.0
    for m in range(POISSON_KMAX):
        cdf += pmf
        pmf = pmf * lam / (m + 1)
    return max(0.0, 1.0 - cdf)


def check_poisson_rate(lam: float, who: str = "lbdas") -> None:
    """The device draws Poisson arrivals from a POISSON_KMAX-entry inverse-CDF table, i.e. truncated at 15 packets
    per slot.  Rates whose truncated mass would bias received_packets against np.random.poisson are rejected
    (> 1e-6, lam above ~3.6) or flagged (> 1e-9, lam above ~2.2); every script of the reference uses lam <= 1."""
    tail = poisson_tail_mass(lam)
    if tail > 1e-6:
        raise ValueError(f"{who}: Poisson rate {lam} puts {tail:.2e} of its mass at >= {POISSON_KMAX} arrivals per "
                         f"slot, which the device sampler truncates; rates up to ~3.6 are supported")
    if tail > 1e-9:
        import warnings
        warnings.warn(f"{who}: Poisson rate {lam}: {tail:.2e} of the mass (>= {POISSON_KMAX} arrivals per slot) is "
                      f"truncated by the device sampler", RuntimeWarning, stacklevel=3)
