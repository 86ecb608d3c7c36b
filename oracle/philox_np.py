"""Philox4x32-10 and the draw transforms of the CUDA "philox" RNG mode, in numpy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference draws from numpy's global Mersenne Twister
(combinatorial_env.py:68,117,180 ...), which a GPU cannot share; the product's
throughput mode instead uses counter-based Philox4x32-10 (Salmon et al.,
"Parallel random numbers: as easy as 1, 2, 3", SC'11; the Random123 library's
``philox4x32_10``) keyed by (seed, env, timestep, device, block).  This file
restates the generator and the exact integer transforms so that tests can
reproduce the device streams on the CPU, feed them to the reference through
``oracle/ref_harness.py`` and demand bit-exact env state.

Known-answer vectors (Random123 ``kat_vectors``) are checked in
``tests/test_philox.py``.

Counter layout (mirrors ``d2d-ppo_b200/csrc/philox.cuh``):
    ctr = (env_global_index, timestep, device | (purpose << 16), block)
    key = (seed & 0xffffffff, seed >> 32)
purpose: 0 = channel switch, 1 = arrival, 2 = policy (random-access action bits).
A 16-bit lane ``c`` lives in block ``c // 8``, word ``(c % 8) // 2``, half ``c % 2``.
Calls are shared between devices where a device needs less than a call yields (``env_common.cuh``):
* arrivals: the counter's device field is ``k // 4`` and device k takes word ``k % 4`` (``word32``);
* single-channel env switch / policy bits: device field ``k // 8``, device k takes 16-bit lane ``k % 8``
  (``lane16_shared``).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

PURPOSE_SWITCH, PURPOSE_ARRIVAL, PURPOSE_POLICY = 0, 1, 2
ENV_LEVEL_DEVICE = 0xFFFF  # "device" id for env-level draws (ChannelSelectionEnv channels)
POISSON_KMAX = 16


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32-valued arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & MASK for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0  # 32x32 -> 64 fits in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def _key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def lanes16(seed, env, t, device, purpose, n_lanes):
    """``n_lanes`` 16-bit uniforms for every env index in ``env`` -> uint32 array [len(env), n_lanes]."""
    env = np.asarray(env, dtype=np.uint64)
    k0, k1 = _key(seed)
    out = np.empty((env.shape[0], n_lanes), dtype=np.uint32)
    c2 = (int(device) & 0xFFFF) | (int(purpose) << 16)
    for blk in range((n_lanes + 7) // 8):
        w = philox4x32_10(env, t, c2, blk, k0, k1)
        for lane in range(blk * 8, min(n_lanes, blk * 8 + 8)):
            word = w[(lane % 8) // 2]
            out[:, lane] = (word >> np.uint32(16 * (lane & 1))) & np.uint32(0xFFFF)
    return out


def word32(seed, env, t, device, purpose):
    """The 32-bit uniform of ``device``: word ``device % 4`` of the call whose device field is ``device // 4``."""
    k0, k1 = _key(seed)
    c2 = ((int(device) // 4) & 0xFFFF) | (int(purpose) << 16)
    return philox4x32_10(np.asarray(env, dtype=np.uint64), t, c2, 0, k0, k1)[int(device) % 4]


def lane16_shared(seed, env, t, device, purpose):
    """One 16-bit lane per device, eight devices per call (single-channel env): lane ``device % 8`` of the call whose
    device field is ``device // 8``."""
    return lanes16(seed, env, t, int(device) // 8, purpose, 8)[:, int(device) % 8]


# ---- integer thresholds (computed on the host in float64, identically in product and oracle) ----

def thr16(p):
    """Bernoulli(p) on a 16-bit lane: event iff lane < thr16(p)."""
    return int(min(65536, max(0, int(np.floor(float(p) * 65536.0 + 0.5)))))


def thr32(p):
    """Bernoulli(p) on a 32-bit word: event iff word < thr32(p) (compare in 64 bits)."""
    return int(min(1 << 32, max(0, int(np.floor(float(p) * 4294967296.0 + 0.5)))))


def poisson_cdf_thresholds(lam):
    """count = #{m < KMAX : u >= thr[m]},  thr[m] = floor(CDF(m) * 2^32) clamped to 2^32 - 1.

    Inverse-CDF sampling at 2^-32 resolution, truncated at POISSON_KMAX.
    """
    lam = float(lam)
    thr = np.empty(POISSON_KMAX, dtype=np.uint64)
    pmf = np.exp(-lam)
    cdf = pmf
    for m in range(POISSON_KMAX):
        thr[m] = min((1 << 32) - 1, int(np.floor(cdf * 4294967296.0)))
        pmf = pmf * lam / (m + 1)
        cdf = min(1.0, cdf + pmf)
    return thr.astype(np.uint32)


def poisson_from_u32(u, thr):
    u = np.asarray(u, dtype=np.uint32)
    return (u[..., None] >= thr[None, :]).sum(-1).astype(np.int64)


def uniform_choice(seed, env, t, device, n_choices):
    """The fused RandomAccess policy of the selection env: floor(u32 * n_choices / 2^32) with the 32-bit word of
    ``device`` from the policy stream (word ``device % 4`` of the call whose device field is ``device // 4``)."""
    u = word32(seed, env, t, device, PURPOSE_POLICY).astype(np.uint64)
    return ((u * np.uint64(n_choices)) >> np.uint64(32)).astype(np.int64)
