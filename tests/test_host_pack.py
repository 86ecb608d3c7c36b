"""Host-side action packing of the host-buffer step (csrc/host_pack.cpp): u8 [B][N][C] -> channel bitmasks [N][B].
No GPU involved: runs in the CPU suite."""
import ctypes as C

import numpy as np
import pytest


def _pack(actions, threads=None):
    from d2d_ppo_b200 import _lib as L
    lib = L.lib()
    if threads is not None:
        L.check(lib.d2d_set_host_threads(threads))
    B, N, Cn = actions.shape
    dt = np.uint8 if Cn <= 8 else (np.uint16 if Cn <= 16 else np.uint32)
    out = np.zeros((N, B), dtype=dt)
    a = np.ascontiguousarray(actions, dtype=np.uint8)
    L.check(lib.d2d_pack_actions_host(C.c_void_p(a.ctypes.data), C.c_void_p(out.ctypes.data), B, N, Cn))
    return out


def _expected(actions):
    B, N, Cn = actions.shape
    w = (1 << np.arange(Cn, dtype=np.uint64))
    return ((actions != 0).astype(np.uint64) * w).sum(-1).T


@pytest.mark.parametrize("B,N,Cn", [(1, 1, 1), (5, 6, 8), (64, 6, 8), (1000, 3, 8), (4099, 6, 8), (257, 4, 4),
                                    (300, 5, 16), (129, 2, 20), (70000, 6, 8)])
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_host_pack_matches_numpy(B, N, Cn, threads):
    rng = np.random.default_rng(B * 131 + N * 7 + Cn)
    actions = (rng.random((B, N, Cn)) < 0.4).astype(np.uint8)
    actions[rng.random((B, N, Cn)) < 0.05] = 7            # any non-zero byte is a "transmit"
    out = _pack(actions, threads)
    assert np.array_equal(out.astype(np.uint64), _expected(actions))


def test_host_threads_setting():
    from d2d_ppo_b200 import _lib as L
    lib = L.lib()
    saved = lib.d2d_get_host_threads()
    L.check(lib.d2d_set_host_threads(5))
    assert lib.d2d_get_host_threads() == 5
    L.check(lib.d2d_set_host_threads(0))                  # default: the affinity mask, at most 16
    assert 1 <= lib.d2d_get_host_threads() <= 16
    assert lib.d2d_set_host_threads(-1) != 0
    L.check(lib.d2d_set_host_threads(saved))
