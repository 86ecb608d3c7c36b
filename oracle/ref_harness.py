"""Drive the UNMODIFIED reference (``/root/reference``) with replayed random streams.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This module works only in
the build container, where ``/root/reference`` is mounted; it never travels to
the GPU box.  ``oracle/gen_golden.py`` uses it to write ``tests/golden/*.npz``
and the ``-m "not gpu"`` tests use it (when the reference is present) to
validate the numpy restatement in ``oracle/envs_np.py`` directly.

Three shims, none of which edits a reference file:

1. ``gym`` is not installed -> ``oracle/shims/gym`` is put on ``sys.path``.
2. numpy 2.x cannot evaluate ``ndarray != []`` (combinatorial_env.py:76,
   env.py:66, channel_selection_env.py:64) -> ``periodic_devices`` is passed as
   a list.
3. The envs draw from the global ``np.random`` inside ``reset``/``step``
   (combinatorial_env.py:68,73,79,83,117,180,185,190,195; env.py:58..179;
   channel_selection_env.py:56..176).  The module attribute ``np`` of each env
   module is replaced by a proxy whose ``.random.poisson`` / ``.random.binomial``
   return pre-drawn values; every other attribute is real numpy.  To know WHICH
   device a scalar draw is for, ``lbdas`` / ``arrival_probs`` / the scalar
   switch probabilities are passed as ``Tagged`` floats that remember their
   index -- the reference only ever indexes those arrays.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("D2D_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "envs", "combinatorial_env.py"))


def import_reference(modname: str):
    """Import ``envs.env`` / ``algorithms.ippo`` ... from the read-only reference tree."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module(modname)


class Tagged(float):
    """A float that remembers which draw it parameterises: ('arr'|'sw', index)."""

    def __new__(cls, value, kind, idx):
        obj = super().__new__(cls, value)
        obj.kind, obj.idx = kind, int(idx)
        return obj


def tagged_array(values, kind):
    out = np.empty(len(values), dtype=object)
    for i, v in enumerate(values):
        out[i] = Tagged(float(v), kind, i)
    return out


class ReplayRandom:
    """Stands in for ``np.random`` inside one reference env module."""

    def __init__(self):
        self.arrivals = None  # [N] values for the current reset()/step()
        self.switch = None    # array for the current step (shape depends on the env)
        self.n_draws = 0

    def poisson(self, lam):
        self.n_draws += 1
        return int(self.arrivals[lam.idx])

    def binomial(self, n, p, size=None):
        assert n == 1
        self.n_draws += 1
        if isinstance(p, Tagged):
            if p.kind == "arr":
                return int(self.arrivals[p.idx])
            return int(self.switch[p.idx])  # ChannelSelectionEnv: one scalar draw per channel
        if isinstance(p, np.ndarray) and p.ndim == 2:
            return np.asarray(self.switch).astype(np.int64)  # CombinatorialEnv.evolve_channel
        if size is not None:
            return np.asarray(self.switch).astype(np.int64)  # D2DEnv.evolve_channel (scalar p, size N)
        # D2DEnv.decode_signal: p is channel_state[idx] in {0., 1.} -> deterministic outcome
        assert float(p) in (0.0, 1.0)
        return int(p)


class _NumpyProxy:
    def __init__(self, rnd):
        self.random = rnd

    def __getattr__(self, name):
        return getattr(np, name)


class RefEnv:
    """One reference env instance driven by replay streams.

    ``arrivals[t, i]`` is what device ``i`` would draw at timestep ``t``
    (``t = 0`` is ``reset``); ``switches[t]`` is the channel-flip draw of the
    step that produces timestep ``t`` (index 0 unused).
    """

    def __init__(self, kind, arrivals, switches, **kw):
        modname, clsname = {
            "combinatorial": ("envs.combinatorial_env", "CombinatorialEnv"),
            "d2d": ("envs.env", "D2DEnv"),
            "channel_selection": ("envs.channel_selection_env", "ChannelSelectionEnv"),
        }[kind]
        mod = import_reference(modname)
        self.rnd = ReplayRandom()
        self._mod, self._saved_np = mod, mod.np
        self.kind = kind
        self.arrivals, self.switches = arrivals, switches
        kw = dict(kw)
        n = kw["n_agents"]
        kw["lbdas"] = tagged_array(kw["lbdas"], "arr")
        if kw.get("arrival_probs") is not None:
            kw["arrival_probs"] = tagged_array(kw["arrival_probs"], "arr")
        if "periodic_devices" in kw:
            kw["periodic_devices"] = [int(i) for i in kw["periodic_devices"]]  # shim 2
        if kind == "channel_selection" and kw.get("channel_switch") is not None:
            kw["channel_switch"] = tagged_array(kw["channel_switch"], "sw")
        assert len(kw["lbdas"]) == n
        self.env = getattr(mod, clsname)(**kw)

    def _patched(self):
        self._mod.np = _NumpyProxy(self.rnd)

    def _restore(self):
        self._mod.np = self._saved_np

    def reset(self):
        self.rnd.arrivals, self.rnd.switch = self.arrivals[0], None
        self._patched()
        try:
            return self.env.reset()
        finally:
            self._restore()

    def step(self, actions):
        t = self.env.timestep + 1
        self.rnd.arrivals, self.rnd.switch = self.arrivals[t], self.switches[t]
        self._patched()
        try:
            return self.env.step(actions)
        finally:
            self._restore()
