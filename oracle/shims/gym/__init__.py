"""Minimal stand-in for the ``gym`` package (not installed in this image).

Only used by ``oracle/ref_harness.py`` so that the unmodified reference
modules (``from gym import spaces`` at envs/env.py:2,
envs/combinatorial_env.py:2, envs/channel_selection_env.py:2) import.
"""
from . import spaces  # noqa: F401
