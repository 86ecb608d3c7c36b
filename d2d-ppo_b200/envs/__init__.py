"""Lockstep channel-access environments (reference: envs/*.py), stepped by sm_100a kernels."""
from .combinatorial_env import CombinatorialEnv  # noqa: F401
from .env import D2DEnv  # noqa: F401
from .channel_selection_env import ChannelSelectionEnv  # noqa: F401
