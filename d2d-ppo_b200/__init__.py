"""d2d-ppo_b200: B200-native hot path of benrobaglia/D2D-PPO.

Lockstep step/reset of the device-to-device channel-access environments and the PPO / IPPO
rollout-and-update loop, as hand-written sm_100a CUDA kernels behind a C ABI
(``include/d2d_b200.h`` -> ``libd2d_b200.so``), with the reference's Python env / agent API on top.

There is no CPU fallback: importing the env or agent classes loads the CUDA library and raises if it
has not been built (``python __graft_entry__.py`` / ``d2d_ppo_b200._build.build()``).
"""
__version__ = "0.1.0"

from . import spaces  # noqa: F401
