"""numpy restatement of the scheduling baselines of the reference (algorithms/baselines.py:48-168).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity pinning: ``edf_act`` against the reference's own
``EarliestDeadlineFirstScheduler.act`` (callable on a buffers array although the class's ``run`` no longer works against
``D2DEnv``, whose state is a flat array since envs/env.py:98-99) through tests/golden/baselines_edf.npz
(oracle/gen_golden_baselines.py).

* ``edf_act``  <- ``preprocess_state`` + ``act`` :55-76, batched over B envs
* ``edf_run``  <- ``run`` :78-111 with ``buffer_state = env.current_buffers`` (what the unpacking at :87 / :98 meant
                 when the env still returned (buffers, channel)); use_channel masks devices whose CURRENT channel is bad
"""
from __future__ import annotations

import numpy as np


def edf_act(buffers, channel=None):
    """buffers [B, N, D] packet counts by slots-to-expiry, channel [B, N] (1 = good) or None.
    Returns (actions [B, N] one-hot uint8 with the all-empty envs left at zero, any_packet [B])."""
    b = np.asarray(buffers).copy()
    if channel is not None:
        b[np.asarray(channel) == 0] = 0
    B, N, D = b.shape
    nz = b != 0
    first = np.where(nz.any(2), nz.argmax(2), D + 1)            # earliest slot holding a packet, D + 1 if none
    any_packet = (first <= D).any(1)
    pick = first.argmin(1)                                      # first device with the smallest slot (numpy argmin)
    actions = np.zeros((B, N), dtype=np.uint8)
    actions[np.arange(B)[any_packet], pick[any_packet]] = 1
    return actions, any_packet


def edf_run(env, use_channel=False):
    """One lockstep batch of episodes on a batched oracle ``D2DOracle`` -> (discarded, received, jains [B],
    channel_errors [B], reward sums [B] over agents and steps, actions [T, B, N], any_packet [T, B])."""
    env.reset()
    done, acts, anyp = False, [], []
    rew = np.zeros(env.B)
    while not done:
        a, ap = edf_act(env.buffers, env.channel_state if use_channel else None)
        _, _, r, done, _ = env.step(a)
        rew += np.asarray(r, dtype=np.float64).sum(1)
        acts.append(a), anyp.append(ap)
    return (float(env.discarded.sum()), float(env.received.sum()), env.compute_jains(), np.asarray(env.channel_errors),
            rew, np.stack(acts), np.stack(anyp))
