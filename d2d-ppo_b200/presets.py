"""The reference's load files and experiment configurations.

``combinatorial_load/setup.p``, ``setup_8_channels.p`` and ``channel_switch_8.p`` are plain pickles of a dict /
ndarray (written by run_ma_baselines.py:21-32).  ``load_setup`` reads that format; the same values are kept here
as literals so that benchmarks and tests do not depend on the reference checkout being present.
"""
from __future__ import annotations

import pickle

import numpy as np

CHANNEL_SWITCH_8 = np.array([
    [0.4, 0.8, 0.2, 0.4, 0.4, 0.2, 0.4, 0.2],
    [0.8, 0.2, 0.6, 0.6, 0.6, 0.2, 0.4, 0.2],
    [0.8, 0.2, 0.4, 0.8, 0.2, 0.2, 0.2, 0.8],
    [0.4, 0.4, 0.4, 0.4, 0.4, 0.6, 0.2, 0.4],
    [0.4, 0.4, 0.2, 0.2, 0.2, 0.2, 0.8, 0.6],
    [0.2, 0.4, 0.4, 0.2, 0.6, 0.6, 0.4, 0.4]])

CHANNEL_SWITCH_16 = np.array([
    [0.6, 0.8, 0.6, 0.6, 0.8, 0.6, 0.2, 0.2, 0.6, 0.8, 0.4, 0.2, 0.4, 0.6, 0.8, 0.2],
    [0.8, 0.6, 0.6, 0.4, 0.4, 0.8, 0.8, 0.2, 0.4, 0.6, 0.6, 0.4, 0.4, 0.6, 0.8, 0.4],
    [0.6, 0.6, 0.8, 0.4, 0.8, 0.2, 0.2, 0.6, 0.4, 0.8, 0.8, 0.6, 0.8, 0.8, 0.2, 0.2],
    [0.8, 0.8, 0.2, 0.4, 0.8, 0.8, 0.6, 0.4, 0.2, 0.8, 0.2, 0.6, 0.4, 0.2, 0.6, 0.8],
    [0.8, 0.6, 0.6, 0.6, 0.2, 0.8, 0.6, 0.4, 0.6, 0.4, 0.4, 0.6, 0.4, 0.8, 0.8, 0.4],
    [0.8, 0.6, 0.8, 0.2, 0.4, 0.4, 0.4, 0.8, 0.4, 0.2, 0.4, 0.6, 0.6, 0.6, 0.8, 0.8]])


def _setup(n_channels, channel_switch):
    return {"n_agents": 6, "n_channels": n_channels, "episode_length": 200,
            "loads_list": [1 / 3, 1 / 2, 1 / 1.5, 1 / 1.25, 1],
            "deadlines": np.array([7, 14] * 3), "arrival_probs": np.array([0.2, 0.4, 0.8, 1, 1, 1]),
            "offsets": np.zeros(6), "periodic_devices": np.array([0, 1, 2]), "channel_switch": channel_switch}


SETUPS = {"setup_8_channels": _setup(8, CHANNEL_SWITCH_8), "setup": _setup(16, CHANNEL_SWITCH_16)}


def load_setup(path):
    """Read a reference load file (dict for setup*.p, ndarray for channel_switch_8.p)."""
    with open(path, "rb") as f:
        return pickle.load(f)


def write_load_files(directory):
    """Write setup.p / setup_8_channels.p / channel_switch_8.p in the reference's format."""
    import os
    os.makedirs(directory, exist_ok=True)
    for name, s in SETUPS.items():
        with open(os.path.join(directory, name + ".p"), "wb") as f:
            pickle.dump(s, f)
    with open(os.path.join(directory, "channel_switch_8.p"), "wb") as f:
        pickle.dump(CHANNEL_SWITCH_8, f)


def combinatorial_kwargs(setup="setup_8_channels", load=1 / 3, episode_length=None, homogeneous_size=True):
    """CombinatorialEnv kwargs as built by xp_load.py:60-75 (c3/c5) and run_ma_baselines.py:58-69 (c1)."""
    s = SETUPS[setup] if isinstance(setup, str) else setup
    n = s["n_agents"]
    return dict(n_agents=n, n_channels=s["n_channels"], deadlines=s["deadlines"], lbdas=np.array([load] * n),
                period=np.array([int(1 / load)] * n), arrival_probs=s["arrival_probs"], offsets=s["offsets"],
                episode_length=s["episode_length"] if episode_length is None else episode_length,
                traffic_model="heterogeneous", homogeneous_size=homogeneous_size,
                periodic_devices=list(s["periodic_devices"]), channel_switch=s["channel_switch"])


def d2d_c2_kwargs(episode_length=200):
    """Config c2: D2DEnv, N = 4, deadlines 7, aperiodic (test.ipynb cells 4-5)."""
    return dict(n_agents=4, deadlines=np.array([7] * 4), lbdas=np.array([1 / 14] * 4), episode_length=episode_length,
                traffic_model="aperiodic", channel_switch=0.2)


def n_agents_sweep_kwargs(n_agents, load=1 / 14, episode_length=200):
    """Config c4: xp_n_agents.py:62-83."""
    return dict(n_agents=n_agents, n_channels=4, deadlines=np.array([7] * n_agents),
                lbdas=np.array([load] * n_agents), period=None, arrival_probs=None, offsets=None,
                episode_length=episode_length, traffic_model="aperiodic", periodic_devices=[],
                channel_switch=np.ones((n_agents, 4)) * 0.8)
