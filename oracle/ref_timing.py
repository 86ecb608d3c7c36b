"""Time the UNMODIFIED reference (staged by oracle/make_ref.py) on host cores.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py): the ``kind: "reference"`` CPU baseline of ``bench.py``.
The reference is driven through its own public API, stock code path, global ``np.random`` (no replay shim: this is
a timing run, not a parity run); the only harness pieces are the ``gym.spaces`` stand-in (gym is not installed) and
``periodic_devices`` passed as a list (numpy 2.x, SURVEY.md section 8c shim 2).

* ``random_access_worker``: ``CombinatorialRandomAccess(env, tp).run(n_episodes)`` (algorithms/baselines.py:193-222 on
  envs/combinatorial_env.py), the loop of run_ma_baselines.py:71-74 -- the reference arm of the headline metric.
* ``ippo_iteration``: ``iPPO.create_rollouts`` + ``iPPO.train`` (ippo.py:277-343, 406-441), GRU actor and critic.
* ``irdqn_episodes``: ``iRDQN.train(n_episodes)`` (irdqn.py:222-302) with the settings of xp_load.py:112-126.
"""
from __future__ import annotations

import os
import sys
import time

from . import make_ref

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def _import(modname):
    import importlib
    root = make_ref.DST if make_ref.available() else make_ref.SRC
    for p in (_SHIMS, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module(modname)


def _env(kw):
    import numpy as np
    mod = _import("envs.combinatorial_env")
    kw = dict(kw)
    kw["periodic_devices"] = [int(i) for i in kw.get("periodic_devices", [])]
    for key in ("deadlines", "lbdas", "arrival_probs", "offsets", "channel_switch", "period"):
        if kw.get(key) is not None and not np.isscalar(kw[key]):
            kw[key] = np.asarray(kw[key])
    return mod.CombinatorialEnv(**kw)


def random_access_worker(args):
    """(env kwargs, tp, n_episodes, seed) -> (agent_steps, seconds) of CombinatorialRandomAccess.run on one core."""
    kw, tp, n_episodes, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import numpy as np
    np.random.seed(seed)
    env = _env(kw)
    policy = _import("algorithms.baselines").CombinatorialRandomAccess(env, transmission_prob=tp)
    policy.run(1)                                           # warm-up episode
    t0 = time.perf_counter()
    policy.run(n_episodes)
    dt = time.perf_counter() - t0
    return n_episodes * env.episode_length * env.n_agents, dt


def random_access_throughput(kw, tp, n_episodes, procs):
    """agent-steps/s of `procs` independent single-env reference processes (total work / slowest process)."""
    if procs == 1:
        steps, dt = random_access_worker((kw, tp, n_episodes, 0))
        return steps / dt, dt
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(random_access_worker, [(kw, tp, n_episodes, i) for i in range(procs)])
    dt = max(r[1] for r in res)
    return sum(r[0] for r in res) / dt, dt


def ippo_iteration(kw, num_episodes=1, n_epoch=5, hidden=64, history_len=6, gamma=0.4, threads=None):
    """One unmodified ``iPPO.train(num_iter=1)`` (GRU actor + critic per agent): dict(rollout_s, total_s, agent_steps)."""
    import numpy as np
    import torch
    if threads:
        torch.set_num_threads(threads)
    np.random.seed(0)
    torch.manual_seed(0)
    env = _env(kw)
    ippo = _import("algorithms.ippo")
    agent = ippo.iPPO(env, hidden_size=hidden, gamma=gamma, policy_lr=3e-4, value_lr=1e-3, device="cpu", useRNN=True,
                      combinatorial=True, history_len=history_len, early_stopping=False)
    agent.test = lambda n: (0.0, 0.0, 0, 0.0)               # SPS excludes the periodic evaluation (iteration 0 tests)
    t0 = time.perf_counter()
    agent.create_rollouts(num_episodes)
    rollout_s = time.perf_counter() - t0
    import contextlib
    import io
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):         # the reference prints a line per epoch of iteration 0
        agent.train(num_iter=1, n_epoch=n_epoch, num_episodes=num_episodes, test_freq=10 ** 9)
    total_s = time.perf_counter() - t0
    return {"rollout_s": rollout_s, "total_s": total_s,
            "agent_steps": num_episodes * env.episode_length * env.n_agents, "threads": torch.get_num_threads()}


def irdqn_episodes(kw, n_episodes=3, threads=None):
    """Unmodified ``iRDQN.train(n_episodes)`` with the (commented-out) settings of xp_load.py:112-126, training from
    episode 1 on so that every timed episode but the first includes a minibatch update; the periodic ``test(50)`` is
    stubbed (SPS excludes evaluation).  dict(total_s, agent_steps, threads)."""
    import contextlib
    import io

    import numpy as np
    import torch
    if threads:
        torch.set_num_threads(threads)
    np.random.seed(0)
    torch.manual_seed(0)
    env = _env(kw)
    mod = _import("algorithms.irdqn")
    agent = mod.iRDQN(env, history_len=env.n_agents, replay_start_size=1, replay_buffer_size=100000, gamma=0.4,
                      update_target_frequency=100, minibatch_size=64, learning_rate=1e-4, update_frequency=1,
                      initial_exploration_rate=1, final_exploration_rate=0.1, adam_epsilon=1e-8, loss='huber')
    agent.test = lambda n, verbose=False: (0.0, 0.0)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        agent.train(n_episodes, early_stopping=False)
    total_s = time.perf_counter() - t0
    return {"total_s": total_s, "agent_steps": n_episodes * env.episode_length * env.n_agents,
            "threads": torch.get_num_threads()}
