"""iRDQN fixtures from the UNMODIFIED reference learner (run via ``python -m oracle.gen_golden_dqn``).

TEST INFRASTRUCTURE ONLY.  Writes tests/golden/dqn_*.npz: one reference ``iRDQN.train(n_episodes=K)``
(algorithms/irdqn.py:222-302) on a replayed CombinatorialEnv -- the actions of every step (teacher-forced in the
tests: ``random`` / ``np.random`` / torch streams cannot be shared with a CUDA kernel), the observations, the
``sample_chunk`` start indices and the loss of every ``train_step``, network and target-network parameters before and
after -- followed by the reference's greedy ``test()`` (:305-353) with the trained networks on fresh replayed
episodes.  Harness-level hooks only, no reference file is edited: the env is the replay wrapper of
``gen_golden_ppo``; ``np.random.randint`` is wrapped by a recorder; ``DQN.train_step`` is wrapped to keep the loss the
reference's ``train`` discards; ``iRDQN.test`` is stubbed DURING ``train`` (it would consume env streams at episode 0);
for the hidden-64 case ``irdqn.RNN`` is wrapped to build its network with ``hidden_size=64`` (the reference exposes no
argument for it: ``DQN`` always builds ``RNN(state_size, action_size)``, hidden 100).
"""
from __future__ import annotations

import functools
import json
import os
import random

import numpy as np
import torch

from .gen_golden import GOLDEN, draw_streams
from .gen_golden_ppo import _EpisodeReplayEnv, _c3, _flat, _sd
from .ref_harness import import_reference


def _case(tag, kw, K, L, start, mb, target_freq, loss, gamma, seed, hidden=100, K_test=3, min_margin=1e-4):
    mod = import_reference("algorithms.irdqn")
    T, N, C = kw["episode_length"], kw["n_agents"], kw["n_channels"]
    rng = np.random.default_rng(seed)
    arr, sw = draw_streams("combinatorial", kw, K + K_test, T, rng)
    env = _EpisodeReplayEnv(kw, arr, sw, "combinatorial")
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    real_rnn = mod.RNN
    if hidden != 100:
        mod.RNN = functools.partial(real_rnn, hidden_size=hidden)
    try:
        agent = mod.iRDQN(env, history_len=L, replay_start_size=start, replay_buffer_size=10 ** 6, gamma=gamma,
                          update_target_frequency=target_freq, minibatch_size=mb, learning_rate=1e-3,
                          update_frequency=1, loss=loss, early_stopping=False)
    finally:
        mod.RNN = real_rnn
    agent.device = torch.device("cpu")
    agent.replay_buffer.device = torch.device("cpu")
    for a in agent.agents:
        a.device = torch.device("cpu")
        a.network.to("cpu"), a.target_network.to("cpu")
    out = {"meta": json.dumps(dict(K=K, T=T, N=N, C=C, L=L, replay_start_size=start, minibatch=mb,
                                   update_target_frequency=target_freq, loss=loss, gamma=gamma, hidden=hidden,
                                   K_test=K_test, lr=1e-3)),
           "config": json.dumps(kw), "arrivals": arr, "switches": sw}
    for i, a in enumerate(agent.agents):
        out.update(_flat(f"init/net{i}", _sd(a.network)))

    starts, losses = [], [[] for _ in range(N)]
    real_randint = np.random.randint

    def randint(low, high=None, size=None, *a, **k):
        r = real_randint(low, high, size, *a, **k)
        if size is not None:                        # sample_chunk's start indices (irdqn.py:25); act() draws scalars
            starts.append(np.asarray(r).copy())
        return r

    for i, a in enumerate(agent.agents):
        orig = a.train_step

        def train_step(transitions, _orig=orig, _i=i):
            l = _orig(transitions)
            losses[_i].append(l)
            return l
        a.train_step = train_step
    real_test = agent.test
    agent.test = lambda n, verbose=False: (0.0, 0.0)
    env.action_log = []
    np.random.randint = randint
    try:
        train_scores, _, _ = agent.train(K, early_stopping=False)
    finally:
        np.random.randint = real_randint
    agent.test = real_test
    acts = np.stack(env.action_log).reshape(K, T, N, C)
    assert (acts.sum(-1) == 1).all()
    out.update(actions=acts.argmax(-1).astype(np.uint8), train_scores=np.asarray(train_scores, dtype=np.float64),
               start_idx=np.stack(starts).astype(np.int32) if starts else np.zeros((0, mb), np.int32),
               losses=np.asarray(losses, dtype=np.float32).T,                    # [train steps, N]
               epsilon=np.asarray([agent.agents[0].epsilon]))
    # the flat transition list the reference sampled from, for the CPU restatement test
    buf = list(agent.replay_buffer.buffer)
    out.update(buf_states=np.stack([b[0].numpy() for b in buf]), buf_next=np.stack([b[3].numpy() for b in buf]),
               buf_rewards=np.stack([np.asarray(b[2]) for b in buf]).astype(np.float32),
               buf_dones=np.asarray([bool(b[4]) for b in buf]))
    for i, a in enumerate(agent.agents):
        out.update(_flat(f"final/net{i}", _sd(a.network)))
        out.update(_flat(f"final/target{i}", _sd(a.target_network)))
    # greedy test() with the trained networks on the next K_test replayed episodes
    margins = []
    for a in agent.agents:
        fwd = a.network.forward

        def forward(x, _fwd=fwd):
            q = _fwd(x)
            top = q.detach().topk(2, dim=1).values
            margins.append(float((top[:, 0] - top[:, 1]).min()))
            return q
        a.network.forward = forward
    env.action_log = []
    res = real_test(K_test)
    tacts = np.stack(env.action_log).reshape(K_test, T, N, C)
    out.update(test_result=np.asarray([float(x) for x in res], dtype=np.float64),
               test_actions=tacts.argmax(-1).astype(np.uint8), test_margin=np.asarray(min(margins)))
    print(f"  {tag}: {len(starts)} train steps, losses[-1] {out['losses'][-1] if len(starts) else None}, "
          f"test -> {res}, smallest greedy margin {min(margins):.3e}")
    if min(margins) <= min_margin:
        return False
    path = os.path.join(GOLDEN, f"dqn_{tag}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    return True


def _clear(tag, *args, seed, **kw):
    while not _case(tag, *args, seed=seed, **kw):
        seed += 100
        print(f"  retrying {tag} with seed {seed}")


def main():
    small = dict(n_agents=3, n_channels=4, deadlines=[3, 5, 3], lbdas=[0.5] * 3, period=[2] * 3,
                 arrival_probs=[0.6, 0.9, 1.0], offsets=[0, 1, 0], episode_length=12, traffic_model="heterogeneous",
                 homogeneous_size=True, periodic_devices=[0], channel_switch=[[0.2, 0.4, 0.6, 0.8]] * 3)
    _clear("small_huber", small, K=8, L=4, start=2, mb=8, target_freq=3, loss="huber", gamma=0.9, seed=31)
    _clear("small_mse", small, K=6, L=3, start=1, mb=8, target_freq=2, loss="mse", gamma=0.6, seed=32)
    _clear("c3_h64", _c3(20), K=6, L=5, start=2, mb=32, target_freq=2, loss="huber", gamma=0.99, seed=33, hidden=64)
    _clear("c3_h100", _c3(20), K=5, L=5, start=2, mb=32, target_freq=2, loss="huber", gamma=0.99, seed=34)


if __name__ == "__main__":
    main()
