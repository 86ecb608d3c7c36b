"""One whole iPPO iteration on the CPU, assembled from the oracle parts (numpy env + torch learner maths).

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py): this is the CPU baseline that ``bench.py`` times next
to the CUDA path for the "env + learned-policy rollout" and "PPO train SPS" figures.  It
follows the reference's control flow (``iPPO.create_rollouts`` ippo.py:277-343, ``iPPO.train`` :406-441:
per-agent GRU actor and critic evaluated on the unpadded window at every env step, lambda-returns and discounted
returns per agent column, then ``n_epoch`` x N ``train_step`` calls on freshly built zero-padded windows), but it
is vectorised over ``n_envs`` lockstep envs, which the single-env reference is not: it is a generous baseline.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import ppo_torch as P
from .envs_np import CombinatorialOracle, NumpySource


def ippo_iteration_cpu(n_envs, env_kwargs, n_epoch=5, hidden=64, history_len=6, gamma=0.4, policy_lr=3e-4,
                       value_lr=1e-3, seed=0):
    """Returns dict(rollout_s, update_s, agent_steps) for one iteration (rollout of n_envs episodes + n_epoch epochs)."""
    E = int(n_envs)
    env = CombinatorialOracle(n_envs=E, source=NumpySource(E, seed), **env_kwargs)
    N, C, T = env.n_agents, env.n_channels, env.episode_length
    gen = torch.Generator().manual_seed(seed)
    obs, _ = env.reset()
    dims = [o.shape[1] for o in obs]
    pols = [P.init_params("gru", dims[i], C, hidden, gen) for i in range(N)]
    vals = [P.init_params("gru", dims[i], 1, hidden, gen) for i in range(N)]
    opt_p = [P.Adam(p, policy_lr) for p in pols]
    opt_v = [P.Adam(p, value_lr) for p in vals]
    Lh = history_len

    t0 = time.perf_counter()
    hist = [[torch.tensor(obs[i], dtype=torch.float32)] for i in range(N)]
    actions = torch.zeros(T, E, N, C)
    logps = torch.zeros(T, E, N)
    values = torch.zeros(T, E, N)
    rewards = np.zeros((T, E, N))
    with torch.no_grad():
        for t in range(T):
            for i in range(N):
                x = torch.stack(hist[i][-Lh:], dim=1)                              # unpadded window (ippo.py:300)
                probs = P.net_forward(pols[i], x, "sigmoid")
                a = torch.bernoulli(probs, generator=gen)
                lp, _ = P.logp_entropy(probs, a, True)
                actions[t, :, i], logps[t, :, i] = a, lp
                values[t, :, i] = P.net_forward(vals[i], x, "identity").squeeze(-1)
            obs, _, rew, done, _ = env.step(actions[t].numpy())
            rewards[t] = rew
            for i in range(N):
                hist[i].append(torch.tensor(obs[i], dtype=torch.float32))
    # episode-major rows (row = e * T + t), as the reference concatenates episodes
    def rows(a):
        return a.transpose(0, 1).reshape((E * T,) + tuple(a.shape[2:]))
    dones = [(r % T) == T - 1 for r in range(E * T)]
    rew_rows = rewards.transpose(1, 0, 2).reshape(E * T, N)
    adv = P.lambda_returns(rew_rows, dones, rows(values).numpy(), gamma, 0.97)
    ret = P.discounted_returns(rew_rows, gamma, dones)
    rollout_s = time.perf_counter() - t0

    t1 = time.perf_counter()
    obs_rows = [rows(torch.stack(hist[i][:T], dim=0)) for i in range(N)]
    act_rows, lp_rows = rows(actions), rows(logps)
    for _ in range(n_epoch):
        for i in range(N):
            x, valid = P.windows(obs_rows[i], T, Lh, pad=True)                     # rebuilt every epoch (ippo.py:419)
            P.ippo_train_step(pols[i], vals[i], opt_p[i], opt_v[i], x, valid, act_rows[:, i], lp_rows[:, i],
                              ret[:, i], adv[:, i], "gru", True)
    update_s = time.perf_counter() - t1
    return {"rollout_s": rollout_s, "update_s": update_s, "agent_steps": E * T * N}
