"""CPU restatement (torch fp32 / numpy) of the independent recurrent DQN of the reference (algorithms/irdqn.py).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity pinning: checked against the unmodified reference through
``tests/golden/dqn_*.npz`` written by ``oracle/gen_golden_dqn.py`` (tests/test_oracle_dqn.py).

Restated (file:line under /root/reference/algorithms/irdqn.py):
* ``q_forward``        <- ``RNN.forward`` :58-86 (GRU gates written out, then Linear-ReLU-Linear-ReLU-Linear)
* ``sample_chunk``     <- ``ReplayBuffer.sample_chunk`` :24-42 on a flat list of transitions (chunks may straddle an
                          episode end; only the last transition's action / reward / done are used, :293-296)
* ``train_step``       <- ``DQN.train_step`` :133-148 (TD target from the target network, Huber / MSE, Adam)
* ``epsilon_at``       <- ``DQN.update_epsilon`` :159-161
* ``greedy_test``      <- ``iRDQN.test`` :305-353 (greedy episodes, (1 - sum discarded / sum received, mean score))
"""
from __future__ import annotations

import numpy as np
import torch

from .ppo_torch import gru_window

Q_KEYS = ["lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0", "layers.0.weight",
          "layers.0.bias", "layers.2.weight", "layers.2.bias", "layers.4.weight", "layers.4.bias"]


def q_forward(p, x, valid=None):
    """x [R, L, I] -> Q [R, A]; ``valid`` [R, L]: window steps that exist (the rollout history is shorter than L at the
    start of an episode, irdqn.py:242-247)."""
    h = gru_window(p, x, valid)
    y = torch.relu(h @ p["layers.0.weight"].t() + p["layers.0.bias"])
    y = torch.relu(y @ p["layers.2.weight"].t() + p["layers.2.bias"])
    return y @ p["layers.4.weight"].t() + p["layers.4.bias"]


def sample_chunk(states, states_next, actions, rewards, dones, start_idx, chunk):
    """Flat transition arrays (deque order): states / states_next [K, N, I], actions [K, N], rewards [K, N],
    dones [K] -> (s [mb, chunk, N, I], a_last [mb, N], r_last [mb, N], s' [mb, chunk, N, I], done_last [mb])."""
    idx = np.asarray(start_idx)[:, None] + np.arange(chunk)[None, :]
    last = idx[:, -1]
    return states[idx], actions[last], rewards[last], states_next[idx], dones[last]


def td_loss(q_value, td_target, loss):
    d = q_value - td_target
    if loss == "mse":
        return (d * d).mean()
    ad = d.abs()
    return torch.where(ad < 1, 0.5 * d * d, ad - 0.5).mean()


class Adam:
    """torch.optim.Adam (betas 0.9 / 0.999) on a dict of tensors, written out."""

    def __init__(self, params, lr, eps=1e-8):
        self.p, self.lr, self.eps, self.t = params, lr, eps, 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, grads):
        self.t += 1
        bc1, bc2 = 1 - 0.9 ** self.t, 1 - 0.999 ** self.t
        for k in self.p:
            g = grads[k]
            self.m[k] = 0.9 * self.m[k] + 0.1 * g
            self.v[k] = 0.999 * self.v[k] + 0.001 * g * g
            denom = self.v[k].sqrt() / np.sqrt(bc2) + self.eps
            self.p[k] = self.p[k] - (self.lr / bc1) * (self.m[k] / denom)


def train_step(p, p_target, opt, s, a_last, r_last, s_next, done_last, gamma, loss):
    """One DQN.train_step of one agent: s / s_next [mb, chunk, I] fp32, a_last [mb] int64, r_last [mb] fp32,
    done_last [mb] fp32.  Updates ``opt.p`` (and returns it with the loss)."""
    with torch.no_grad():
        q_t = q_forward(p_target, s_next).max(1, True)[0]
    td_target = r_last[:, None] + (1 - done_last[:, None]) * gamma * q_t
    leaf = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    q = q_forward(leaf, s).gather(1, a_last[:, None])
    l = td_loss(q, td_target, loss)
    grads = torch.autograd.grad(l, [leaf[k] for k in Q_KEYS])
    opt.step(dict(zip(Q_KEYS, grads)))
    return float(l.detach()), dict(zip(Q_KEYS, grads))


def epsilon_at(ep, initial=1.0, final=0.1, horizon=1000):
    return max(initial - (initial - final) * (ep / horizon), final)


def greedy_test(env, params, L):
    """iRDQN.test on the B lockstep episodes of a batched oracle env (oracle/envs_np.py; B = n_episodes).
    Returns ((score, mean reward score), actions [B, T, N])."""
    N = len(params)
    obs, _ = env.reset()
    hist = [np.stack([np.asarray(o) for o in obs], axis=1)]                     # [B, N, I]
    done, score, acts = False, np.zeros(env.B), []
    while not done:
        h = torch.tensor(np.stack(hist[-L:], axis=1), dtype=torch.float32)      # [B, l, N, I]
        a = np.stack([q_forward(params[i], h[:, :, i]).argmax(1).numpy() for i in range(N)], axis=1)   # [B, N]
        onehot = (a[:, :, None] == np.arange(env.n_channels)[None, None, :]).astype(np.uint8)
        obs, _, reward, done, _ = env.step(onehot)
        hist.append(np.stack([np.asarray(o) for o in obs], axis=1))
        score += np.maximum(np.asarray(reward, dtype=np.float64), 0).mean(1)
        acts.append(a)
    return (1 - float(env.discarded.sum()) / float(env.received.sum()), float(score.mean())), np.stack(acts, axis=1)
