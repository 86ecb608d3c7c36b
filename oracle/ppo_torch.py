"""CPU restatement (torch, fp32) of the PPO / IPPO learner maths of the reference.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity pinning: checked against the unmodified reference
(``algorithms/ippo.py``, ``algorithms/d2d_ppo.py``) through ``tests/golden/ppo_*.npz`` written by
``oracle/gen_golden_ppo.py`` and, when the reference is mounted, directly (tests/test_oracle_ppo.py).

Restated (file:line under /root/reference):
* ``gru_window`` / ``net_forward``  <- ``RNN`` / ``Policy`` / ``Value`` forward, d2d_ppo.py:24-98, ippo.py:14-90
                                     (GRU gates written out: torch order r, z, n; zero initial hidden state)
* ``bernoulli_logp_entropy`` / ``categorical_logp_entropy`` <- ``select_action`` / ``evaluate``,
                                     d2d_ppo.py:159-196, with torch.distributions' clamping (eps = 2^-23)
* ``windows``                       <- rollout windows (d2d_ppo.py:302, unpadded) and training windows
                                     (``preprocess_input_for_rnn``, d2d_ppo.py:385-398, left zero padding)
* ``lambda_returns``                <- ``compute_gae`` d2d_ppo.py:100-110 (numpy float64, population std)
* ``discounted_returns``            <- ``discount_rewards`` d2d_ppo.py:112-124 (float64 scan, fp32 unbiased std)
* ``ippo_train_step`` / ``d2dppo_epoch`` <- ippo.py:194-217, d2d_ppo.py:198-216 and :413-446
* ``greedy_test``                   <- ``test`` d2d_ppo.py:341-383, ippo.py:345-388 (greedy rollout, 4-tuple)

Row order: the reference concatenates episodes along axis 0 (row = e * T + t); B lockstep envs running one
episode each ARE ``num_episodes = B``.  Every function below takes rows in that episode-major order.
"""
from __future__ import annotations

import numpy as np
import torch

EPS = float(torch.finfo(torch.float32).eps)


# ----------------------------------------------------------------------------------------------
# networks: parameters are plain dicts with the reference's state_dict keys
# ----------------------------------------------------------------------------------------------
def init_params(arch, n_in, n_out, hidden, generator):
    """Same initialisation law as the reference (orthogonal gains 2 / 3, zero biases, GRU U(+-1/sqrt(H)))."""
    def ortho(rows, cols, gain):
        w = torch.empty(rows, cols)
        torch.nn.init.orthogonal_(w, gain, generator=generator)
        return w
    if arch == "mlp":
        return {"linear1.weight": ortho(hidden, n_in, 2), "linear1.bias": torch.zeros(hidden),
                "linear2.weight": ortho(n_out, hidden, 2), "linear2.bias": torch.zeros(n_out)}
    k = 1.0 / np.sqrt(hidden)

    def uni(*shape):
        return (torch.rand(*shape, generator=generator) * 2 - 1) * k
    return {"lstm.weight_ih_l0": uni(3 * hidden, n_in), "lstm.weight_hh_l0": uni(3 * hidden, hidden),
            "lstm.bias_ih_l0": uni(3 * hidden), "lstm.bias_hh_l0": uni(3 * hidden),
            "layers.0.weight": ortho(hidden, hidden, 3), "layers.0.bias": torch.zeros(hidden),
            "layers.2.weight": ortho(n_out, hidden, 3), "layers.2.bias": torch.zeros(n_out)}


def gru_window(p, x, valid=None):
    """x [R, L, I] -> last hidden [R, H].  ``valid`` [R, L] bool: steps to run (False = skipped, h unchanged);
    the unpadded rollout window of d2d_ppo.py:302 is 'skip the missing leading steps'."""
    w_ih, w_hh = p["lstm.weight_ih_l0"], p["lstm.weight_hh_l0"]
    b_ih, b_hh = p["lstm.bias_ih_l0"], p["lstm.bias_hh_l0"]
    H = w_hh.shape[1]
    h = torch.zeros(x.shape[0], H, dtype=x.dtype)
    for s in range(x.shape[1]):
        gi = x[:, s] @ w_ih.t() + b_ih
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h_new = (1 - z) * n + z * h
        h = h_new if valid is None else torch.where(valid[:, s:s + 1], h_new, h)
    return h


def net_forward(p, x, out, valid=None):
    """out: 'softmax' | 'sigmoid' | 'identity'.  x [R, I] for the MLP, [R, L, I] for the GRU net."""
    if "linear1.weight" in p:
        y = torch.relu(x @ p["linear1.weight"].t() + p["linear1.bias"]) @ p["linear2.weight"].t() + p["linear2.bias"]
    else:
        h = gru_window(p, x, valid)
        y = torch.relu(h @ p["layers.0.weight"].t() + p["layers.0.bias"]) @ p["layers.2.weight"].t() \
            + p["layers.2.bias"]
    if out == "softmax":
        return torch.softmax(y, dim=1)
    if out == "sigmoid":
        return torch.sigmoid(y)
    return y


def policy_out_kind(arch, combinatorial):
    """MLP policies always end in softmax (d2d_ppo.py:81), even under Bernoulli; the GRU switches (:55-58)."""
    return "softmax" if (arch == "mlp" or not combinatorial) else "sigmoid"


def windows(obs, T, L, pad):
    """obs [R, I] episode-major rows -> (x [R, L, I], valid [R, L]).
    pad=True : training windows, missing leading steps are ZERO INPUTS that still run (d2d_ppo.py:385-398).
    pad=False: rollout windows, missing leading steps do not exist (d2d_ppo.py:302)."""
    R, I = obs.shape
    x = torch.zeros(R, L, I, dtype=obs.dtype)
    valid = torch.ones(R, L, dtype=torch.bool)
    t = torch.arange(R) % T
    for s in range(L):
        back = L - 1 - s
        ok = t >= back
        src = (torch.arange(R) - back).clamp(min=0)
        x[:, s] = torch.where(ok[:, None], obs[src], torch.zeros_like(obs))
        if not pad:
            valid[:, s] = ok
    return x, valid


# ----------------------------------------------------------------------------------------------
# distributions (torch.distributions numerics written out)
# ----------------------------------------------------------------------------------------------
def _bce_with_logits(logits, target):
    # ATen: (1 - target) * input - log_sigmoid(input),  log_sigmoid(x) = min(x, 0) - log1p(exp(-|x|))
    return (1 - target) * logits - (torch.clamp(logits, max=0) - torch.log1p(torch.exp(-logits.abs())))


def bernoulli_logp_entropy(probs, actions):
    """Bernoulli(probs): log_prob(a).mean(-1), entropy().mean(-1)  (d2d_ppo.py:162-169)."""
    pc = probs.clamp(EPS, 1 - EPS)
    logits = torch.log(pc) - torch.log1p(-pc)
    logp = -_bce_with_logits(logits, actions.to(probs.dtype))
    ent = _bce_with_logits(logits, probs)
    return logp.mean(-1), ent.mean(-1)


def categorical_logp_entropy(probs, actions):
    """Categorical(probs): probs renormalised, logits = log(clamp(p))  (d2d_ppo.py:172-179)."""
    p = probs / probs.sum(-1, keepdim=True)
    logits = torch.log(p.clamp(EPS, 1 - EPS))
    logp = logits.gather(-1, actions.long().reshape(-1, 1)).squeeze(-1)
    ent = -(logits * p).sum(-1)
    return logp, ent


def logp_entropy(probs, actions, combinatorial):
    return bernoulli_logp_entropy(probs, actions) if combinatorial else categorical_logp_entropy(probs, actions)


# ----------------------------------------------------------------------------------------------
# returns
# ----------------------------------------------------------------------------------------------
def lambda_returns(rewards, dones, values, gamma, lam):
    """compute_gae (d2d_ppo.py:100-110): lambda-return, float64 scan, population-std normalisation, cast to fp32.
    rewards/values [R] or [R, N] float64-able, dones [R] bool."""
    r = np.asarray(rewards, dtype=np.float64)
    v = np.asarray(values, dtype=np.float64)
    nd = 1.0 - np.asarray(dones, dtype=np.float64)
    out = np.empty_like(np.broadcast_to(r, np.broadcast_shapes(r.shape, v.shape)), dtype=np.float64)
    out[-1] = r[-1] - v[-1]
    gae = np.zeros_like(out[-1])
    for s in range(len(r) - 2, -1, -1):
        delta = r[s] + gamma * v[s + 1] * nd[s] - v[s]
        gae = delta + gamma * lam * nd[s] * gae
        out[s] = gae + v[s]
    sd = out.std(0)
    if (sd > 0).all():
        out = (out - out.mean(0)) / sd
    return torch.tensor(out, dtype=torch.float32)


def discounted_returns(rewards, gamma, dones):
    """discount_rewards (d2d_ppo.py:112-124): float64 scan -> fp32 -> (x - mean) / unbiased std in fp32."""
    r = np.asarray(rewards, dtype=np.float64)
    nd = 1.0 - np.asarray(dones, dtype=np.float64)
    out = np.empty_like(r)
    run = np.zeros_like(r[0])
    for s in range(len(r) - 1, -1, -1):
        run = r[s] + run * gamma * nd[s]
        out[s] = run
    ret = torch.tensor(out, dtype=torch.float32)
    if (ret.std(0) > 0).all():
        ret = (ret - ret.mean(0)) / ret.std(0)
    return ret


# ----------------------------------------------------------------------------------------------
# updates
# ----------------------------------------------------------------------------------------------
class Adam:
    """torch.optim.Adam defaults (betas .9/.999, eps 1e-8, no weight decay), written out."""

    def __init__(self, params, lr):
        self.p, self.lr, self.t = params, lr, 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, grads):
        self.t += 1
        b1, b2 = 0.9, 0.999
        for k in self.p:
            g = grads[k]
            self.m[k] = b1 * self.m[k] + (1 - b1) * g
            self.v[k] = b2 * self.v[k] + (1 - b2) * g * g
            mhat = self.m[k] / (1 - b1 ** self.t)
            denom = (self.v[k] / (1 - b2 ** self.t)).sqrt() + 1e-8
            self.p[k] = self.p[k] - self.lr * mhat / denom


def _grads(loss, params):
    names = list(params)
    gs = torch.autograd.grad(loss, [params[n] for n in names])
    return dict(zip(names, gs))


def clip_grads(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (d2d_ppo.py:211,445)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return {k: g * coef for k, g in grads.items()}, total


def surrogate(probs, actions, logp_old, weight, combinatorial, cliprange, beta):
    """-mean(min(ratio w, clip(ratio) w)) - beta mean(entropy); also returns the ratio (d2d_ppo.py:201-207)."""
    logp, ent = logp_entropy(probs, actions, combinatorial)
    ratio = torch.exp(logp - logp_old)
    loss = -torch.min(ratio * weight, ratio.clamp(1 - cliprange, 1 + cliprange) * weight).mean() - beta * ent.mean()
    return loss, ratio


def ippo_train_step(pol, val, opt_p, opt_v, x, valid, actions, logp_old, returns, adv, arch, combinatorial,
                    cliprange=0.1, beta=0.01):
    """ippo.py:194-217: policy step (no grad clip), then critic MSE step.  Returns (policy_loss, value_loss)."""
    pp = {k: v.clone().requires_grad_(True) for k, v in pol.items()}
    probs = net_forward(pp, x, policy_out_kind(arch, combinatorial), valid)
    loss, _ = surrogate(probs, actions, logp_old, adv, combinatorial, cliprange, beta)
    opt_p.p = pol
    opt_p.step(_grads(loss, pp))
    pol.update(opt_p.p)
    vp = {k: v.clone().requires_grad_(True) for k, v in val.items()}
    value = net_forward(vp, x, "identity", valid).squeeze(-1)
    vloss = ((value - returns) ** 2).mean()
    opt_v.p = val
    opt_v.step(_grads(vloss, vp))
    val.update(opt_v.p)
    return float(loss.detach()), float(vloss.detach())


def d2dppo_epoch(pols, opts, critic, opt_c, xs, valids, states, actions, logp_old, rewards, dones, returns, cycle,
                 arch, combinatorial, gamma, beta, cliprange=0.1, lam=0.97, max_norm=20.0):
    """One epoch of d2d_ppo.py:413-446 for a given agent order ``cycle``.
    xs[i]/valids[i]: agent i's inputs, actions [R, N(, C)], logp_old [R, N], rewards/returns [R] (agent means)."""
    cp = {k: v.clone().requires_grad_(True) for k, v in critic.items()}
    values = net_forward(cp, states, "identity").squeeze(-1)
    M = lambda_returns(rewards, dones, values.detach().numpy(), gamma, lam)
    losses = []
    for i in cycle:
        pp = {k: v.clone().requires_grad_(True) for k, v in pols[i].items()}
        probs = net_forward(pp, xs[i], policy_out_kind(arch, combinatorial), valids[i])
        loss, ratio = surrogate(probs, actions[:, i], logp_old[:, i], M.detach(), combinatorial, cliprange, beta)
        g, _ = clip_grads(_grads(loss, pp), max_norm)
        opts[i].p = pols[i]
        opts[i].step(g)
        pols[i].update(opts[i].p)
        M = (ratio * M).detach()
        losses.append(float(loss.detach()))
    vloss = ((values - returns) ** 2).mean()
    g, _ = clip_grads(_grads(vloss, cp), max_norm)
    opt_c.p = critic
    opt_c.step(g)
    critic.update(opt_c.p)
    return losses, float(vloss.detach())


# ----------------------------------------------------------------------------------------------
# evaluation
# ----------------------------------------------------------------------------------------------
def greedy_test(pols, env, arch, combinatorial, history_len):
    """``test(num_episodes)`` (d2d_ppo.py:341-383): greedy rollouts (probs > 0.5 per channel, or argmax) of the B
    lockstep episodes of the oracle env ``env`` (= num_episodes = B).  Returns (actions [B*T, N(, C)] episode-major,
    (mean URLLC score, mean Jain index, sum of channel errors, mean per-episode sum of reward.mean()))."""
    B, N, T = env.B, env.n_agents, env.episode_length
    obs, _ = env.reset()
    hist = [[torch.tensor(np.asarray(obs[i]), dtype=torch.float32)] for i in range(N)]
    acts, rew_sum = [], np.zeros(B)
    done = False
    while not done:
        step_actions = []
        for i in range(N):
            if arch == "gru":
                x = torch.stack(hist[i][-history_len:], dim=1)                  # unpadded window (d2d_ppo.py:361)
            else:
                x = hist[i][-1]
            probs = net_forward(pols[i], x, policy_out_kind(arch, combinatorial))
            step_actions.append((probs > 0.5).to(torch.int64) if combinatorial else probs.argmax(1))
        a = torch.stack(step_actions, dim=1).numpy()                            # [B, N(, C)]
        obs, _, reward, done, _ = env.step(a)
        for i in range(N):
            hist[i].append(torch.tensor(np.asarray(obs[i]), dtype=torch.float32))
        acts.append(a)
        rew_sum += np.asarray(reward, dtype=np.float64).mean(1)
    acts = np.stack(acts, axis=1).reshape((B * T,) + acts[0].shape[1:])
    errors = int(np.sum(getattr(env, "channel_errors", 0)))
    return acts, (float(np.mean(env.compute_urllc())), float(np.mean(env.compute_jains())), errors,
                  float(np.mean(rew_sum)))
