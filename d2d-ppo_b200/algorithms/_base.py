"""Shared rollout / evaluation / checkpoint logic of the two learners (reference: create_rollouts, test, save,
load in algorithms/d2d_ppo.py:269-383 and algorithms/ippo.py:266-388).

B lockstep envs each running one episode ARE the reference's ``num_episodes = B`` sequential episodes
(SURVEY.md section 7): rollouts live on the device as env-minor matrices [time][row][env], the env kernels write
observations straight into them, and the policy / critic kernels read them in place.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import _lib as L
from . import _dist
from ._nets import NetSet, action_dtype, policy_head


class PPOBase:
    def _setup(self, env, hidden_size, gamma, policy_lr, value_lr, device, useRNN, save_path, combinatorial,
               history_len, early_stopping, seed, scratch_bytes):
        self.env = env
        self.history_len = int(history_len)
        self.n_agents = env.n_agents
        self.hidden_size = int(hidden_size)
        self.gamma = gamma
        self.policy_lr, self.value_lr = policy_lr, value_lr
        self.early_stopping = early_stopping
        self.useRNN = bool(useRNN)
        self.save_path = save_path
        self.combinatorial = bool(combinatorial)
        self.device = env.device if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the learners run on CUDA devices only (no CPU fallback)")
        if getattr(env, "compat", False):
            raise ValueError("construct the env with n_envs=B: the learners roll out B lockstep episodes per "
                             "iteration (B plays the role of the reference's num_episodes)")
        kind = env.action_kind
        if combinatorial and kind != "bernoulli_mask":
            raise ValueError("combinatorial=True (Bernoulli per channel) needs a CombinatorialEnv")
        if not combinatorial and kind == "bernoulli_mask":
            raise ValueError("CombinatorialEnv takes MultiBinary actions: use combinatorial=True")
        self.B, self.T = env.n_envs, env.episode_length
        self.arch = L.NET_GRU if self.useRNN else L.NET_MLP
        self.lead = self.history_len - 1 if self.useRNN else 0
        self.dist_kind = L.DIST_BERNOULLI if combinatorial else L.DIST_CATEGORICAL
        # MLP policies always end in softmax (d2d_ppo.py:81); only the GRU switches to sigmoid (:55-58)
        self.policy_out = L.OUT_SIGMOID if (self.useRNN and combinatorial) else L.OUT_SOFTMAX
        self.n_actions = env.action_space[0].n
        self.obs_rows, self.obs_off, self.obs_dim = env.obs_layout
        self.state_rows = env.state_space.shape[0]
        self.seed = int(seed)
        self._gen = torch.Generator().manual_seed(self.seed)      # identical initial weights on every rank
        self._scratch = int(scratch_bytes)
        # integer-valued observations (buffers, channel bits, ack in {-1, 0, 1}) are exact in bf16, which lets the
        # networks' rollout forward run on the tcgen05 GRU-window kernel; the selection env's 1/count acks are not
        self.exact_obs = kind in ("bernoulli_mask", "binary")
        self.policies = NetSet(self.arch, self.policy_out, self.n_agents, self.B, self.obs_dim, self.obs_off,
                               self.obs_rows, self.hidden_size, self.n_actions, self.history_len, self.device,
                               policy_lr, scratch_bytes, self._gen, inputs_bf16_exact=self.exact_obs)
        self._act_dtype = action_dtype(self.dist_kind, self.n_actions)
        self._iter = 0
        T, B, N, dev = self.T, self.B, self.n_agents, self.device
        # rollout storage, allocated once: observation blocks lead .. lead + T hold times 0 .. T
        self.obs_buf = torch.zeros((self.lead + T + 1, self.obs_rows, B), dtype=torch.float32, device=dev)
        self.act_buf = torch.zeros((T, N, B), dtype=self._act_dtype, device=dev)
        self.logp_buf = torch.zeros((T, N, B), dtype=torch.float32, device=dev)
        self.reward_buf = torch.zeros((T, B), dtype=torch.int32, device=dev)

    # ------------------------------------------------------------------ checkpoints (d2d_ppo.py:269-277)
    def save(self, checkpoint_path):
        if _dist.rank() == 0:
            os.makedirs(checkpoint_path, exist_ok=True)
            for i in range(self.n_agents):
                torch.save(self.policies.state_dict(i), f"{checkpoint_path}/agent_{i}.pth")
            print("Models saved!")

    def load(self, checkpoint_path):
        for i in range(self.n_agents):
            self.policies.load_state_dict(i, torch.load(f"{checkpoint_path}/agent_{i}.pth", map_location="cpu"))
        print("Models loaded!")

    # ------------------------------------------------------------------ rollout
    def _sample_seed(self):
        # one Philox stream per (seed, iteration): timestep t and env index key the draws inside an iteration
        return (self.seed * 0x9E3779B97F4A7C15 + self._iter * 0xD1B54A32D192ED03 + 0x2545F4914F6CDD1D) & (2 ** 64 - 1)

    def _act(self, t, mode, forced=None):
        """select_action for all agents at time t: actions into act_buf[t], log-probs into logp_buf[t]."""
        logits = self.policies.rollout_step(self.obs_buf, self.lead, t)
        if forced is not None:
            self.act_buf[t].copy_(forced)
            mode = L.ACT_GIVEN
        policy_head(logits, self.n_agents, self.B, self.n_actions, self.policy_out, self.dist_kind, mode,
                    self.act_buf[t:t + 1], self.logp_buf[t:t + 1], seed=self._sample_seed(),
                    env_offset=self.env.env_offset, t_abs0=t)

    def _run_episode(self, mode, forced_actions=None, state_buf=None, per_step=None):
        """One lockstep episode of all B envs.  forced_actions: optional [T, N, B] device-layout actions
        (teacher forcing for parity runs).  per_step(t) runs after the policy acted and before the env steps."""
        env = self.env
        env.reset_into(self.obs_buf[self.lead], None if state_buf is None else state_buf[0])
        for t in range(self.T):
            self._act(t, mode, None if forced_actions is None else forced_actions[t])
            if per_step is not None:
                per_step(t)
            done = env.step_into(self.act_buf[t], self.obs_buf[self.lead + t + 1],
                                 None if state_buf is None else state_buf[t + 1], self.reward_buf[t])
        assert done
        self._iter += 1
        return env.compute_urllc()          # score per episode: 1 - discarded / received (d2d_ppo.py:329)

    # ------------------------------------------------------------------ evaluation (d2d_ppo.py:341-383)
    def test(self, num_episodes):
        """Greedy rollouts.  Runs ceil(num_episodes / B) lockstep batches and averages over every episode run.
        Returns (mean URLLC score, mean Jain index, total channel errors, mean per-episode reward sum)."""
        batches = max(1, -(-int(num_episodes) // self.B))
        scores, jains, rewards, errors = [], [], [], 0
        for _ in range(batches):
            s = self._run_episode(L.ACT_GREEDY)
            scores.append(s)
            jains.append(self.env.compute_jains())
            rewards.append(self.reward_buf.to(torch.float64).sum(0))     # reward.mean() over identical agents
            ce = self.env.channel_errors
            errors += int(ce.sum().item()) if torch.is_tensor(ce) else int(ce) * self.B
        stats = torch.stack([torch.cat(scores).sum(), torch.cat(jains).sum(), torch.cat(rewards).sum(),
                             torch.tensor(float(errors), dtype=torch.float64, device=self.device),
                             torch.tensor(float(batches * self.B), dtype=torch.float64, device=self.device)])
        _dist.all_reduce_sum_(stats)
        n = stats[4].item()
        return stats[0].item() / n, stats[1].item() / n, int(stats[3].item()), stats[2].item() / n

    # ------------------------------------------------------------------ helpers shared by the train loops
    def _norm_stats(self, stats, cols, ddof):
        """mean / std / normalise-flags from (all-reduced) [n_cols, 4] sums; the reference normalises only if
        EVERY column has a positive std (d2d_ppo.py:108, :122)."""
        n = float(self.B * self.T * _dist.world_size())
        mean = stats[:, cols[0]] / n
        var = (stats[:, cols[1]] - n * mean * mean) / (n - ddof)
        std = var.clamp(min=0).sqrt()
        # stays on the device (no host round trip between the statistics pass and the emit pass)
        flags = (std > 0).all().to(torch.int32).expand(stats.shape[0]).contiguous()
        return mean.contiguous(), std.contiguous(), flags

    def _check_episodes(self, num_episodes):
        if num_episodes is not None and int(num_episodes) != self.B:
            raise ValueError(f"num_episodes={num_episodes} but the env runs n_envs={self.B} lockstep episodes per "
                             f"rollout; build the env with n_envs=num_episodes")

    def _maybe_test(self, it, epoch, test_freq, scores, score_test_list):
        """The reference's test / best-model / early-stop block, run once per epoch when iter % test_freq == 0
        (d2d_ppo.py:450-459, ippo.py:428-438).  Returns True to stop training."""
        if it % test_freq != 0:
            return False
        score_test, jains, channel_loss, avg_rewards = self.test(50)
        score_test_list.append(score_test)
        if _dist.rank() == 0:
            print(f"Iteration: {it}, Epoch: {epoch}, score rollout: {np.mean(scores)} "
                  f"Score test: {(score_test, jains, channel_loss, avg_rewards)}")
        if np.max(score_test_list) == score_test and self.save_path is not None:
            self.save(self.save_path)
        return bool((score_test == 1) & bool(self.early_stopping))
