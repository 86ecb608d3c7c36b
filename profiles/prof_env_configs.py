"""ncu target: a few steps of the non-flagship env configs (c2 D2DEnv, c4 N = 64, ChannelSelectionEnv)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from d2d_ppo_b200 import presets
from d2d_ppo_b200.envs import ChannelSelectionEnv, CombinatorialEnv, D2DEnv
dev = torch.device("cuda", 0)
which = sys.argv[1:] or ["c2", "c4", "sel"]
def run(env, tp=None, actions=None, n=6):
    obs = torch.empty((env.obs_layout[0], env.n_envs), dtype=torch.float32, device=dev)
    rew = torch.empty(env.n_envs, dtype=torch.int32, device=dev)
    env._reset_device(True, False, out_obs=obs)
    for _ in range(n):
        env._step_device(actions, True, False, obs, None, random_access_tp=tp, out_reward=rew)
    torch.cuda.synchronize()
if "c2" in which:
    run(D2DEnv(n_envs=1 << 22, device=dev, seed=2, **presets.d2d_c2_kwargs()), tp=0.2)
if "c4" in which:
    run(CombinatorialEnv(n_envs=(1 << 22) // 64, device=dev, seed=3, **presets.n_agents_sweep_kwargs(64, load=1 / 3)), tp=0.2)
if "sel" in which:
    N, C, B = 5, 16, 1 << 20
    sel = ChannelSelectionEnv(n_agents=N, n_channels=C, deadlines=np.array([7] * N), lbdas=np.array([1 / 3] * N),
                              period=None, arrival_probs=None, offsets=None, episode_length=200,
                              traffic_model="aperiodic", periodic_devices=[], channel_switch=np.array([0.2] * (C + 1)),
                              n_envs=B, device=dev, seed=4)
    run(sel, actions=torch.randint(0, C + 1, (N, B), dtype=torch.uint8, device=dev))
