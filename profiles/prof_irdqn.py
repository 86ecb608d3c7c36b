"""ncu target: one iRDQN iteration (B lockstep epsilon-greedy episodes + one minibatch update) on the c3 env.
usage: python profiles/prof_irdqn.py [hidden=100] [B=4096] [T=20]"""
import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from d2d_ppo_b200 import presets
from d2d_ppo_b200.algorithms.irdqn import iRDQN
from d2d_ppo_b200.envs import CombinatorialEnv
dev = torch.device("cuda", 0)
hidden = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = int(sys.argv[3]) if len(sys.argv) > 3 else 20
kw = presets.combinatorial_kwargs("setup_8_channels", load=1 / 3, episode_length=T)
env = CombinatorialEnv(n_envs=B, device=dev, seed=7, **kw)
ag = iRDQN(env, history_len=6, replay_start_size=1, replay_buffer_size=100000, gamma=0.4, update_target_frequency=100,
           minibatch_size=64, learning_rate=1e-4, loss="huber", early_stopping=False, hidden_size=hidden, seed=5)
ag.test = lambda *a, **k: (0.0, 0.0)
with contextlib.redirect_stdout(io.StringIO()):
    ag.train(2, early_stopping=False)
    ag.replay_start_size = 0
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ag.train(1, early_stopping=False)
    torch.cuda.profiler.stop()
torch.cuda.synchronize()
