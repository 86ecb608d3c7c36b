import sys, time
sys.path.insert(0, "/root/repo")
import torch
from d2d_ppo_b200 import presets, _lib
from d2d_ppo_b200.algorithms.ippo import iPPO
from d2d_ppo_b200.envs import CombinatorialEnv
dev = torch.device("cuda", 0)
B = 65536
kw = presets.combinatorial_kwargs("setup_8_channels", load=1 / 3)
env = CombinatorialEnv(n_envs=B, device=dev, seed=7, **kw)
ag = iPPO(env, hidden_size=64, gamma=0.4, policy_lr=3e-4, value_lr=1e-3, useRNN=True, combinatorial=True, history_len=6, early_stopping=False, seed=1, scratch_bytes=6 << 30)
ag._run_episode(_lib.ACT_SAMPLE)
def timeit(name, fn, n=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name:28s} host {1e3*(t1-t0)/n:7.3f} ms/call   total {1e3*(t2-t0)/n:7.3f} ms/call")
timeit("policy.rollout_step", lambda i: ag.policies.rollout_step(ag.obs_buf, ag.lead, 10 + i % 100))
timeit("values.rollout_step", lambda i: ag.values.rollout_step(ag.obs_buf, ag.lead, 10 + i % 100))
timeit("_act", lambda i: ag._act(10 + i % 100, _lib.ACT_SAMPLE))
env.reset_into(ag.obs_buf[ag.lead])
timeit("env.step_into", lambda i: env.step_into(ag.act_buf[i], ag.obs_buf[ag.lead + i + 1], None, ag.reward_buf[i]))
torch.cuda.synchronize(); t0 = time.perf_counter(); ag._run_episode(_lib.ACT_SAMPLE); torch.cuda.synchronize(); print("episode (policy only) s:", time.perf_counter() - t0)
torch.cuda.synchronize(); t0 = time.perf_counter(); ag.create_rollouts(B); torch.cuda.synchronize(); print("create_rollouts s:", time.perf_counter() - t0)
import subprocess, threading
rows = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in p.stdout], daemon=True).start()
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ag.create_rollouts(B); torch.cuda.synchronize(); print("create_rollouts s:", time.perf_counter() - t0)
p.terminate()
print("clock samples:", rows[::3])
