"""Diagnostic: where does the time of a short multi-rank timed window go?  (torchrun, N ranks)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from d2d_ppo_b200 import presets
from d2d_ppo_b200.envs import CombinatorialEnv
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B = 1 << 20
kw = presets.combinatorial_kwargs("setup_8_channels", load=1 / 3)
env = CombinatorialEnv(n_envs=B, device=dev, seed=42, env_offset=rank * B, **kw)
obs = torch.empty((180, B), device=dev); rew = torch.zeros(B, dtype=torch.int32, device=dev)
def run(n): env.run_random_access(0.2, n, auto_reset=True, out_obs=obs, out_reward=rew)
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
run(100); torch.cuda.synchronize()
for mode in ("nobarrier", "barrier", "barrier+sleep", "barrier_noafter"):
    for K in (20, 20, 200):
        if mode != "nobarrier": barrier()
        if mode == "barrier+sleep": time.sleep(0.01)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); run(K); e1.record()
        t1 = time.perf_counter()
        if mode in ("barrier", "barrier+sleep"): barrier()
        torch.cuda.synchronize()
        print(f"rank {rank} {mode:16s} K={K:4d} device {e0.elapsed_time(e1):8.3f} ms ({e0.elapsed_time(e1)/K:.4f}/step) host enqueue {1e3*(t1-t0):.3f} ms", flush=True)
if world > 1: dist.destroy_process_group()
