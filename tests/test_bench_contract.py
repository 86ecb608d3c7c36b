"""bench.py contract checks that run without a GPU: the reference arm (CPU port of the path) prints one JSON line with
the keys the driver reads, and under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "3",
                           "--warmup", "1", "--cpu-envs", "64"], capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "agent-steps/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("agent-steps/sec") and j["value"] > 0 and j["steps"] == 3
    # kind "reference" when the unmodified reference is staged in oracle/_ref (build container, GPU box), else the port
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["port_value"] > 0
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and j["scaling"] == "weak" and j["vs_baseline"] is None and j["dtype"] == "u8"


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_learner_baseline_runs():
    """oracle/ippo_cpu.py (the CPU baseline of bench.py's rollout / train figures) completes a tiny iteration."""
    import sys
    sys.path.insert(0, ROOT)
    from d2d_ppo_b200 import presets
    from oracle.ippo_cpu import ippo_iteration_cpu
    kw = presets.combinatorial_kwargs("setup_8_channels", load=0.5, episode_length=8)
    r = ippo_iteration_cpu(3, kw, n_epoch=1, hidden=16, history_len=3)
    assert r["agent_steps"] == 3 * 8 * 6 and r["rollout_s"] > 0 and r["update_s"] > 0


def test_staged_reference_is_unmodified_and_times():
    """oracle/make_ref.py stages the reference files byte for byte (hash manifest) and oracle/ref_timing.py drives them
    through their own API; skipped where neither /root/reference nor a staged copy exists."""
    import pytest
    sys.path.insert(0, ROOT)
    from oracle import make_ref, ref_timing
    if not make_ref.available() and make_ref.stage() is None:
        pytest.skip("no reference tree and no staged copy")
    assert make_ref.verify()
    if os.path.isdir(make_ref.SRC):
        import filecmp
        for rel in make_ref.FILES:
            assert filecmp.cmp(os.path.join(make_ref.SRC, rel), os.path.join(make_ref.DST, rel), shallow=False), rel
    from d2d_ppo_b200 import presets
    kw = presets.combinatorial_kwargs("setup_8_channels", load=1 / 3, episode_length=20)
    v, dt = ref_timing.random_access_throughput(kw, 0.2, 2, 1)
    assert v > 0 and dt > 0
