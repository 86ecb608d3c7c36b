#!/bin/bash
# tests -> bench -> ncu launch list -> ncu full captures (each ncu pass only after its command exited 0 without ncu)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_a.log 2>&1; echo "pytest=$?"
tail -3 gpurun_out/pytest_a.log
( time python bench.py --steps 2000 --warmup 20 > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err ) 2> gpurun_out/bench_a.time; echo "bench=$?"
cat gpurun_out/bench_a.time | tail -3
python bench.py --steps 40 --warmup 3 --no-cpu-baseline --rollout-envs 16384 --train-envs 1024 > gpurun_out/bench_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench_small.csv \
  python bench.py --steps 40 --warmup 3 --no-cpu-baseline --rollout-envs 16384 --train-envs 1024 > gpurun_out/ncu_small.log 2>&1
echo "launchlist=$?"
# full captures: env step (skip warm-up launches), at the bench shape
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-learner > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:comb_step_kernel -s 20 -c 2 -o gpurun_out/prof_env_step_v5 -f \
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-learner > gpurun_out/ncu_env.log 2>&1
echo "ncu_env=$?"
ncu --set full --clock-control none --import-source on -k regex:"returns_scan_kernel|gru_window_tc_kernel" -s 40 -c 6 -o gpurun_out/prof_gae_gru -f \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --rollout-envs 65536 --train-envs 256 > gpurun_out/ncu_gae.log 2>&1
echo "ncu_gae=$?"
