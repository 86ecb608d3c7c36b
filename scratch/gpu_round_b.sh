#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -2
python profiles/prof_learner.py 65536 8 > gpurun_out/prof_plain_b.log 2>&1 && cat gpurun_out/prof_plain_b.log && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_learner_b.csv \
  python profiles/prof_learner.py 65536 8 > gpurun_out/prof_ncu_b.log 2>&1
echo "launchlist=$?"
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel|gru_gate_bwd_kernel|dense_tile_kernel" -s 60 -c 6 -o gpurun_out/prof_bwd_kernels -f \
  python profiles/prof_learner.py 65536 8 > gpurun_out/ncu_bwd.log 2>&1
echo "ncu_bwd=$?"
