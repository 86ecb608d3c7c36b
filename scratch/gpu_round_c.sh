#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:"returns_scan_kernel" -s 2 -c 2 -o gpurun_out/prof_gae -f \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --train-envs 256 > gpurun_out/ncu_gae2.log 2>&1
echo "ncu_gae=$?"
ncu --set full --clock-control none --import-source on -k regex:"sc_step_kernel|comb_step_kernel|sel_step_kernel" -s 60 -c 1400 --kernel-id :::"40|280|520|760|1000|1240|1390" -o gpurun_out/prof_env_cfgs -f \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-learner > gpurun_out/ncu_cfgs.log 2>&1
echo "ncu_cfgs=$?"
