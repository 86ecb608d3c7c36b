#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest=$?"; tail -2 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err ) 2> gpurun_out/bench_final.time; echo "bench=$?"; tail -3 gpurun_out/bench_final.time
python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/bench_ref_final.log 2>&1; echo "ref=$?"
# launch list of the headline bench command (env step section only: the timed region of `value`)
python bench.py --steps 300 --warmup 5 --no-learner --no-configs --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench_final.csv \
  python bench.py --steps 300 --warmup 5 --no-learner --no-configs --no-cpu-baseline > /dev/null 2>&1; echo "launchlist=$?"
# full captures of the learner kernels at the bench's training shape
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gru_bwd_tc|gru_window_tc|head_fused|wgrad_tc" -c 5 -o gpurun_out/prof_learner_final -f python profiles/prof_learner.py 4096 200 > /dev/null 2>&1; echo "ncu_learner=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/epoch_final.csv python profiles/prof_learner.py 4096 200 > /dev/null 2>&1; echo "epoch_list=$?"
