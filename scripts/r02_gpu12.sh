#!/bin/bash
# compile-time-rows per-team kernel (linear circuits): parity, then C4 with it on / off
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_q9_isource.py tests/test_fullspan.py -m gpu -x -q -k "mid_size or team or q9 or current_source or ladder or c4 or tran" > gpurun_out/g12_tests.log 2>&1
echo "rc=$?" >> gpurun_out/g12_tests.log; tail -n 8 gpurun_out/g12_tests.log
for opt in "team_rows=1" "team_rows=0"; do
  timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu --opt $opt > gpurun_out/g12_c4_$opt.json 2> gpurun_out/g12_c4_$opt.err
  echo "bench c4 $opt rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/g12_c4_$opt.json'));print(b['value'],b['ms_per_step'],b['config']['variant'],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'])"
done
