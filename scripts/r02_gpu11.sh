#!/bin/bash
# C5 through the per-team kernel (factors in shared memory, direct streams): parity, then the C5 bench slice with team on / off
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_q9_isource.py -m gpu -x -q -k "mid_size or sparse_kernel or tran_inverter or q9 or current_source" > gpurun_out/g11_tests.log 2>&1
echo "rc=$?" >> gpurun_out/g11_tests.log; tail -n 15 gpurun_out/g11_tests.log
for opt in "team=1" "team=0"; do
  timeout 900 python bench.py --workload c5 --scale 0.25 --steps 2 --warmup 3 --no-cpu --opt $opt > gpurun_out/g11_c5_$opt.json 2> gpurun_out/g11_c5_$opt.err
  echo "bench c5 $opt rc=$?"
done
