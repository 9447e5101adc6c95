#!/bin/bash
# first GPU pass of round 2: new tests, then the whole suite, then C4 A/B of the per-team kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/g1_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mid_size or sparse_kernel" > gpurun_out/g1_team_tests.log 2>&1
echo "team tests rc=$?" >> gpurun_out/g1_team_tests.log
timeout 900 python -m pytest tests/test_q9_isource.py tests/test_fullspan.py -m gpu -q > gpurun_out/g1_parity_tests.log 2>&1
echo "parity tests rc=$?" >> gpurun_out/g1_parity_tests.log
for opt in "team=0" "team_lanes=8" "team_lanes=4" "team_lanes=16" "team_lanes=32" "team_lanes=8 --opt team_place=1"; do
  tag=$(echo "$opt" | tr -c 'a-z0-9=' '_')
  timeout 600 python bench.py --workload c4 --scale 0.25 --tran-stop 0.54e-6 --steps 2 --warmup 3 --no-cpu --opt $opt > gpurun_out/g1_c4_$tag.json 2> gpurun_out/g1_c4_$tag.err
  echo "bench $opt rc=$?" >> gpurun_out/g1_bench.log
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/g1_all_tests.log 2>&1
echo "all tests rc=$?" >> gpurun_out/g1_all_tests.log
tail -3 gpurun_out/g1_team_tests.log gpurun_out/g1_parity_tests.log gpurun_out/g1_all_tests.log
cat gpurun_out/g1_bench.log
