#!/bin/bash
# second GPU pass: team kernel v2 — tests first, then C4 A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_team_plan.py tests/test_q9_isource.py tests/test_multi_device.py -m gpu -q -x > gpurun_out/g2_tests_a.log 2>&1
echo "rc=$?" >> gpurun_out/g2_tests_a.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mid_size or sparse_kernel or tran_ladder" > gpurun_out/g2_tests_b.log 2>&1
echo "rc=$?" >> gpurun_out/g2_tests_b.log
for opt in "team_lanes=8" "team_lanes=16" "team_lanes=4" "team_lanes=32"; do
  tag=$(echo "$opt" | tr -c 'a-z0-9=' '_')
  timeout 600 python bench.py --workload c4 --scale 0.25 --tran-stop 0.54e-6 --steps 2 --warmup 3 --no-cpu --opt $opt > gpurun_out/g2_c4_$tag.json 2> gpurun_out/g2_c4_$tag.err
  echo "bench $opt rc=$?" >> gpurun_out/g2_bench.log
done
timeout 900 python -m pytest tests/test_fullspan.py -m gpu -q > gpurun_out/g2_tests_c.log 2>&1
echo "rc=$?" >> gpurun_out/g2_tests_c.log
tail -n 4 gpurun_out/g2_tests_a.log gpurun_out/g2_tests_b.log gpurun_out/g2_tests_c.log
cat gpurun_out/g2_bench.log
