#!/bin/bash
# team kernel v4: parity tests, C4 quarter bench (rows in flight 3 / 5, lanes 8 / 16), ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_q9_isource.py tests/test_team_plan.py tests/test_multi_device.py -m gpu -x -q -k "mid_size or sparse_kernel or tran_ladder or q9 or current_source or division or multi" > gpurun_out/g4_tests.log 2>&1
echo "rc=$?" >> gpurun_out/g4_tests.log
for opt in "team_rows=5" "team_rows=3" "team_rows=5 --opt team_lanes=16" "team_rows=5 --opt l2_persist=0"; do
  tag=$(echo "$opt" | tr -c 'a-z0-9=' '_')
  timeout 600 python bench.py --workload c4 --scale 0.25 --tran-stop 0.54e-6 --steps 2 --warmup 3 --no-cpu --opt $opt > gpurun_out/g4_c4_$tag.json 2> gpurun_out/g4_c4_$tag.err
  echo "bench $opt rc=$?" >> gpurun_out/g4_bench.log
done
tail -n 3 gpurun_out/g4_tests.log; cat gpurun_out/g4_bench.log
bash scripts/r02_ncu_c4.sh > /dev/null 2>&1
