#!/bin/bash
# round-2 validation pass: whole GPU suite, the default bench line (all configs), reference arm, C5 with / without L2 persistence
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/g5_all_tests.log 2>&1
echo "rc=$?" >> gpurun_out/g5_all_tests.log
tail -n 5 gpurun_out/g5_all_tests.log
( time timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/g5_bench_all.json 2> gpurun_out/g5_bench_all.err ) 2> gpurun_out/g5_bench_all.time
echo "bench all rc=$?"; tail -n 3 gpurun_out/g5_bench_all.time
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g5_bench_ref.json 2> gpurun_out/g5_bench_ref.err
echo "bench ref rc=$?"
for opt in "l2_persist=1" "l2_persist=0"; do
  timeout 900 python bench.py --workload c5 --scale 0.25 --steps 2 --warmup 3 --no-cpu --opt $opt > gpurun_out/g5_c5_$opt.json 2> gpurun_out/g5_c5_$opt.err
  echo "bench c5 $opt rc=$?"
done
