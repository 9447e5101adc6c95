#!/bin/bash
# bisect of the per-attempt micro-optimisations: library variants A (no prologue/update batching), B (old streams), C (only b build + PLTE)
mkdir -p gpurun_out
for v in D E F; do
  cp build_variants/libftsb_$v.so ftspice_b200/libftsb.so
  timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/g21_$v.json 2> gpurun_out/g21_$v.err
  echo "[$v] rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/g21_$v.json'));print(b['value'],b['ms_per_step'],b['config']['variant'][:50],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'])"
done
