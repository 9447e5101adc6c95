#!/bin/bash
# level-scheduled elimination / substitutions of the per-team kernel: parity (nonlinear mid-size circuits, C5 cases), then C5 slices
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mid_size or sparse_kernel or tran_inverter or inverter or c5" > gpurun_out/g22_tests.log 2>&1
echo "rc=$?" >> gpurun_out/g22_tests.log; tail -n 5 gpurun_out/g22_tests.log
i=0
for opts in "--opt team_place=6"; do
  i=$((i+1))
  timeout 900 python bench.py --workload c5 --scale 0.25 --steps 2 --warmup 3 --no-cpu $opts > gpurun_out/g22_$i.json 2> gpurun_out/g22_$i.err
  echo "[$opts] rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/g22_$i.json'));print(b['value'],b['ms_per_step'],b['config']['variant'][:60],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'],b['config']['converged_instances'])"
done
