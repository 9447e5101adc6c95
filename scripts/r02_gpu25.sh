#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/dbg_c5_op.py 4096 > gpurun_out/g25_op.log 2>&1; tail -2 gpurun_out/g25_op.log | cut -c1-600
timeout 900 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu > gpurun_out/g25_c5.json 2> gpurun_out/g25_c5.err
echo "rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/g25_c5.json'));print(b['value'],b['ms_per_step'],b['config']['variant'][:120],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'],b['config']['converged_instances'])"
