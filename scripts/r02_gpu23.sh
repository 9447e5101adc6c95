#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/g23_1.json 2> gpurun_out/g23_1.err
echo "rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/g23_1.json'));print(b['value'],b['ms_per_step'],b['config']['variant'][:50],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'])"
