#!/bin/bash
# C5 bench slice (4 096 instances x 4 ns, start-up included): the A/B line of the round's C5 experiments
mkdir -p gpurun_out
timeout 900 python bench.py --workload c5 --scale 0.25 --tran-stop 4e-9 --steps 2 --warmup 3 --no-cpu $EXTRA > gpurun_out/c5s.json 2> gpurun_out/c5s.err
echo "rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/c5s.json'));print(b['value'],b['ms_per_step'],b['config']['variant'][:60],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'],b['config']['converged_instances'])"
