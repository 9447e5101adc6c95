#!/bin/bash
# C3: points per .dc chain (1 M points in total): a chain is sequential, so the number of chains decides how full the one wave is
mkdir -p gpurun_out
for p in 32 24 20 19 16; do
  timeout 300 python bench.py --workload c3 --dc-points $p --steps 20 --warmup 3 --no-cpu > gpurun_out/c3_p$p.json 2> gpurun_out/c3_p$p.err
  python -c "
import json;b=json.load(open('gpurun_out/c3_p$p.json'));print($p,b['value'],b['ms_per_step'],b['e2e']['value'],b['config']['variant'][:40],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'])"
done
