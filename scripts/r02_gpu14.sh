#!/bin/bash
# C4: lanes per instance x placement of A for the compile-time-rows kernel
mkdir -p gpurun_out
i=0
for opts in "" "--opt team_place=4" "--opt team_lanes=4" "--opt team_lanes=4 --opt team_place=4" "--opt team_rows=0 --opt team_lanes=4"; do
  i=$((i+1))
  timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu $opts > gpurun_out/g14_$i.json 2> gpurun_out/g14_$i.err
  echo "[$opts] rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/g14_$i.json'));print(b['value'],b['ms_per_step'],b['config']['variant'][:50],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'])"
done
