#!/bin/bash
# ncu of the per-team kernel (levels) on a C5 slice
mkdir -p gpurun_out
CMD="python bench.py --workload c5 --scale 0.11 --tran-stop 3e-9 --steps 1 --warmup 3 --no-cpu --opt team_place=6"
$CMD > gpurun_out/g24_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_team -s 3 -c 1 -o gpurun_out/g24_team $CMD > gpurun_out/g24_f.log 2>&1
tail -1 gpurun_out/g24_plain.log | cut -c1-200; tail -3 gpurun_out/g24_f.log
