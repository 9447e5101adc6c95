#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --scale 0.1084 --tran-stop 0.2e-6 --steps 1 --warmup 3 --no-cpu"
$CMD > gpurun_out/g18_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_team -s 3 -c 1 -o gpurun_out/g18_team $CMD > gpurun_out/g18_f.log 2>&1
tail -1 gpurun_out/g18_plain.log | cut -c1-200; tail -3 gpurun_out/g18_f.log
