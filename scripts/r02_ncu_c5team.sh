#!/bin/bash
mkdir -p gpurun_out
C5="python bench.py --workload c5 --scale 0.09 --tran-stop 3e-9 --steps 1 --warmup 3 --no-cpu"
$C5 > gpurun_out/w_c5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_team -s 3 -c 1 -o gpurun_out/w_c5_team $C5 > gpurun_out/w_c5_f.log 2>&1
tail -c 300 gpurun_out/w_c5_f.log
