#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/full_config.py c5 > gpurun_out/g30_full_c5.json 2> gpurun_out/g30_full_c5.err; echo "full c5 rc=$?"; cut -c1-300 gpurun_out/g30_full_c5.json
CMD="python bench.py --workload c5 --scale 0.11 --tran-stop 3e-9 --steps 1 --warmup 3 --no-cpu"
$CMD > gpurun_out/g30_c5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_team -s 3 -c 1 -o gpurun_out/g30_c5_team $CMD > gpurun_out/g30_f5.log 2>&1
tail -1 gpurun_out/g30_c5_plain.log | cut -c1-160
