#!/bin/bash
# launch list of the default bench command (final build): plain run first, then the gpu__time_duration pass
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/ll_plain.json 2> gpurun_out/ll_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ll_launches.csv $CMD > gpurun_out/ll_ncu.log 2>&1
echo "rc=$?"; wc -l gpurun_out/ll_launches.csv
