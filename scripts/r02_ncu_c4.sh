#!/bin/bash
# ncu capture of the per-team .tran kernel on a C4 slice (2 048 instances, 0.1 us): plain run first, then launch list + full set
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --scale 0.03125 --tran-stop 0.1e-6 --steps 1 --warmup 3 --no-cpu $EXTRA"
$CMD > gpurun_out/n_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 8 --csv --log-file gpurun_out/n_launches.csv $CMD > gpurun_out/n_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_team -s 3 -c 1 -o gpurun_out/n_team $CMD > gpurun_out/n_f.log 2>&1
tail -2 gpurun_out/n_plain.log; tail -3 gpurun_out/n_f.log
