#!/bin/bash
# rows kernel, one 14-warp block per SM: C4 bench, then ncu of one full wave (8 192 instances x 0.05 us)
mkdir -p gpurun_out
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/g13_c4.json 2> gpurun_out/g13_c4.err
echo "bench c4 rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/g13_c4.json'));print(b['value'],b['ms_per_step'],b['config']['variant'],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'])"
CMD="python bench.py --workload c4 --scale 0.125 --tran-stop 0.05e-6 --steps 1 --warmup 3 --no-cpu"
$CMD > gpurun_out/g13_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 8 --csv --log-file gpurun_out/g13_launches.csv $CMD > gpurun_out/g13_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_team -s 3 -c 1 -o gpurun_out/g13_team $CMD > gpurun_out/g13_f.log 2>&1
tail -2 gpurun_out/g13_plain.log | cut -c1-300; tail -3 gpurun_out/g13_f.log
