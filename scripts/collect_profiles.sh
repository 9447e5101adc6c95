#!/bin/bash
# Collects the bench lines and ncu captures committed under profiles/ (run under gpurun; writes gpurun_out/).
# usage: scripts/collect_profiles.sh [tag]   (tag names the files: r01c -> gpurun_out/r01c_*.json ...)
cd "$(dirname "$0")/.."
O=gpurun_out
T=${1:-r01c}
python bench.py > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_c2_reference.json 2>> $O/${T}_bench_c2.err
python bench.py --workload c3 --steps 5 --no-cpu > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err
python bench.py --workload c4 --scale 0.25 --tran-stop 0.54e-6 --steps 3 --no-cpu > $O/${T}_bench_c4_quarter.json 2> $O/${T}_bench_c4.err
python bench.py --workload c5 --scale 0.18066 --tran-stop 4e-9 --steps 2 --no-cpu > $O/${T}_bench_c5.json 2> $O/${T}_bench_c5.err
# ncu: launch list + full capture of the C2 kernel (default bench command, short).  -s 1: the second ftsb_jit launch is a
# whole-batch launch of the device-buffer warm-up (the host-buffer steps launch one kernel per chunk)
C2="python bench.py --steps 2 --warmup 3 --no-cpu"
$C2 > $O/${T}_ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${T}_c2_jit_launches.csv $C2 > $O/${T}_ncu_l.log 2>&1
$C2 > $O/${T}_ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ftsb_jit -s 1 -c 1 -o $O/${T}_c2_jit $C2 > $O/${T}_ncu_f.log 2>&1
ls -la $O | tail -12
