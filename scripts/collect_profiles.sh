#!/bin/bash
# Collects the bench lines and ncu captures committed under profiles/ (run under gpurun; writes gpurun_out/).
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py > $O/bench_c2.json 2> $O/bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_c2_reference.json 2>> $O/bench_c2.err
python bench.py --opt jit=0 --opt plan=0 --no-cpu > $O/bench_c2_dense.json 2>> $O/bench_c2.err
python bench.py --workload c3 --steps 5 > $O/bench_c3.json 2> $O/bench_c3.err
python bench.py --workload c4 --scale 0.25 --tran-stop 0.54e-6 --steps 3 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py --workload c5 --scale 0.18066 --tran-stop 4e-9 --steps 2 > $O/bench_c5.json 2> $O/bench_c5.err
# ncu: launch list + full capture of the C2 kernel (default bench command, short)
C2="python bench.py --steps 2 --warmup 3 --no-cpu"
$C2 > $O/ncu_c2_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r01_c2_jit_launches.csv $C2 > $O/ncu_c2_l.log 2>&1
$C2 > $O/ncu_c2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ftsb_jit -s 3 -c 1 -o $O/r01_c2_jit $C2 > $O/ncu_c2_f.log 2>&1
ls -la $O | tail -20
