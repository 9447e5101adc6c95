#!/bin/bash
# Collects the bench lines and ncu captures committed under profiles/ (run under gpurun; writes gpurun_out/).
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py > $O/bench_c2.json 2> $O/bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_c2_reference.json 2>> $O/bench_c2.err
python bench.py --workload c3 --steps 5 > $O/bench_c3.json 2> $O/bench_c3.err
python bench.py --workload c4 --scale 0.25 --tran-stop 0.54e-6 --steps 3 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py --workload c5 --scale 0.125 --tran-stop 2e-9 --steps 3 > $O/bench_c5.json 2> $O/bench_c5.err
# ncu: launch list + full capture of the C4 .tran kernel (small command: 2048 instances, 0.1 us)
C4="python bench.py --workload c4 --scale 0.03125 --tran-stop 0.1e-6 --steps 1 --warmup 3 --no-cpu"
$C4 > $O/ncu_c4_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r01_c4_spw_launches.csv $C4 > $O/ncu_c4_l.log 2>&1
$C4 > $O/ncu_c4_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_spw -s 7 -c 1 -o $O/r01_c4_spw $C4 > $O/ncu_c4_f.log 2>&1
ls -la $O | tail -20
