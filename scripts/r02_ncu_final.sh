#!/bin/bash
# final ncu evidence of round 2: C4 per-team kernel (full set + launch list), C5 k_spw<tran> with / without the L2 window
mkdir -p gpurun_out
bash scripts/r02_ncu_c4.sh > /dev/null 2>&1
C5="python bench.py --workload c5 --scale 0.09 --tran-stop 3e-9 --steps 1 --warmup 3 --no-cpu"
$C5 > gpurun_out/v_c5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_spw -s 7 -c 1 -o gpurun_out/v_c5_persist $C5 > gpurun_out/v_c5_f1.log 2>&1
ncu --set full --clock-control none -k regex:k_spw -s 7 -c 1 -o gpurun_out/v_c5_nopersist $C5 --opt l2_persist=0 > gpurun_out/v_c5_f0.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 8 --csv --log-file gpurun_out/v_c5_launches.csv $C5 > gpurun_out/v_c5_l.log 2>&1
ls -la gpurun_out/v_* gpurun_out/n_*
