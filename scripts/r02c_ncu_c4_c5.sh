#!/bin/bash
# final kernels of the round: ncu of the C4 rows kernel (one full wave, 0.4 us) and of the C5 per-team kernel, full-size C4 / C5 passes
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --scale 0.1084 --tran-stop 0.4e-6 --steps 1 --warmup 3 --no-cpu"
$CMD > gpurun_out/g27_c4_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 12 --csv --log-file gpurun_out/g27_c4_launches.csv $CMD > gpurun_out/g27_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_team -s 3 -c 1 -o gpurun_out/g27_c4_team $CMD > gpurun_out/g27_f.log 2>&1
tail -1 gpurun_out/g27_c4_plain.log | cut -c1-160
CMD="python bench.py --workload c5 --scale 0.11 --tran-stop 3e-9 --steps 1 --warmup 3 --no-cpu"
$CMD > gpurun_out/g27_c5_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 12 --csv --log-file gpurun_out/g27_c5_launches.csv $CMD > gpurun_out/g27_l5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_team -s 3 -c 1 -o gpurun_out/g27_c5_team $CMD > gpurun_out/g27_f5.log 2>&1
tail -1 gpurun_out/g27_c5_plain.log | cut -c1-160
timeout 300 python scripts/full_config.py c4 > gpurun_out/g27_full_c4.json 2> gpurun_out/g27_full_c4.err; echo "full c4 rc=$?"; cut -c1-200 gpurun_out/g27_full_c4.json
timeout 600 python scripts/full_config.py c5 > gpurun_out/g27_full_c5.json 2> gpurun_out/g27_full_c5.err; echo "full c5 rc=$?"; cut -c1-300 gpurun_out/g27_full_c5.json
