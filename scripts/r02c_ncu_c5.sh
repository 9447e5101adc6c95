#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mid_size or sparse_kernel or tran_inverter or inverter or c5" > gpurun_out/g28_tests.log 2>&1
echo "rc=$?" >> gpurun_out/g28_tests.log; tail -n 3 gpurun_out/g28_tests.log
CMD="python bench.py --workload c5 --scale 0.11 --tran-stop 3e-9 --steps 1 --warmup 3 --no-cpu"
$CMD > gpurun_out/g28_c5_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 12 --csv --log-file gpurun_out/g28_c5_launches.csv $CMD > gpurun_out/g28_l5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_team -s 3 -c 1 -o gpurun_out/g28_c5_team $CMD > gpurun_out/g28_f5.log 2>&1
tail -1 gpurun_out/g28_c5_plain.log | cut -c1-160
