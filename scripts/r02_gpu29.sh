#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mid_size or sparse_kernel or tran_inverter or inverter or c5" > gpurun_out/g29_tests.log 2>&1
echo "rc=$?" >> gpurun_out/g29_tests.log; tail -n 3 gpurun_out/g29_tests.log
timeout 900 python bench.py --workload c5 --scale 0.25 --tran-stop 4e-9 --steps 2 --warmup 3 --no-cpu > gpurun_out/g29_1.json 2> gpurun_out/g29_1.err
echo "rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/g29_1.json'));print(b['value'],b['ms_per_step'],b['config']['variant'][:60],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'],b['config']['converged_instances'])"
