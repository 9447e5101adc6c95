#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_q9_isource.py tests/test_multi_device.py -m gpu -x -q -k "mid_size or sparse_kernel or tran_ladder or q9 or current_source or multi or param_major" > gpurun_out/g8_tests.log 2>&1
echo "rc=$?" >> gpurun_out/g8_tests.log; tail -n 3 gpurun_out/g8_tests.log
timeout 600 python bench.py --workload c4 --scale 0.25 --tran-stop 0.54e-6 --steps 2 --warmup 3 --no-cpu > gpurun_out/g8_c4_quarter.json 2> gpurun_out/g8_c4_quarter.err; echo "rc=$?"
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/g8_c4_full.json 2> gpurun_out/g8_c4_full.err; echo "rc=$?"
bash scripts/r02_ncu_c4.sh > /dev/null 2>&1
