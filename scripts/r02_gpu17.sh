#!/bin/bash
# slot streams in the rows kernel: parity subset, then C4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_q9_isource.py -m gpu -x -q -k "mid_size or team or q9 or current_source" > gpurun_out/g17_tests.log 2>&1
echo "rc=$?" >> gpurun_out/g17_tests.log; tail -n 5 gpurun_out/g17_tests.log
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu $opts > gpurun_out/g17_1.json 2> gpurun_out/g17_1.err
echo "rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/g17_1.json'));print(b['value'],b['ms_per_step'],b['config']['variant'][:50],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'])"
