#!/bin/bash
# C4 A/B line of the round's experiments: parity subset of the per-team kernels, then the C4 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_q9_isource.py -m gpu -x -q -k "mid_size or team or q9 or current_source" > gpurun_out/c4ab_tests.log 2>&1
echo "rc=$?" >> gpurun_out/c4ab_tests.log; tail -n 3 gpurun_out/c4ab_tests.log
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu $EXTRA > gpurun_out/c4ab.json 2> gpurun_out/c4ab.err
echo "rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/c4ab.json'));print(b['value'],b['ms_per_step'],b['config']['variant'][:50],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'])"
