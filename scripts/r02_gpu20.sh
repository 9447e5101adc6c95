#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_q9_isource.py tests/test_fullspan.py -m gpu -x -q -k "mid_size or team or q9 or current_source or ladder or c4 or tran" > gpurun_out/g20_tests.log 2>&1
echo "rc=$?" >> gpurun_out/g20_tests.log; tail -n 4 gpurun_out/g20_tests.log
for opts in "" "--opt team_rows=0"; do
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu $opts > gpurun_out/g20_1.json 2> gpurun_out/g20_1.err
echo "[$opts] rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/g20_1.json'));print(b['value'],b['ms_per_step'],b['config']['variant'][:50],b['roofline']['newton_iters_per_launch'],b['config']['solve_points_per_step'])"
done
