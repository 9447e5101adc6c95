#!/bin/bash
# A/B of the damping's ln_1p in the generated kernels (jit_log1p=0: CUDA's log1p) + the workloads the built kernels serve
cd "$(dirname "$0")/.."
run() { timeout 600 python bench.py --no-cpu --steps 5 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*'.ljust(60), d['config']['variant'], '| ms/step %.3f'%d['ms_per_step'], 'Mpts/s %.3f'%(d['value']/1e6), 'e2e %.3f'%(d['e2e']['value']/1e6), 'ok', d['config']['converged_instances'], 'newton', d['roofline']['newton_iters_per_launch'])" || echo "FAILED: $*"; }
run --workload c2
run --workload c2 --opt jit_log1p=0
run --workload c3
run --workload c3 --opt jit_log1p=0
run --workload c4 --scale 0.25 --tran-stop 0.27e-6
run --workload c5 --scale 0.125 --tran-stop 2e-9
