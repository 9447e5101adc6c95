#!/bin/bash
# quick survey of all workloads at reduced size (own measurements, not the driver's bench line)
cd "$(dirname "$0")/.."
run() { timeout 400 python bench.py --no-cpu --steps 3 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$*'.ljust(52), d['config']['variant'], '| ms/step %.2f'%d['ms_per_step'], 'Mpts/s %.3f'%(d['value']/1e6), 'e2e %.3f'%(d['e2e']['value']/1e6), 'frac %.4f'%r['frac'], 'newton', r['newton_iters_per_launch'], 'factor', r['lu_factor_per_launch'], 'redo', r.get('redo_instances'), 'pts', d['config']['solve_points_per_step'], 'ok', d['config']['converged_instances'])" || echo "FAILED: $*"; }
run --workload c3
run --workload c4 --scale 0.25 --tran-stop 0.27e-6
run --workload c4 --scale 0.25 --tran-stop 0.27e-6 --opt sparse_warp=0
run --workload c5 --scale 0.125 --tran-stop 2e-9
run --workload c5 --scale 0.125 --tran-stop 2e-9 --opt sparse_warp=0
