#!/bin/bash
# usage: scripts/bench_opts.sh <workload> "<opts1>" "<opts2>" ...   (each opts = e.g. "--opt lanes=2")
cd "$(dirname "$0")/.."
wl=$1; shift
for o in "$@"; do
  timeout 200 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu $o 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$o'.ljust(28), d['config']['variant'].ljust(18), 'ms/step %.3f'%d['ms_per_step'], 'Mpts/s %.2f'%(d['value']/1e6), 'e2e %.2f'%(d['e2e']['value']/1e6), 'frac %.4f'%d['roofline']['frac'], 'ok', d['config']['converged_instances'])" || echo "FAILED: $o"
done
