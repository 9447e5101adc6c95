#!/bin/bash
# lane-per-instance (G = 1) variants of the team kernel: parity on a short ladder, then the C4 quarter bench
mkdir -p gpurun_out
python - > gpurun_out/g7_parity.log 2>&1 <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from ftspice_b200 import netlist as nl
from ftspice_b200.engine import BatchEngine
for nodes, drive in ((24, "PULSE( 0.0 1.0 0.0 10n 10n 50n 100n )"), (64, "SIN( 0.0 1.0 10M )")):
    circ = nl.load(nl.ladder_netlist(nodes, drive, stop="40n", step="1n"))
    vals, fns = nl.monte_carlo(circ, 70)
    tr = circ.command("tran")
    outs = {}
    for name, opts in (("g8", {}), ("g1p60", {"team_lanes": 1, "team_place": 60}), ("g1p52", {"team_lanes": 1, "team_place": 52}),
                       ("g1p36", {"team_lanes": 1, "team_place": 36})):
        be = BatchEngine(circ)
        for k, v in opts.items():
            be.set_option(k, v)
        be.set_params(vals, fns)
        r = be.run_tran(tr.stop, tr.step, capacity=300)
        outs[name] = (be.variant, r.status, r.n_rec, r.n_iters, r.t, r.x, r.x_final, be.counters())
        print(nodes, name, be.variant.split("+")[0], outs[name][7])
    for name in ("g1p60", "g1p52", "g1p36"):
        same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(outs["g8"][1:7], outs[name][1:7]))
        print(nodes, name, "bitwise equal to g8:", same)
PY
cat gpurun_out/g7_parity.log
for opt in "team_lanes=1 --opt team_place=60" "team_lanes=1 --opt team_place=52" "team_lanes=1 --opt team_place=36"; do
  tag=$(echo "$opt" | tr -c 'a-z0-9=' '_')
  timeout 600 python bench.py --workload c4 --scale 0.25 --tran-stop 0.54e-6 --steps 2 --warmup 3 --no-cpu --opt $opt > gpurun_out/g7_c4_$tag.json 2> gpurun_out/g7_c4_$tag.err
  echo "bench $opt rc=$?"
done
timeout 900 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --opt team_lanes=1 --opt team_place=60 > gpurun_out/g7_c4_full_p60.json 2> gpurun_out/g7_c4_full_p60.err
echo "full p60 rc=$?"
