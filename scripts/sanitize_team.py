"""tiny per-team-kernel cases for compute-sanitizer (memcheck / racecheck): the compile-time-rows kernel (ladders), the level-scheduled
placements 3 and 6 (inverter chains), the generic loop"""
import sys
sys.path.insert(0, ".")
import numpy as np
from ftspice_b200 import netlist as nl
from ftspice_b200.engine import BatchEngine

cases = [(nl.ladder_netlist(24, "PULSE( 0.0 1.0 0.0 10n 10n 50n 100n )", stop="3n", step="1n"), 9, [{}, {"team_rows": 0}]),
         (nl.ladder_netlist(64, "SIN( 0.0 1.0 10M )", stop="3n", step="1n"), 5, [{}]),
         (nl.inverter_chain_netlist(40, stop="0.3n", step="50p"), 5, [{}, {"team_levels": 0}, {"team_lanes": 32, "team_place": 6}]),
         (nl.inverter_chain_netlist(256, stop="0.15n", step="50p"), 2, [{}])]
for text, B, optsets in cases:
    circ = nl.load(text)
    vals, fns = nl.monte_carlo(circ, B, rel=0.1)
    for opts in optsets:
        be = BatchEngine(circ)
        for k, v in opts.items():
            be.set_option(k, v)
        be.set_params(vals, fns)
        t = circ.command("tran")
        r = be.run_tran(t.stop, t.step, capacity=64)
        print(be.variant.split("+")[0], r.status.tolist(), r.n_rec.tolist(), be.counters()["redo"])
