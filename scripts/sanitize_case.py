"""tiny sparse-kernel cases for compute-sanitizer (memcheck / racecheck): ladder .tran, inverter chain .tran, .op"""
import sys
sys.path.insert(0, ".")
import numpy as np
from ftspice_b200 import netlist as nl
from ftspice_b200.engine import BatchEngine

for text, B in [(nl.ladder_netlist(24, "PULSE( 0.0 1.0 0.0 10n 10n 50n 100n )", stop="3n", step="1n"), 5),
                (nl.inverter_chain_netlist(40, stop="0.2n", step="50p"), 3),
                (nl.PROJ1_NL, 70)]:
    circ = nl.load(text)
    vals, fns = nl.monte_carlo(circ, B, rel=0.3)
    for opts in ({}, {"fill_cap": 2}, {"sparse": 1}):
        be = BatchEngine(circ)
        for k, v in opts.items():
            be.set_option(k, v)
        be.set_params(vals, fns)
        t = circ.command("tran")
        if t:
            r = be.run_tran(t.stop, t.step, capacity=64)
            print(be.variant, r.status.tolist(), r.n_rec.tolist(), be.counters()["redo"])
        else:
            r = be.run_op()
            print(be.variant, int((r.status == 0).sum()), be.counters()["redo"])
