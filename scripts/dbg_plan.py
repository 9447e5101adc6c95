import sys, time
import numpy as np
sys.path.insert(0, ".")
from ftspice_b200 import netlist as nl
from ftspice_b200.engine import BatchEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
circ = nl.load(nl.PROJ1_NL)
vals, fns = nl.monte_carlo(circ, B, first=first)
for opts in ({}, {"jit": 0}, {"plan": 0}):
    be = BatchEngine(circ)
    for k, v in opts.items():
        be.set_option(k, v)
    be.set_params(vals, fns)
    for rep in range(3):
        t0 = time.perf_counter(); r = be.run_op(); t1 = time.perf_counter()
        print(opts, be.variant, "%.2f ms" % (1e3 * (t1 - t0)), be.counters(), int((r.status == 0).sum()))
