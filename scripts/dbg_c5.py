"""quick timing + counters of one synthetic workload through BatchEngine (own measurements)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from ftspice_b200 import netlist as nl
from ftspice_b200.engine import BatchEngine

which = sys.argv[1] if len(sys.argv) > 1 else "c5"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
stop = float(sys.argv[3]) if len(sys.argv) > 3 else 2e-9
opts = dict(kv.split("=") for kv in sys.argv[4:])
if which == "c5":
    circ = nl.load(nl.inverter_chain_netlist(256, stop="168n", step="50p")); step = 50e-12
else:
    circ = nl.load(nl.ladder_netlist(64, "SIN( 0.0 1.0 10M )", stop="5.4u", step="1n")); step = 1e-9
vals, fns = nl.monte_carlo(circ, B)
be = BatchEngine(circ)
for k, v in opts.items():
    be.set_option(k, int(v))
be.set_params(vals, fns)
for rep in range(2):
    t0 = time.perf_counter(); r = be.run_op(); t1 = time.perf_counter()
    print("op  ", be.variant, "%.2f ms" % (1e3 * (t1 - t0)), be.counters(), "ok", int((r.status == 0).sum()))
    t0 = time.perf_counter(); r = be.run_tran(stop, step, capacity=0); t1 = time.perf_counter()
    c = be.counters()
    print("tran", be.variant, "%.2f ms" % (1e3 * (t1 - t0)), c, "ok", int((r.status == 0).sum()),
          "us/newton/instance-slot %.1f" % (1e3 * (t1 - t0) * 1e3 / max(c["newton_iters"], 1) * min(B, 148 * 3)))
