"""Times the start-up .op of the C5 chain (n = 258) alone: usage dbg_c5_op.py [instances]"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from ftspice_b200 import netlist as nl
from ftspice_b200.engine import BatchEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
circ = nl.load(nl.inverter_chain_netlist(256, stop="168n", step="50p").replace(".TRAN 168n 50p", ".OP"))
vals, fns = nl.monte_carlo(circ, B)
be = BatchEngine(circ)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    be.set_option(k, int(v))
be.set_params(vals, fns)
for rep in range(3):
    t0 = time.perf_counter()
    r = be.run_op()
    dt = time.perf_counter() - t0
    c = be.counters()
    import ctypes as ct
    from ftspice_b200 import capi
    why = (ct.c_int64 * 3)()
    capi.lib().ftsb_debug_why.argtypes = [ct.c_void_p, ct.c_void_p]
    capi.lib().ftsb_debug_why(be.h, why)
    c["why_pivot_fill_nonfinite"] = list(why)
    print(json.dumps({"instances": B, "seconds": dt, "ok": int((r.status == 0).sum()), "iters_mean": float(r.n_iters.mean()), "variant": be.variant, "counters": c}))
