"""C2 .op through the per-topology kernel with instance-major (argv[1] = 0) or parameter-major (1) parameters: the same
65 536 instances, four calls (plans learned, kernel generated, then steady state).  Used under ncu to compare the L1
sectors per global load request of the two layouts (scripts/r02_ncu_all.sh)."""
import ctypes as ct
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ftspice_b200 import capi, netlist as nl  # noqa: E402
from ftspice_b200.engine import BatchEngine  # noqa: E402

layout = int(sys.argv[1]) if len(sys.argv) > 1 else 0
circ = nl.load(nl.PROJ1_NL)
B = 65536
vals, fns = nl.monte_carlo(circ, B)
be = BatchEngine(circ)
v = np.ascontiguousarray(vals.T if layout else vals)
for rep in range(4):
    capi.check(capi.lib().ftsb_set_params(be.h, B, v.ctypes.data_as(ct.c_void_p), None, layout, capi.MEM_HOST), be.h)
    be.batch = B
    r = be.run_op()
print(be.variant, int((r.status == 0).sum()), float(np.abs(r.x).sum()))
