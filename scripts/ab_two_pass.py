"""A/B of the two-pass .op step (jit_two_pass) on the reference's nonlinear netlists, 65 536 Monte-Carlo instances each
(own measurement; device-resident parameters, CUDA events around run_op)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from ftspice_b200 import netlist as nl
from ftspice_b200.engine import BatchEngine
sys.path.insert(0, "tests")
from harness import load

dev = torch.device("cuda", 0)
B = 65536
for name in ("proj1_nl", "npn_test", "proj2", "r_d_direct", "nmos_test"):
    circ = nl.load(nl.PROJ1_NL) if name == "proj1_nl" else load(name)
    vals, fns = nl.monte_carlo(circ, B)
    d_vals = torch.from_numpy(vals).to(dev)
    d_x = torch.empty((B, circ.n_start), dtype=torch.float64, device=dev)
    d_it = torch.empty(B, dtype=torch.int32, device=dev)
    d_st = torch.empty(B, dtype=torch.int32, device=dev)
    out = []
    for k in (0, 4, 8, 12):
        be = BatchEngine(circ)
        be.set_option("jit_two_pass", k)
        be.set_params_device(B, d_vals.data_ptr(), 0)
        for _ in range(4):
            be.run_op_device(d_x.data_ptr(), d_it.data_ptr(), d_st.data_ptr())
        be.sync()
        t0 = time.perf_counter()
        for _ in range(10):
            be.run_op_device(d_x.data_ptr(), d_it.data_ptr(), d_st.data_ptr())
        be.sync()
        out.append("k=%d %.3f ms" % (k, 100 * (time.perf_counter() - t0)))
    print(name.ljust(12), be.variant.split("+")[0].ljust(28), "mean iters %.1f" % d_it.float().mean().item(), " | ".join(out))
