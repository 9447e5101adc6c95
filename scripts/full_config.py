"""One pass of a named config at its FULL size through the C-ABI with device-resident parameters (own measurement;
bench.py is the driver's contract).  usage: full_config.py c4|c5"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from ftspice_b200 import netlist as nl
from ftspice_b200.engine import BatchEngine

which = sys.argv[1]
if which == "c4":
    circ = nl.load(nl.ladder_netlist(64, "SIN( 0.0 1.0 10M )", stop="5.4u", step="1n")); B, stop, step = 65536, 5.4e-6, 1e-9
else:
    circ = nl.load(nl.inverter_chain_netlist(256, stop="168n", step="50p")); B, stop, step = 16384, 168e-9, 50e-12
vals, fns = nl.monte_carlo(circ, B)
be = BatchEngine(circ)
dev = torch.device("cuda", 0)
d_vals = torch.from_numpy(vals).to(dev)
d_fns = torch.from_numpy(np.ascontiguousarray(be.compact_fns(fns))).to(dev)
be.set_params_device(B, d_vals.data_ptr(), d_fns.data_ptr())
d_x = torch.empty((B, circ.n_run), dtype=torch.float64, device=dev)
d_nrec = torch.empty(B, dtype=torch.int64, device=dev)
d_st = torch.empty(B, dtype=torch.int32, device=dev)
be.sync()
t0 = time.perf_counter()
be.run_tran_device(stop, step, 0, 1, 0, 0, 0, d_nrec.data_ptr(), d_x.data_ptr(), d_st.data_ptr())
be.sync()
dt = time.perf_counter() - t0
pts = int(d_nrec.sum().item())
print(json.dumps({"config": which, "instances": B, "tran_stop": stop, "seconds": dt, "solve_points": pts,
                  "solve_points_per_s": pts / dt, "ok_instances": int((d_st == 0).sum().item()), "variant": be.variant,
                  "counters": be.counters()}))
