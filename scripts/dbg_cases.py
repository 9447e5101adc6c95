"""Debug helper: run one analysis of one golden netlist on the GPU and print a summary."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from ftspice_b200.engine import BatchEngine
from ftspice_b200 import capi
from harness import load, golden, fl

name, an = sys.argv[1], sys.argv[2]
g = golden()[name][an]
circ = load(name)
be = BatchEngine(circ); be.set_params()
t0 = time.time()
if an == "dc":
    d = circ.command("dc")
    r = be.run_dc(d.source, d.start, d.stop, d.step)
    k = int(r.points_done[0])
    x = fl(g["x"])
    print(name, an, "status", r.status[0], g["status"], "points", k, len(g["n_iters"]), "iters eq",
          (r.n_iters[0,:k] == np.array(g["n_iters"][:k])).all() if k else None, be.variant, be.counters(), "%.3fs" % (time.time()-t0))
    if k:
        err = np.abs(r.x[0,:k]-x[:k]); tol = 1e-9*np.abs(x[:k])+1e-12
        print("   bad", (err>tol).sum(), "max rel %.2e" % (err/(np.abs(x[:k])+1e-12)).max())
else:
    tr = circ.command("tran")
    cap = max(g["n_rec"],1)+8
    r = be.run_tran(tr.stop, tr.step, capacity=cap)
    k = int(r.n_rec[0])
    print(name, an, "status", r.status[0], g["status"], "nrec", k, g["n_rec"], be.variant, be.counters(), "%.3fs" % (time.time()-t0))
    kk = min(k, g["n_rec"], cap)
    if kk:
        x = fl(g["x"]); t = fl(g["t"])
        print("   iters eq", (r.n_iters[0,:kk] == np.array(g["n_iters"][:kk])).all(), "t eq", (r.t[0,:kk]==t[:kk]).all())
        err = np.abs(r.x[0,:kk]-x[:kk]); tol = 1e-9*np.abs(x[:kk])+1e-12
        print("   bad", (err>tol).sum(), "max rel %.2e" % (err/(np.abs(x[:kk])+1e-12)).max())
        ne = np.argwhere(r.n_iters[0,:kk] != np.array(g["n_iters"][:kk]))
        if len(ne): print("   first iter mismatch at rec", ne[0], r.n_iters[0,ne[0][0]], g["n_iters"][ne[0][0]], "t", r.t[0,ne[0][0]], t[ne[0][0]])
