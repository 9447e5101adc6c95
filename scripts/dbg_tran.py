import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from ftspice_b200.engine import BatchEngine
from harness import load, golden, fl
np.set_printoptions(precision=17, linewidth=200)
name = sys.argv[1]; stop = float(sys.argv[2]) if len(sys.argv) > 2 else None
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 1
g = golden()[name]["tran"]
circ = load(name); tr = circ.command("tran")
be = BatchEngine(circ); be.set_option("skip_refactor", skip); be.set_params()
cap = 600
r = be.run_tran(stop or tr.stop, tr.step, capacity=cap)
k = min(int(r.n_rec[0]), cap)
print(name, "status", r.status[0], "nrec", r.n_rec[0], be.counters())
x = fl(g["x"]); t = fl(g["t"]); it = np.array(g["n_iters"])
for i in range(min(k, len(t), 40)):
    ok = np.all(np.abs(r.x[0,i]-x[i]) <= 1e-9*np.abs(x[i])+1e-12)
    print(i, "t", r.t[0,i], t[i], "it", r.n_iters[0,i], it[i], "OK" if ok else "BAD", r.x[0,i], x[i] if not ok else "")
