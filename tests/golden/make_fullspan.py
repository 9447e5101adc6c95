"""Generates tests/golden/fullspan_c4.json and fullspan_c5.json: ORACLE results of sampled instances of the two named
.tran configs at their FULL time span (C4 ladder 5.4 us / 1 ns, C5 inverter chain 168 ns / 50 ps).

    python tests/golden/make_fullspan.py [c4|c5] [threads]

Per sampled instance (global Monte-Carlo index, the same draw the GPU batch makes): status, number of accepted steps,
Newton / PLTE rejection counts, total Newton iterations, SHA-256 of the time stamps (f64 bytes) and of the per-record
iteration counts (i32 bytes) — both must be reproduced bit for bit — and, for the value tolerance (1e-9 relative +
1e-12: 13 significant digits are stored), x at two (C4) / one (C5) records inside the run and at the last one.  About 4 s (C4) / 30 s (C5) of CPU per instance.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from ftspice_b200 import netlist as nl  # noqa: E402
from oracle import orc  # noqa: E402

CONFIGS = {
    # name: (netlist, stop, step, capacity, global Monte-Carlo indices of the sampled instances)
    "c4": (lambda: nl.ladder_netlist(64, "SIN( 0.0 1.0 10M )", stop="5.4u", step="1n"), 5.4e-6, 1e-9, 12000,
           [0] + [1 + k * 1039 + (k % 7) for k in range(63)]),
    "c5": (lambda: nl.inverter_chain_netlist(256, stop="168n", step="50p"), 168e-9, 50e-12, 7000,
           [0] + [1 + k * 409 + (k % 5) for k in range(39)]),
}


def sample_params(circ, idx):
    vs, fs = [], []
    for i in idx:
        v, f = nl.monte_carlo(circ, 1, first=int(i))
        vs.append(v[0])
        fs.append(f[0])
    return np.array(vs), np.array(fs)


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c4"
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    text_fn, stop, step, cap, idx = CONFIGS[which]
    orc.build()
    circ = nl.load(text_fn())
    vals, fns = sample_params(circ, idx)
    st, nrec, it, t, x, xf, stats = orc.Oracle(circ).batch_tran(vals, fns, stop, step, cap, threads=threads)
    out = {"config": which, "stop": stop, "step": step, "n": circ.n_run, "instances": []}
    for j, i in enumerate(idx):
        k = int(nrec[j])
        assert k <= cap, (i, k)
        rows = (sorted({k // 3, (2 * k) // 3}) if which == "c4" else [k // 2]) if k > 3 else []
        out["instances"].append({
            "index": int(i), "status": int(st[j]), "n_rec": k, "rej_newton": int(stats[j].rej_newton),
            "rej_plte": int(stats[j].rej_plte), "newton_iters": int(it[j, :k].sum()),
            "t_sha256": digest(t[j, :k]), "n_iters_sha256": digest(it[j, :k].astype(np.int32)),
            "t_last": repr(float(t[j, k - 1])) if k else None,
            "rows": rows, "x_rows": [["%.13g" % float(v) for v in x[j, r]] for r in rows],
            "x_final": ["%.13g" % float(v) for v in xf[j]],
        })
    path = os.path.join(HERE, f"fullspan_{which}.json")
    with open(path, "w") as f:
        json.dump(out, f)
    ok = int((st == 0).sum())
    print(f"{which}: {len(idx)} instances, {ok} OK, {len(idx) - ok} failing; accepted steps {int(nrec.sum())}; wrote {path} "
          f"({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
