"""Generates tests/golden/reference_netlists.json.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py

For each of the reference's 17 example netlists (test/*.sp — the inputs the north_star names as the parity
set) it stores the netlist text and the ORACLE's result for every analysis the netlist asks for.  The
reference itself cannot be run here (Rust, no cargo), so these are oracle outputs, cross-checked against the
independent restatement of SURVEY.md App. A (iteration counts, record counts, rejections, spot values).
The GPU box has no /root/reference: tests read this file only.
"""
import glob
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from ftspice_b200 import netlist as nl  # noqa: E402
from oracle import orc  # noqa: E402

REF = "/root/reference/test"


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(REF, "*.sp"))):
        name = os.path.basename(path)[:-3]
        text = open(path).read()
        c = nl.load(text)
        o = orc.Oracle(c)
        ent = {"netlist": text, "n_run": c.n_run, "n_start": c.n_start}
        if c.command("op"):
            s, it, x, st = o.run_op()
            ent["op"] = {"status": s, "n_iters": it, "headers": c.headers(True),
                         "x": [repr(float(v)) for v in x], "lu_solves": st.lu_solves}
        dc = c.command("dc")
        if dc:
            s, npts, it, x, sw, st = o.run_dc(c.elem_index(dc.source), dc.start, dc.stop, dc.step)
            ent["dc"] = {"status": s, "points": npts, "n_iters": [int(v) for v in it], "headers": c.headers(False),
                         "sweep": [repr(float(v)) for v in sw],
                         "x": [[repr(float(v)) for v in row] for row in x], "newton_iters": st.newton_iters}
        tr = c.command("tran")
        if tr:
            s, nrec, it, t, x, xf, st = o.run_tran(tr.stop, tr.step)
            ent["tran"] = {"status": s, "n_rec": nrec, "n_iters": [int(v) for v in it], "headers": c.headers(False),
                           "t": [repr(float(v)) for v in t], "x": [[repr(float(v)) for v in row] for row in x],
                           "newton_iters": st.newton_iters, "rej_newton": st.rej_newton, "rej_plte": st.rej_plte}
        out[name] = ent
    with open(os.path.join(HERE, "reference_netlists.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", len(out), "netlists")


if __name__ == "__main__":
    main()
