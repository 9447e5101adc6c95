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

# Arithmetic-sensitivity variants of the oracle.  Two rebuilds of the SAME source: (a) with FMA contraction
# (what rustc would produce with a different codegen choice / what nvcc does by default), (b) with exp / expm1 /
# log1p taken from the long-double libm and rounded (another high-quality libm).  Where the oracle's own result
# moves by more than the parity tolerance between these builds, the reference's answer is not reproducible to
# 1e-9 across compilers/libms; those entries are stored as `spread` and the parity tests use them as an explicit,
# counted allowance (see tests/harness.py).
LD_WRAPPER = r"""
#include <math.h>
static double exp_ld(double x){ return (double)expl((long double)x); }
static double expm1_ld(double x){ return (double)expm1l((long double)x); }
static double log1p_ld(double x){ return (double)log1pl((long double)x); }
#define exp exp_ld
#define expm1 expm1_ld
#define log1p log1p_ld
#include "%s"
"""


def build_variants(tmp):
    import subprocess
    src = os.path.join(ROOT, "oracle", "ftsp_oracle.c")
    inc = os.path.join(ROOT, "oracle")
    fma = os.path.join(tmp, "liborc_fma.so")
    ld = os.path.join(tmp, "liborc_ld.so")
    subprocess.run(["gcc", "-O2", "-ffp-contract=fast", "-mfma", "-fPIC", "-shared", "-I", inc, "-o", fma, src, "-lm",
                    "-lpthread"], check=True)
    wrap = os.path.join(tmp, "orc_ld.c")
    with open(wrap, "w") as f:
        f.write(LD_WRAPPER % src)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", inc, "-o", ld, wrap, "-lm", "-lpthread"],
                   check=True)
    return [fma, ld]


def run_netlist(c):
    """(op, dc, tran) raw oracle results with the currently selected oracle library."""
    o = orc.Oracle(c)
    res = {}
    if c.command("op"):
        res["op"] = o.run_op()
    dc = c.command("dc")
    if dc:
        res["dc"] = o.run_dc(c.elem_index(dc.source), dc.start, dc.stop, dc.step)
    tr = c.command("tran")
    if tr:
        res["tran"] = o.run_tran(tr.stop, tr.step)
    return res


def xs_of(an, r):
    import numpy as np
    if an == "op":
        return np.atleast_2d(r[2]), np.array([r[1]])
    if an == "dc":
        return r[3], r[2]
    return r[4], r[2]


def main():
    import tempfile
    import numpy as np
    out = {}
    texts = {os.path.basename(p)[:-3]: open(p).read() for p in sorted(glob.glob(os.path.join(REF, "*.sp")))}
    # sensitivity pass first (switches the oracle library), then the strict pass
    spread = {}
    with tempfile.TemporaryDirectory() as tmp:
        variants = build_variants(tmp)
        base = {n: run_netlist(nl.load(t)) for n, t in texts.items()}
        for lib in variants:
            orc._lib = None
            orc._LIB_PATH = lib
            for n, t in texts.items():
                other = run_netlist(nl.load(t))
                for an, r in base[n].items():
                    x0, it0 = xs_of(an, r)
                    x1, it1 = xs_of(an, other[an])
                    if x0.shape != x1.shape or not np.array_equal(it0, it1):
                        raise SystemExit(f"{n} {an}: iteration path changes under {lib}")
                    d = np.abs(x1 - x0)
                    spread[(n, an)] = np.maximum(spread.get((n, an), 0.0), d)
        orc._lib = None
        orc._LIB_PATH = os.path.join(ROOT, "oracle", "libftsp_oracle.so")
    for name, text in texts.items():
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
        for an in ("op", "dc", "tran"):
            if an in ent and (name, an) in spread and ent[an]["status"] == 0:
                sp = spread[(name, an)]
                x0, _ = xs_of(an, base[name][an])
                idx = np.argwhere(sp > 0.1 * (1e-9 * np.abs(x0) + 1e-12))
                ent[an]["spread"] = [[int(i), int(j), repr(float(sp[i, j]))] for i, j in idx]
        out[name] = ent
    with open(os.path.join(HERE, "reference_netlists.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", len(out), "netlists")


if __name__ == "__main__":
    main()
