"""CPU test of the per-topology kernel generator (ftspice_b200/csrc/jit.cpp, plan.cpp): for every reference netlist
with 2..16 unknowns, build plans from a structurally valid pivot sequence, generate the CUDA source and compile it with
NVRTC for sm_100a — no GPU needed (nothing is launched)."""
import ctypes as ct

import numpy as np
import pytest

from ftspice_b200 import capi
from ftspice_b200 import netlist as nl

from harness import load, netlist_names


def _pattern(circ, mode):
    """Structural pattern of the Jacobian in analysis `mode` (rows of column sets), mirroring compile.cpp's stamps."""
    n = circ.n_start if mode == "op" else circ.n_run
    rows = [set() for _ in range(n)]

    def add(i, j):
        if i >= 0 and j >= 0:
            rows[i].add(j)

    for e in range(circ.n_elems):
        k, nd, br = int(circ.kind[e]), [int(x) for x in circ.node[e]], int(circ.branch[e])
        if k in (nl.R, nl.D):
            for i in nd[:2]:
                for j in nd[:2]:
                    add(i, j)
        elif k == nl.V or (k == nl.L and mode == "op"):
            for x in nd[:2]:
                add(br, x)
                add(x, br)
        elif k == nl.Q:
            for i in nd:
                for j in nd:
                    add(i, j)
        elif k == nl.M:
            d, g, s = nd
            for i in (d, s):
                for j in (d, g, s):
                    add(i, j)
    return n, rows


def _greedy_perm(n, rows):
    """A pivot sequence that is structurally possible: at each step the unpivoted row with an entry in the column
    that sits at the smallest position (positions follow the kernel's swap rule)."""
    rows = [set(r) for r in rows]
    pos = list(range(n))
    at = list(range(n))
    perm = 0xFEDCBA9876543210
    for k in range(n):
        cands = [r for r in range(n) if pos[r] >= k and k in rows[r]]
        if not cands:
            return None  # structurally singular in this mode
        P = min(cands, key=lambda r: pos[r])
        rk, pp = at[k], pos[P]
        at[k], at[pp] = P, rk
        pos[P], pos[rk] = k, pp
        cols = {j for j in rows[P] if j > k}
        for r in range(n):
            if pos[r] > k and k in rows[r]:
                rows[r] |= cols
        perm = (perm & ~(0xF << (4 * k))) | (P << (4 * k))
    return perm


@pytest.mark.parametrize("name", netlist_names())
def test_generated_kernel_compiles(name):
    circ = load(name)
    lib = capi.lib()
    fn = lib.ftsb_debug_jit
    fn.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_void_p, ct.c_int, ct.c_void_p, ct.c_void_p, ct.c_int,
                   ct.c_char_p, ct.c_char_p, ct.c_int]
    fn.restype = ct.c_int
    elems = (capi.FtsbElem * max(circ.n_elems, 1))()
    for i in range(circ.n_elems):
        elems[i].kind = int(circ.kind[i])
        for j in range(3):
            elems[i].node[j] = int(circ.node[i, j])
        elems[i].branch = int(circ.branch[i])
        elems[i].fn_kind = int(circ.fn_kind[i])
    cls = np.ascontiguousarray(circ.cls, np.int32)
    compiled = 0
    for an, mode in ((0, "op"), (1, "dc")):
        n, rows = _pattern(circ, mode)
        if not 2 <= n <= 16:
            continue
        perm = _greedy_perm(n, rows)
        if perm is None:
            continue
        perms = np.array([perm], np.uint64)
        counts = np.array([10], np.int64)
        log = ct.create_string_buffer(1 << 16)
        rc = fn(elems, circ.n_elems, circ.n_run, circ.n_start, cls.ctypes.data, an, perms.ctypes.data, counts.ctypes.data, 1,
                None, log, len(log))
        msg = log.value.decode(errors="replace")
        if rc == -3 and "libnvrtc not available" in msg:
            pytest.skip("NVRTC is not installed here")
        assert rc == 0, f"{name} {mode}: rc {rc}: {msg[:2000]}"
        assert msg.startswith("ok K=1"), msg[:200]
        compiled += 1
    if compiled == 0:
        pytest.skip("no analysis mode of this netlist has 2..16 unknowns and a structurally regular matrix")
