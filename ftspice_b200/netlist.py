"""Netlist reader, index lowering and synthetic-circuit generators (host side, Python).

This is the harness-level mirror of the reference front end: it turns netlist text into the
index-based descriptor that crosses the C-ABI (``include/ftsb.h``).  It follows

* grammar: ``src/spice.pest:1-59``  (values ``:44-46``, comments ``:54-59``)
* element construction incl. the terminal swaps: ``src/parser.rs:64-219``
* SI prefixes ``src/parser.rs:310-340`` (``value *= mult`` in f64)
* commands ``src/parser.rs:221-252``
* unknown numbering: ``src/node_collection.rs:13-74`` (sorted voltage nodes, then sorted G2 names,
  then — start-up only — sorted inductor names; byte-lexicographic order)
* output column order = BTreeMap key order over all names (``src/engine.rs:79-80``)

No arithmetic of the hot path lives here.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# element kinds / waveform kinds: same numbering as include/ftsb.h and oracle/ftsp_oracle.h
R, C, L, V, I, D, Q, M = range(8)
FN_NONE, FN_SIN, FN_PULSE, FN_EXP = range(4)
KIND_OF_LETTER = {"R": R, "C": C, "L": L, "V": V, "I": I, "D": D, "Q": Q, "M": M}
GND = "0"  # src/node.rs:1

_PREFIX = {  # src/parser.rs:321-334
    "G": 1e9, "M": 1e6, "k": 1e3, "h": 1e2, "da": 1e1, "d": 1e-1, "c": 1e-2,
    "m": 1e-3, "u": 1e-6, "n": 1e-9, "p": 1e-12, "f": 1e-15,
}
# number = "-"? ("0" | [1-9][0-9]*) ("." [0-9]*)? ; prefix tried in grammar order ("da" before "d")
_NUM = r"-?(?:0|[1-9][0-9]*)(?:\.[0-9]*)?"
_VALUE_RE = re.compile(rf"({_NUM})(G|M|k|h|da|d|c|m|u|n|p|f)?")
_NAME_RE = re.compile(r"[A-Za-z0-9]+")


class NetlistError(ValueError):
    pass


def parse_value(tok: str, suffix: str = "") -> float:
    """``value`` (+ an optional literal unit suffix such as V or A), src/parser.rs:310-340."""
    m = _VALUE_RE.match(tok)
    if not m:
        raise NetlistError(f"bad value {tok!r}")
    rest = tok[m.end():]
    num, pre = m.group(1), m.group(2)
    if suffix:
        if rest.upper() != suffix.upper():
            # PEG backtracking: a prefix letter may have eaten nothing useful — only exact forms parse
            raise NetlistError(f"bad value {tok!r} (expected unit {suffix})")
    elif rest:
        raise NetlistError(f"bad value {tok!r}")
    val = float(num)
    if pre:
        val *= _PREFIX[pre]
    return val


@dataclass
class Elem:
    kind: int
    name: str
    nodes: List[str]          # the Rust struct's `nodes` Vec, i.e. AFTER the parser's swaps
    val: float = 0.0
    fn_kind: int = FN_NONE
    fn: Tuple[float, ...] = ()


@dataclass
class Command:
    kind: str                 # "op" | "dc" | "tran"
    source: str = ""
    start: float = 0.0
    stop: float = 0.0
    step: float = 0.0


def _parse_fn(text: str) -> Tuple[int, Tuple[float, ...]]:
    m = re.match(r"(?i)(SIN|PULSE|EXP)\(\s*(.*?)\s*\)\s*$", text)
    if not m:
        raise NetlistError(f"bad function {text!r}")
    kind = {"SIN": FN_SIN, "PULSE": FN_PULSE, "EXP": FN_EXP}[m.group(1).upper()]
    vals = tuple(parse_value(t) for t in m.group(2).split())
    need = {FN_SIN: 3, FN_PULSE: 7, FN_EXP: 6}[kind]
    if len(vals) != need:
        raise NetlistError(f"{m.group(1)} needs {need} values, got {len(vals)}")
    return kind, vals


def spice_fn_eval_t0(fn_kind: int, p: Sequence[float]) -> float:
    """f(0) — the parser stores it as the source value (src/parser.rs:95-99, 126-130).

    Only t = 0 is ever needed on the host; the general evaluation is device code
    (src/spice_fn.rs:8-124)."""
    import math
    t = 0.0
    if fn_kind == FN_SIN:
        return p[0] + p[1] * math.sin(2.0 * math.pi * p[2] * t)
    if fn_kind == FN_PULSE:
        v1, v2, delay, t_rise, t_fall, pw, period = p
        tn = abs(math.fmod(t, period))
        if tn < delay:
            return v1
        if tn < delay + t_rise:
            return v1 + (tn - delay) / t_rise * (v2 - v1)
        if tn < delay + pw - t_fall:
            return v2
        if tn < delay + pw:
            return v2 + (tn - (delay + pw - t_fall)) / t_fall * (v1 - v2)
        return v1
    if fn_kind == FN_EXP:
        v1, v2, rd, rtau, fd, ftau = p
        if t < rd:
            return v1
        if t < fd:
            return v1 + (v1 - v2) * math.expm1(-(t - rd) / rtau)
        return v1 + (v1 - v2) * math.expm1(-(t - rd) / rtau) + (v2 - v1) * math.expm1(-(t - fd) / ftau)
    raise NetlistError("no function")


def parse_netlist(text: str) -> Tuple[List[Elem], List[Command]]:
    """Netlist text → (elements in file order, commands in file order)."""
    elems: List[Elem] = []
    cmds: List[Command] = []
    ended = False
    for raw in text.split("\n"):
        line = raw.strip(" ")
        if not line:
            continue
        if line[0] in "*$":
            continue
        low = line.lower()
        if low.startswith(".end"):
            ended = True
            break
        if low == ".op":
            cmds.append(Command("op"))
            continue
        if low.startswith(".dc"):
            tk = line.split()
            if len(tk) != 5 or tk[1][0].upper() not in "VI":
                raise NetlistError(f"bad .dc: {line!r}")
            cmds.append(Command("dc", tk[1], parse_value(tk[2]), parse_value(tk[3]), parse_value(tk[4])))
            continue
        if low.startswith(".tran"):
            tk = line.split()
            if len(tk) != 3:
                raise NetlistError(f"bad .tran: {line!r}")
            cmds.append(Command("tran", "", 0.0, parse_value(tk[1]), parse_value(tk[2])))
            continue
        letter = line[0].upper()
        if letter not in KIND_OF_LETTER:
            raise NetlistError(f"unknown line {line!r}")
        kind = KIND_OF_LETTER[letter]
        if kind in (V, I):
            m = re.match(r"(\S+) +(\S+) +(\S+) +(.*)$", line)
            if not m:
                raise NetlistError(f"bad source {line!r}")
            name, a, b, vtxt = m.groups()
            vtxt = vtxt.strip(" ")
            if "(" in vtxt:
                fk, fp = _parse_fn(vtxt)
                val = spice_fn_eval_t0(fk, fp)
            else:
                fk, fp = FN_NONE, ()
                val = parse_value(vtxt, "V" if kind == V else "A")
            # parser.rs:83-84,114-115: nodes = [second, first]
            elems.append(Elem(kind, name, [b, a], val, fk, fp))
            continue
        tk = line.split()
        name = tk[0]
        for t in tk[:1]:
            if not _NAME_RE.fullmatch(t):
                raise NetlistError(f"bad name {t!r}")
        if kind in (R, C, L):
            if len(tk) != 4:
                raise NetlistError(f"bad element {line!r}")
            key, _, vtok = tk[3].partition("=")
            if key.upper() != letter or not vtok:
                raise NetlistError(f"bad element value {line!r}")
            val = parse_value(vtok)
            # R keeps netlist order (parser.rs:67-69); C and L store [second, first] (:145-146,161-162)
            nodes = [tk[1], tk[2]] if kind == R else [tk[2], tk[1]]
            elems.append(Elem(kind, name, nodes, val))
        elif kind == D:
            if len(tk) != 4 or tk[3] != "d_model":
                raise NetlistError(f"bad diode {line!r}")
            elems.append(Elem(kind, name, [tk[2], tk[1]]))  # parser.rs:177-178
        else:  # Q, M: 4 netlist nodes, the 4th ignored (parser.rs:186-219)
            model = "q_model" if kind == Q else "t_model"
            if len(tk) != 6 or tk[5] != model:
                raise NetlistError(f"bad transistor {line!r}")
            elems.append(Elem(kind, name, [tk[1], tk[2], tk[3]]))
    if not ended:
        raise NetlistError("missing .end")
    return elems, cmds


def check_elems(elems: Sequence[Elem]) -> None:
    """src/parser/check_elems.rs:6-22."""
    names = {e.name for e in elems}
    if len(names) != len(elems):
        raise NetlistError("Duplicate elems found!")
    if not any(n == GND for e in elems for n in e.nodes):
        raise NetlistError("GND node not found!")


@dataclass
class Circuit:
    """Index-lowered circuit: what crosses the C-ABI plus the names the result writer needs."""
    elems: List[Elem]
    cmds: List[Command]
    n_run: int
    n_start: int
    names: List[str]              # unknown index -> name (start-up numbering; first n_run = run numbering)
    cls: np.ndarray               # [n_start] int32, 0 = Voltage, 1 = Current
    order_run: np.ndarray         # [n_run] unknown indices in BTreeMap key order
    order_start: np.ndarray       # [n_start]
    kind: np.ndarray              # [n_elems] int32
    node: np.ndarray              # [n_elems,3] int32, -1 = GND / unused
    branch: np.ndarray            # [n_elems] int32
    fn_kind: np.ndarray           # [n_elems] int32
    vals: np.ndarray              # [n_elems] float64 (nominal)
    fns: np.ndarray               # [n_elems,7] float64 (nominal)
    index_of: Dict[str, int] = field(default_factory=dict)

    @property
    def n_elems(self) -> int:
        return len(self.elems)

    def elem_index(self, name: str) -> int:
        for i, e in enumerate(self.elems):
            if e.name == name:
                return i
        raise KeyError(name)

    def command(self, kind: str) -> Optional[Command]:
        """First command of a kind (src/engine.rs:32-43)."""
        for c in self.cmds:
            if c.kind == kind:
                return c
        return None

    def headers(self, startup: bool) -> List[str]:
        """Column names in the reference's order (BTreeMap keys), src/engine.rs:79-80,112-113."""
        order = self.order_start if startup else self.order_run
        return [self.names[i] for i in order]


def lower(elems: Sequence[Elem], cmds: Sequence[Command] = ()) -> Circuit:
    """NodeCollection::from_elems / from_startup_elems as index tables (src/node_collection.rs:13-74)."""
    elems = list(elems)
    v_names = sorted({n for e in elems for n in e.nodes if n != GND})
    g2_names = sorted({e.name for e in elems if e.kind == V})
    l_names = sorted({e.name for e in elems if e.kind == L})
    index_of: Dict[str, int] = {}
    for i, n in enumerate(v_names):
        index_of[n] = i
    for i, n in enumerate(g2_names):
        index_of[n] = i + len(v_names)
    n_run = len(v_names) + len(g2_names)
    for i, n in enumerate(l_names):
        index_of[n] = i + n_run
    n_start = n_run + len(l_names)
    names = [""] * n_start
    for k, i in index_of.items():
        names[i] = k
    cls = np.zeros(n_start, np.int32)
    cls[len(v_names):] = 1
    order_run = np.array([index_of[k] for k in sorted(v_names + g2_names)], np.int32).reshape(-1)
    order_start = np.array([index_of[k] for k in sorted(v_names + g2_names + l_names)], np.int32).reshape(-1)

    ne = len(elems)
    kind = np.zeros(ne, np.int32)
    node = -np.ones((ne, 3), np.int32)
    branch = -np.ones(ne, np.int32)
    fn_kind = np.zeros(ne, np.int32)
    vals = np.zeros(ne, np.float64)
    fns = np.zeros((ne, 7), np.float64)
    for i, e in enumerate(elems):
        kind[i] = e.kind
        for j, n in enumerate(e.nodes):
            node[i, j] = -1 if n == GND else index_of[n]
        if e.kind in (V, L):
            branch[i] = index_of[e.name]
        fn_kind[i] = e.fn_kind
        vals[i] = e.val
        fns[i, :len(e.fn)] = e.fn
    return Circuit(elems, list(cmds), n_run, n_start, names, cls, order_run, order_start, kind, node, branch,
                   fn_kind, vals, fns, index_of)


def load(text: str) -> Circuit:
    elems, cmds = parse_netlist(text)
    check_elems(elems)
    return lower(elems, cmds)


# ------------------------------------------------------------------------------------------------
# synthetic configs of SURVEY.md §8(d)
# ------------------------------------------------------------------------------------------------
PROJ1_NL = """*Proj 1 - nonlinear variant (SURVEY.md 8d, config C2b)
V50 5 0 2V
V76 7 6 2V
V32 3 2 0.2V
I1 6 0 1mA
I2 8 4 1mA
R1 5 1 R=1.5
R2 5 2 R=50
R3 5 6 R=0.1
R4 1 2 R=1
R5 2 6 R=1.5
R6 3 4 R=0.1
R7 4 0 R=10
R8 8 0 R=1000
R9 8 9 R=1000
D90 0 9 d_model
R10 1 10 R=10k
R11 5 11 R=640
Q1 11 10 0 0 q_model
.OP
.END
"""


def ladder_netlist(n_nodes: int = 64, drive: str = "SIN( 0.0 1.0 10M )", stop: str = "5.4u", step: str = "1n",
                   ind_every: int = 4) -> str:
    """Config C4: RC/RL ladder, n = n_nodes + 1 unknowns (zero-padded names: lexicographic = physical)."""
    w = len(str(n_nodes - 1))
    nn = lambda i: f"n{i:0{w}d}"
    out = [f"* synthetic {n_nodes}-node RC/RL ladder", f"V1 {nn(0)} 0 {drive}"]
    for i in range(1, n_nodes):
        if ind_every and i % ind_every == 0:
            out.append(f"L{i} {nn(i - 1)} {nn(i)} L=10n")
        else:
            out.append(f"R{i} {nn(i - 1)} {nn(i)} R=100")
        out.append(f"C{i} {nn(i)} 0 C=1p")
    out += [f".TRAN {stop} {step}", ".END", ""]
    return "\n".join(out)


def inverter_chain_netlist(n_nodes: int = 256, stop: str = "168n", step: str = "50p") -> str:
    """Config C5: NMOS inverter chain with resistive pull-ups; n = n_nodes + 2 unknowns."""
    w = len(str(n_nodes - 2))
    nn = lambda i: f"n{i:0{w}d}"
    out = [f"* synthetic {n_nodes}-node NMOS inverter chain", "Vdd vdd 0 3V",
           f"Vin {nn(0)} 0 PULSE( 0.0 3 1n 2n 2n 10n 20n )"]
    for i in range(1, n_nodes - 1):
        out.append(f"R{i} vdd {nn(i)} R=10k")
        out.append(f"M{i} {nn(i)} {nn(i - 1)} 0 0 t_model")
        out.append(f"C{i} {nn(i)} 0 C=1p")
    out += [f".TRAN {stop} {step}", ".END", ""]
    return "\n".join(out)


# ------------------------------------------------------------------------------------------------
# Monte-Carlo perturbation: counter-based splitmix64, factor 1 + 0.1*u, u ~ U(-1,1)
# ------------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def monte_carlo(circ: Circuit, batch: int, rel: float = 0.1, seed: int = 0x5EED5EED, first: int = 0
                ) -> Tuple[np.ndarray, np.ndarray]:
    """Per-instance element values and waveform parameters.

    Every R/C/L value, every DC source value and every waveform amplitude-like parameter is multiplied by
    ``1 + rel*u``; global instance 0 is the unperturbed netlist.  ``first`` = global index of the first
    instance (so that shards of one batch draw the same numbers whatever the GPU count).
    Returns (vals[batch, n_elems], fns[batch, n_elems, 7]).
    """
    ne = circ.n_elems
    inst = (np.arange(batch, dtype=np.uint64) + np.uint64(first))[:, None]
    par = np.arange(ne * 8, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        key = np.uint64(seed) ^ (inst * np.uint64(0x100000001B3)) ^ (par * np.uint64(0xD6E8FEB86659FD93))
    bits = _splitmix64(_splitmix64(key))
    u = (bits >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0   # U(-1,1)
    fac = (1.0 + rel * u).reshape(batch, ne, 8)
    fac[inst[:, 0] == 0] = 1.0
    vals = np.broadcast_to(circ.vals, (batch, ne)).copy()
    fns = np.broadcast_to(circ.fns, (batch, ne, 7)).copy()
    scal = np.isin(circ.kind, (R, C, L, V, I))
    vals[:, scal] *= fac[:, scal, 0]
    # amplitude-like waveform parameters: SIN amplitude (1); PULSE v2 (1); EXP v2 (1)
    has_fn = circ.fn_kind != FN_NONE
    fns[:, has_fn, 1] *= fac[:, has_fn, 1]
    # sources with a waveform: their DC value is f(0), re-derived from fns by the engine (src/engine.rs:47-51)
    return vals, fns
