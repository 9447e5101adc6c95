"""The C++ host side (ftspice_b200/host: parser, NodeCollection, Engine, SimResult, CLI mirroring src/main.rs).

CPU part: `ftspice_b200 --lower` (parse + index lowering, no GPU) must produce exactly the descriptor the Python
harness produces, for the 17 reference netlists and the synthetic configs.  GPU part: the CLI's CSV output equals
the Python Engine mirror's, byte for byte, and the exit codes follow the reference's `main`."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

from ftspice_b200 import netlist as nl
from ftspice_b200.engine import rust_f64_str

from harness import ROOT, golden, netlist_names

CLI = os.path.join(ROOT, "ftspice_b200", "host", "ftspice_b200")


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(CLI):
        subprocess.run(["make", "-C", os.path.dirname(CLI)], check=True, capture_output=True)
    return CLI


def run_cli(cli, text, *flags):
    with tempfile.NamedTemporaryFile("w", suffix=".sp", delete=False) as f:
        f.write(text)
        path = f.name
    try:
        return subprocess.run([cli, *flags, path], capture_output=True, text=True, timeout=120)
    finally:
        os.unlink(path)


def texts():
    out = {n: golden()[n]["netlist"] for n in netlist_names()}
    out["proj1_nl"] = nl.PROJ1_NL
    out["ladder64"] = nl.ladder_netlist(64)
    out["ladder_pulse"] = nl.ladder_netlist(12, "PULSE( 0.0 1.0 0.0 10n 10n 50n 100n )", stop="60n", step="1n")
    out["inverter256"] = nl.inverter_chain_netlist(256)
    return out


@pytest.mark.parametrize("name", sorted(texts()))
def test_cpp_lowering_equals_python_lowering(cli, name):
    text = texts()[name]
    r = run_cli(cli, text, "--lower")
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout)
    c = nl.load(text)
    assert d["n_run"] == c.n_run and d["n_start"] == c.n_start
    assert d["names"] == c.names
    assert d["cls"] == [int(v) for v in c.cls]
    assert d["headers_start"] == c.headers(True) and d["headers_run"] == c.headers(False)
    assert len(d["elems"]) == c.n_elems
    for i, e in enumerate(d["elems"]):
        assert e["name"] == c.elems[i].name and e["kind"] == int(c.kind[i])
        assert e["node"] == [int(v) for v in c.node[i]]
        assert e["branch"] == int(c.branch[i]) and e["fn_kind"] == int(c.fn_kind[i])
        assert float(e["val"]) == float(c.vals[i]), (e["name"], e["val"], c.vals[i])
        assert [float(v) for v in e["fn"]] == [float(v) for v in c.fns[i]]
        assert e["val"] == rust_f64_str(float(c.vals[i]))  # same Rust-style formatting in both hosts
    kinds = [k["kind"] for k in d["cmds"]]
    assert kinds == [k for k in ("op", "dc", "tran") if c.command(k)]
    for k in d["cmds"]:
        cmd = c.command(k["kind"])
        assert k["source"] == cmd.source and float(k["start"]) == cmd.start
        assert float(k["stop"]) == cmd.stop and float(k["step"]) == cmd.step


def test_cpp_cli_errors(cli):
    r = run_cli(cli, "R1 1 0 R=1\nR1 1 0 R=2\n.end\n", "--lower")
    assert r.returncode == 101 and "Duplicate elems found!" in r.stderr
    r = run_cli(cli, "R1 1 2 R=1\n.end\n", "--lower")
    assert r.returncode == 101 and "GND node not found!" in r.stderr
    r = run_cli(cli, "R1 1 0 R=1e3\n.end\n", "--lower")  # no exponent notation in the grammar
    assert r.returncode == 101
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 101 and "Insufficient arguments" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", netlist_names())
def test_cpp_cli_output_equals_python_engine(cli, name):
    """Reference behaviour of main.rs: op, dc, tran in that order; `?` stops at the first failing analysis."""
    from ftspice_b200.engine import Engine, NotConvergedError
    text = golden()[name]["netlist"]
    want, failed = "", False
    e = Engine.from_netlist(text)
    try:
        if e.op_cmd:
            want += e.run_op().to_csv()
        if e.dc_cmd:
            want += e.run_dc().to_csv()
        if e.tran_cmd:
            want += e.run_tran().to_csv()
    except NotConvergedError:
        failed = True
    r = run_cli(cli, text)
    assert r.stdout == want
    assert r.returncode == (1 if failed else 0), r.stderr
    if failed:
        assert "NotConvergedError" in r.stderr
    g = golden()[name]
    if "op" in g and g["op"]["status"] == 0:  # first data row = the golden .op row
        row = r.stdout.splitlines()[1].split(",")
        assert int(row[0]) == g["op"]["n_iters"]
