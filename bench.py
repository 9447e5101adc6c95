#!/usr/bin/env python
"""bench.py — batched solve-points/s of the B200 circuit-solve engine (BASELINE.json metric: .dc / .tran).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|c2|c3|c4|c5] [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic circuit instances; a *solve-point* is one row the
reference would print (one .op result, one .dc sweep point, one accepted .tran step).

Headline (the JSON line's top level) = config C4 of SURVEY.md §8(d): the 64-node RC/RL ladder (n = 65), SIN drive,
.TRAN with 1 ns maximum step, 65 536 Monte-Carlo instances per GPU.  The named span (5.4 us, ~10 k accepted steps per
instance, 6.6e8 solve-points per GPU) takes ~12 s per pass, so one timed step runs every instance from t = 0 over 1/5
of it (1.08 us, ~2070 accepted steps per instance, start-up .op and the h ramp-up included); `--full-span` times the whole span instead
(the full-span numbers of the round are committed under profiles/).  `configs` holds the other named configs measured in
the same run, each with its own roofline / cpu_baseline / e2e: C2 (.op, 65 536 instances per GPU), C3 (.dc, 1 M sweep
points in total) and C5 (.tran, 256-node inverter chain, 16 384 instances in total, 12 ns of the 168 ns span per step).

`value` = solve-points/s with the parameters already resident in HBM (device pointers through the C-ABI); `e2e` = the
same through the C-ABI with pinned HOST buffers, host->device and device->host copies inside the timed region.  One
process per GPU (torchrun for N > 1), no data-path collective: C4 and C2 keep the per-GPU batch (weak scaling), C3 and
C5 split their fixed total (strong scaling).  Times are CUDA-event times, max over ranks.

`--impl reference` times the reference's algorithm on the host cores (the C oracle port — the Rust reference cannot be
built in this image) on a bounded sample of the headline workload, all host threads.
"""
from __future__ import annotations

import argparse
import ctypes as ct
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from ftspice_b200 import netlist as nl  # noqa: E402

UNIT = "solve-points/s"

NMOS_TEST = """* NMOS with pull up resistor (reference test/nmos_test.sp shape)
V01 1 0 0.8V
V02 2 0 5V
R23 2 3 R=1000
M310 3 1 0 0 t_model
.OP
.END
"""


# ----------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------
def make_workload(name: str, scale: float, world: int, full_span: bool = False, tran_stop: float | None = None,
                  dc_points: int = 20):
    if name == "c2":
        circ = nl.load(nl.PROJ1_NL)
        return dict(name="c2", kind="op", circ=circ, batch=max(1, int(65536 * scale)), scaling="weak",
                    label="C2 proj1-nl (diode+NPN, n=14, m=4) .op x 65536 instances/GPU, +-10% Monte Carlo")
    if name == "c3":
        circ = nl.load(NMOS_TEST)
        chains_total = max(world, int((1 << 20) // dc_points * scale))
        return dict(name="c3", kind="dc", circ=circ, batch=chains_total // world, points=dc_points, sweep="V01", lo=0.0, hi=3.4,
                    chains_total=chains_total, scaling="strong",
                    label=f"C3 nmos_test .dc V01 0..3.4 as {chains_total} chains x {dc_points} points (1M sweep points at scale 1), "
                          f"sharded over the GPUs")
    if name == "c4":
        circ = nl.load(nl.ladder_netlist(64, "SIN( 0.0 1.0 10M )", stop="5.4u", step="1n"))
        stop = tran_stop or (5.4e-6 if full_span else 5.4e-6 / 5.0)
        span = "the full named span" if stop == 5.4e-6 else f"{stop:g}s = 1/{5.4e-6 / stop:g} of the named 5.4us span per step"
        return dict(name="c4", kind="tran", circ=circ, batch=max(1, int(65536 * scale)), stop=stop, step=1e-9, scaling="weak",
                    full_stop=5.4e-6,
                    label=f"C4 64-node RC/RL ladder (n=65), SIN drive, .TRAN 1n max step, 65536 instances/GPU, +-10% Monte Carlo; "
                          f"every step runs all instances from t=0 over {span}")
    if name == "c5":
        circ = nl.load(nl.inverter_chain_netlist(256, stop="168n", step="50p"))
        stop = tran_stop or (168e-9 if full_span else 12e-9)
        total = max(world, int(16384 * scale))
        span = "the full named span" if stop == 168e-9 else f"{stop:g}s of the named 168ns span per step (through the input's first pulse; the start-up .op of every instance is inside the step)"
        return dict(name="c5", kind="tran", circ=circ, batch=total // world, stop=stop, step=50e-12, scaling="strong",
                    full_stop=168e-9,
                    label=f"C5 256-node NMOS inverter chain (n=258) .TRAN 50p max step, {total} instances in total sharded over the "
                          f"GPUs, +-10% Monte Carlo; every step runs all instances from t=0 over {span}")
    raise SystemExit(f"unknown workload {name}")


def flop_model(circ: nl.Circuit, mode: str):
    """Reference-equivalent flops per Newton iteration (SURVEY.md §8d / BASELINE.md §4)."""
    n = circ.n_start if mode == "op" else circ.n_run
    total_lu = (n - 1) * n * (2 * n - 1) / 3 + 1.5 * n * (n - 1) + n * n + n
    factor = (n - 1) * n * (2 * n - 1) / 3 + 0.5 * n * (n - 1)
    solve = total_lu - factor
    nnz_h = 0
    dev = 0
    for e in circ.elems:
        gnd = sum(1 for x in e.nodes if x == nl.GND)
        if e.kind == nl.D:
            nnz_h += 2 - gnd
            dev += 12
        elif e.kind == nl.Q:
            nnz_h += 3 - gnd
            dev += 44
        elif e.kind == nl.M:
            nnz_h += 3 - gnd
            dev += 20
    rest = 2 * n * n + 2 * nnz_h + 4 * n + dev
    # executed residual: one multiply-add per linear / companion stamp term (or per assembled entry), not a dense n x n product
    ax_terms = 0
    for e in circ.elems:
        live = sum(1 for x in e.nodes[:2] if x != nl.GND)
        if e.kind == nl.R or (e.kind in (nl.C, nl.L) and mode != "op"):
            ax_terms += live * live
        elif e.kind == nl.V or (e.kind == nl.L and mode == "op"):
            ax_terms += 2 * live
    rest_exec = 2 * ax_terms + 2 * nnz_h + 4 * n + dev
    return dict(n=n, factor=factor, solve=solve, rest=rest, rest_exec=rest_exec)


# ----------------------------------------------------------------------------------------------------------
# clocks, NUMA
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.out = None

    def start(self):
        try:
            self.out = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.out,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.flush()
        self.out.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.out.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        try:
            os.unlink(self.out.name)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def pin_to_gpu_numa_node(torch, local_rank: int):
    """Run this rank (and first-touch its pinned buffers) on the cores of the GPU's NUMA node: the host<->device copy leg of
    the end-to-end calls is what stops scaling when every rank sits on node 0 (round-1 SCALE record).  Best effort."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for tok in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = tok.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the ORACLE, executed as the checker-side baseline only)
# ----------------------------------------------------------------------------------------------------------
def dc_ranges(wl, count: int, first: int):
    width = (wl["hi"] - wl["lo"]) / wl["chains_total"]
    idx = np.arange(first, first + count, dtype=np.float64)
    start = wl["lo"] + idx * width
    step = np.full(count, width / wl["points"])
    stop = start + width - 0.25 * step
    return start, stop, step


def cpu_run(wl, sample: int, threads: int, first: int = 0):
    """Runs `sample` instances of the workload on the host with the oracle; returns (points, seconds)."""
    from oracle import orc
    orc.build()
    circ = wl["circ"]
    o = orc.Oracle(circ)
    if wl["kind"] == "op":
        vals, fns = nl.monte_carlo(circ, sample, first=first)
        t0 = time.perf_counter()
        st, it, x = o.batch_op(vals, fns, threads=threads)
        dt = time.perf_counter() - t0
        return int((st == 0).sum()), dt
    if wl["kind"] == "dc":
        start, stop, step = dc_ranges(wl, sample, first)
        vals = np.broadcast_to(circ.vals, (sample, circ.n_elems)).copy()
        t0 = time.perf_counter()
        st, done, it, x = o.batch_dc(vals, None, circ.elem_index(wl["sweep"]), start, stop, step, wl["points"], threads=threads)
        dt = time.perf_counter() - t0
        return int(done.sum()), dt
    vals, fns = nl.monte_carlo(circ, sample, first=first)
    t0 = time.perf_counter()
    st, nrec, *_ = o.batch_tran(vals, fns, wl["stop"], wl["step"], 0, threads=threads, want_wave=False)
    dt = time.perf_counter() - t0
    return int(nrec.sum()), dt


def cpu_baseline(wl, sample: int):
    """Bounded single-thread sample of the same workload (the reference is single-threaded)."""
    p, s = cpu_run(wl, sample, 1)
    samp = f"{sample} instances of the same workload (same seed, instances 0..{sample - 1})"
    if wl["kind"] == "tran":
        samp += f", the step's span ({wl['stop']:g}s)"
    return {"value": p / s, "unit": UNIT, "cores": 1, "kind": "port", "sample": samp, "seconds": s,
            "note": "C oracle restatement of the reference's src/engine, single thread like the reference; not the Rust "
                    "binary (no cargo in this image)"}


def reference_arm(args, wl):
    threads = os.cpu_count() or 1
    sample = args.cpu_sample or {"op": wl["batch"], "dc": min(wl["batch"], 1024), "tran": 2 * threads}[wl["kind"]]
    for _ in range(max(args.warmup, 0)):
        cpu_run(wl, max(threads, sample // 8), threads)
    pts, secs = 0, 0.0
    for _ in range(max(args.steps, 1)):
        p, s = cpu_run(wl, sample, threads)
        pts += p
        secs += s
    v = pts / secs
    samp = f"{sample} instances/step of the same workload" + (f", the step's span ({wl['stop']:g}s)" if wl["kind"] == "tran" else "")
    print(json.dumps({
        "impl": "reference", "metric": f"batched solve-points/sec (.{wl['kind']})", "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
        "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "note": "reference algorithm on host cores: C oracle port of src/engine "
                   "(the Rust reference cannot be built here: no cargo); pthreads over instances"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": samp},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------------------
# one config on this rank's GPU
# ----------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def run_config(cx: Ctx, wl, steps: int, warmup: int, e2e_steps: int, opts, with_cpu: bool, sampler=None, cpu_sample: int = 0):
    torch, dist, capi, BatchEngine = cx.torch, cx.dist, cx.capi, cx.BatchEngine
    rank, world, local_rank, dev, stream = cx.rank, cx.world, cx.local_rank, cx.dev, cx.stream
    circ = wl["circ"]
    B = wl["batch"]
    first = rank * B
    be = BatchEngine(circ, device=local_rank)
    for kv in opts:
        k, v = kv.split("=")
        be.set_option(k, int(v))
    capi.check(capi.lib().ftsb_set_stream(be.h, stream.cuda_stream), be.h)
    n_out = circ.n_start if wl["kind"] == "op" else circ.n_run
    L = capi.lib()
    vp = lambda t: ct.c_void_p(t.data_ptr())

    # inputs: pinned host copies (for e2e) and device-resident copies (for `value`)
    if wl["kind"] == "dc":
        vals = np.broadcast_to(circ.vals, (B, circ.n_elems)).copy()
        fns = np.zeros((B, max(be.n_fn, 1), 7))
        start, stop, stp = dc_ranges(wl, B, first)
    else:
        vals, fns_full = nl.monte_carlo(circ, B, first=first)
        fns = be.compact_fns(fns_full) if be.n_fn else np.zeros((B, 1, 7))
    h_vals = torch.from_numpy(vals).pin_memory()
    h_fns = torch.from_numpy(np.ascontiguousarray(fns)).pin_memory()
    with torch.cuda.stream(stream):
        d_vals = h_vals.to(dev, non_blocking=True)
        d_fns = h_fns.to(dev, non_blocking=True)
        if wl["kind"] == "op":
            d_x = torch.empty((B, n_out), dtype=torch.float64, device=dev)
            d_it = torch.empty(B, dtype=torch.int32, device=dev)
        elif wl["kind"] == "dc":
            P = wl["points"]
            d_start, d_stop, d_step = (torch.from_numpy(a).to(dev) for a in (start, stop, stp))
            d_x = torch.zeros((B, P, n_out), dtype=torch.float64, device=dev)
            d_it = torch.zeros((B, P), dtype=torch.int32, device=dev)
            d_done = torch.empty(B, dtype=torch.int64, device=dev)
        else:
            d_x = torch.empty((B, n_out), dtype=torch.float64, device=dev)  # x_final only: full waveforms of the
            d_nrec = torch.empty(B, dtype=torch.int64, device=dev)          # named config are 333 GB (SURVEY hard part 6)
        d_st = torch.empty(B, dtype=torch.int32, device=dev)
    stream.synchronize()
    be.set_params_device(B, d_vals.data_ptr(), d_fns.data_ptr() if be.n_fn else 0)
    be.sync()

    def step_device():
        if wl["kind"] == "op":
            be.run_op_device(d_x.data_ptr(), d_it.data_ptr(), d_st.data_ptr())
        elif wl["kind"] == "dc":
            be.run_dc_device(circ.elem_index(wl["sweep"]), d_start.data_ptr(), d_stop.data_ptr(), d_step.data_ptr(), P,
                             d_x.data_ptr(), d_it.data_ptr(), d_done.data_ptr(), d_st.data_ptr())
        else:
            be.run_tran_device(wl["stop"], wl["step"], 0, 1, 0, 0, 0, d_nrec.data_ptr(), d_x.data_ptr(), d_st.data_ptr())

    # host-buffer (e2e) path: pinned outputs
    if wl["kind"] == "op":
        ho_x = torch.empty((B, n_out), dtype=torch.float64).pin_memory()
        ho_it = torch.empty(B, dtype=torch.int32).pin_memory()
    elif wl["kind"] == "dc":
        ho_x = torch.empty((B, P, n_out), dtype=torch.float64).pin_memory()
        ho_it = torch.empty((B, P), dtype=torch.int32).pin_memory()
        ho_done = torch.empty(B, dtype=torch.int64).pin_memory()
        h_rng = [torch.from_numpy(a).pin_memory() for a in (start, stop, stp)]
    else:
        ho_x = torch.empty((B, n_out), dtype=torch.float64).pin_memory()
        ho_nrec = torch.empty(B, dtype=torch.int64).pin_memory()
    ho_st = torch.empty(B, dtype=torch.int32).pin_memory()

    def step_host():
        # the call a user makes with host buffers: parameters in, results out, blocking
        fnp = vp(h_fns) if be.n_fn else None
        if wl["kind"] == "op":
            capi.check(L.ftsb_solve_op(be.h, B, vp(h_vals), fnp, capi.LAYOUT_INSTANCE_MAJOR, vp(ho_x), vp(ho_it), vp(ho_st)), be.h)
        elif wl["kind"] == "dc":
            capi.check(L.ftsb_solve_dc(be.h, B, vp(h_vals), fnp, capi.LAYOUT_INSTANCE_MAJOR, circ.elem_index(wl["sweep"]),
                                       vp(h_rng[0]), vp(h_rng[1]), vp(h_rng[2]), P, vp(ho_x), vp(ho_it), vp(ho_done),
                                       vp(ho_st)), be.h)
        else:
            capi.check(L.ftsb_set_params(be.h, B, vp(h_vals), fnp, capi.LAYOUT_INSTANCE_MAJOR, capi.MEM_HOST), be.h)
            o = capi.FtsbTranOpts(0, 1, 0)
            capi.check(L.ftsb_run_tran(be.h, wl["stop"], wl["step"], ct.byref(o), None, None, None, vp(ho_nrec),
                                       vp(ho_x), vp(ho_st), capi.MEM_HOST), be.h)

    h2d = h_vals.numel() * 8 + (h_fns.numel() * 8 if be.n_fn else 0) + (3 * B * 8 if wl["kind"] == "dc" else 0)
    d2h = ho_x.numel() * 8 + ho_st.numel() * 4 + (ho_it.numel() * 4 if wl["kind"] != "tran" else 0) + \
        (B * 8 if wl["kind"] != "op" else 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, k):
        """k steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps (untimed)."""
        evs = []
        for _ in range(k):
            with torch.cuda.stream(stream):
                cx.flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fn()
            e1.record(stream)
            evs.append((e0, e1))
        stream.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    for _ in range(warmup):
        step_device()
    be.sync()
    for _ in range(min(warmup, 2)):
        step_host()

    barrier()
    if sampler is not None and rank == 0:
        sampler.start()
    launches0 = be.launch_count
    t_wall0 = time.perf_counter()
    ms_dev = timed(step_device, steps)
    launches = be.launch_count - launches0
    dev_variant = be.variant
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if (sampler is not None and rank == 0) else None
    cnt = be.counters()
    ms_e2e = timed(step_host, e2e_steps)
    barrier()
    e2e_variant = be.variant
    # the host-buffer path must have produced what the device path produced (same instances, same kernel arithmetic)
    if not (torch.equal(ho_x.view(torch.int64), d_x.cpu().view(torch.int64)) and torch.equal(ho_st, d_st.cpu())):
        raise RuntimeError(f"bench {wl['name']}: host-buffer (e2e) results differ from the device-buffer results")

    if wl["kind"] == "op":
        pts_rank = int((d_st == 0).sum().item())
    elif wl["kind"] == "dc":
        pts_rank = int(d_done.sum().item())
    else:
        pts_rank = int(d_nrec.sum().item())
    ok_rank = int((d_st == 0).sum().item())

    tot = torch.tensor([sum(ms_dev) / steps, sum(ms_e2e) / e2e_steps, float(pts_rank), float(ok_rank)], dtype=torch.float64, device=dev)
    mx = tot.clone()
    gather_ms = 0.0
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        pts_all, ok_all = sm[2].item(), sm[3].item()
        # the only NVLink traffic of the job: gather the operating points / final states on every rank (untimed)
        g0 = time.perf_counter()
        rows = d_x.reshape(B, -1)
        gathered = torch.empty((world * B, rows.shape[1]), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, rows.contiguous())
        torch.cuda.synchronize()
        gather_ms = 1e3 * (time.perf_counter() - g0)
        del gathered
    else:
        pts_all, ok_all = float(pts_rank), float(ok_rank)
    dev_ms, e2e_ms = mx[0].item(), mx[1].item()

    out = None
    if rank == 0:
        value = pts_all / (dev_ms * 1e-3)
        e2e_v = pts_all / (e2e_ms * 1e-3)
        fm = flop_model(circ, wl["kind"] if wl["kind"] == "op" else "run")
        ref_flops = cnt["newton_iters"] * (fm["factor"] + fm["solve"] + fm["rest"])
        sparse = dev_variant.startswith(("spw", "jit", "plan", "team"))  # kernels that skip structural zeros
        if sparse:  # flops actually executed (device counter / plan counts) + the per-iteration rest
            exec_flops = cnt["sparse_flops"] + cnt["newton_iters"] * fm["rest_exec"]
        else:
            exec_flops = cnt["lu_factor"] * fm["factor"] + cnt["lu_solve"] * fm["solve"] + cnt["newton_iters"] * fm["rest_exec"]
        kern_ms = float(np.mean(ms_dev))
        v0 = dev_variant.split("/")[0]
        kname = ("k_team_tran" if v0.startswith("team") else "ftsb_jit (NVRTC, per topology)" if v0.startswith("jit") else
                 "k_plan" if v0.startswith("plan") else "k_spw" if v0.startswith("spw") else
                 "k_sync" if v0.startswith(("sync", "zskip")) else "k_lanes" if v0.startswith("lanes") else
                 "k_warp" if "r_n" in v0 else "k_cta" if v0.startswith("cta") else "k_subwarp")
        ncu_extra, traffic, traffic_note = {}, None, None
        try:  # dram bytes of the dominant kernel from the committed ncu capture, scaled to this run's work
            ent = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wl["name"], {})
            if ent.get("variant_prefix") and dev_variant.startswith(ent["variant_prefix"]):
                if "dram_bytes_per_solve_point" in ent:
                    traffic = ent["dram_bytes_per_solve_point"] * pts_rank
                    traffic_note = f"{ent['dram_bytes_per_solve_point']:.1f} B per solve-point in the ncu capture ({ent.get('source')}) x this step's solve-points"
                elif cx.scale == 1.0:
                    traffic = ent["dram_bytes_per_launch"]
                    traffic_note = ent.get("source")
                ncu_extra = {k: ent[k] for k in ("fp64_pipe_pct", "issue_active_pct", "warp_instructions_per_launch",
                                                 "warp_instructions_per_solve_point") if k in ent}
        except Exception:
            pass
        achieved = exec_flops / (kern_ms * 1e-3) / 1e12
        n_par = int(np.isin(circ.kind, (nl.R, nl.C, nl.L, nl.V, nl.I)).sum()) + 7 * be.n_fn
        pts_per_inst = max(pts_rank / max(B, 1), 1e-9)
        alg_bytes = 8.0 * (n_par / pts_per_inst + fm["n"] + 2) * pts_rank
        out = {
            "metric": f"batched solve-points/sec (.{wl['kind']})", "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["label"], "instances_per_gpu": B, "n_unknowns": fm["n"], "variant": dev_variant,
                       "l2": "flushed (256 MiB memset) between timed steps", "converged_instances": int(ok_all),
                       "solve_points_per_step": int(pts_all), "arithmetic": "fp64, no FMA contraction on the solution path",
                       "wall_s_timed_region": t_wall, "final_gather_ms_untimed": gather_ms},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "steps": e2e_steps, "variant": e2e_variant},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": cx.fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / cx.fp64_peak if cx.fp64_peak > 0 else None, "traffic": traffic,
                         "traffic_note": traffic_note,
                         "kernel": f"{kname}<{wl['kind']}> ({dev_variant})",
                         "kernel_ms": kern_ms, "launches_per_step": launches / max(steps, 1),
                         "note": "flops and kernel_ms cover the whole step (for .tran: start-up .op launches + the .tran kernel + "
                                 "its redo passes; for the .op of a nonlinear small circuit: two passes of the generated kernel + the "
                                 "dense redo pass); the named kernel takes > 90 % of it (profiles/)",
                         "executed_flops_per_launch": exec_flops,
                         "reference_equivalent_flops_per_launch": ref_flops,
                         "peak_source": cx.fp64_peak_source,
                         "newton_iters_per_launch": cnt["newton_iters"], "lu_factor_per_launch": cnt["lu_factor"],
                         "lu_solve_per_launch": cnt["lu_solve"], "redo_instances": cnt.get("redo", 0),
                         "ncu": ncu_extra or None,
                         "hbm": {"achieved": alg_bytes / (kern_ms * 1e-3) / 1e9, "peak": cx.hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / cx.hbm_peak,
                                 "algorithmic_bytes_per_launch": alg_bytes, "peak_source": cx.hbm_peak_source}},
        }
        if clocks is not None:
            out["clocks"] = clocks
        if with_cpu and world == 1:
            sample = cpu_sample or {"op": min(B, 131072), "dc": min(B, (1 << 20) // wl.get("points", 1)),
                                    "tran": 8 if wl["name"] == "c4" else 4}[wl["kind"]]
            out["cpu_baseline"] = cpu_baseline(wl, min(sample, B))
    be.close()
    return out


def multi_api_leg(cx: Ctx, wl, steps: int):
    """N > 1: the same C2 batch (all ranks' instances) through ONE ftsb_multi handle over all N devices from one host process —
    the call the Rust side of src/main.rs makes on a multi-GPU box (no torch, no process group).  Host clock around the
    blocking calls; the other ranks wait at a barrier meanwhile."""
    torch, capi = cx.torch, cx.capi
    from ftspice_b200.engine import MultiEngine
    circ = wl["circ"]
    B = wl["batch"] * cx.world
    me = MultiEngine(circ, list(range(cx.world)))
    vals, fns_full = nl.monte_carlo(circ, B)
    fns = np.ascontiguousarray(fns_full[:, me.fn_elems, :]) if me.n_fn else np.zeros((B, 1, 7))
    h_vals = torch.from_numpy(vals).pin_memory()
    h_fns = torch.from_numpy(fns).pin_memory()
    n = circ.n_start
    ho_x = torch.empty((B, n), dtype=torch.float64).pin_memory()
    ho_it = torch.empty(B, dtype=torch.int32).pin_memory()
    ho_st = torch.empty(B, dtype=torch.int32).pin_memory()
    L = capi.lib()
    vp = lambda t: ct.c_void_p(t.data_ptr())

    def call():
        capi.check_multi(L.ftsb_multi_solve_op(me.h, B, vp(h_vals), vp(h_fns) if me.n_fn else None, vp(ho_x), vp(ho_it), vp(ho_st)), me.h)

    for _ in range(4):
        call()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    dt = time.perf_counter() - t0
    pts = int((ho_st == 0).sum().item())
    out = {"value": pts * steps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / steps, "steps": steps, "devices": cx.world,
           "instances": B, "variant": me.variant, "clock": "host perf_counter around the blocking ftsb_multi_solve_op calls",
           "h2d_bytes_per_step": int(h_vals.numel() * 8 + (h_fns.numel() * 8 if me.n_fn else 0)),
           "d2h_bytes_per_step": int(ho_x.numel() * 8 + 8 * B)}
    me.close()
    return out


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps of the headline config (default 5)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="all", choices=["all", "c2", "c3", "c4", "c5"],
                    help="all = headline C4 + sub-results C2, C3, C5 in `configs`; cN = that config alone as the headline")
    ap.add_argument("--scale", type=float, default=1.0, help="batch-size multiplier (1.0 = the named config)")
    ap.add_argument("--tran-stop", type=float, default=None, help="override the .tran span of a step")
    ap.add_argument("--full-span", action="store_true", help=".tran configs: one step = the whole named span")
    ap.add_argument("--dc-points", type=int, default=0, help="points per .dc chain of workload c3 (1M points in total); 0 = 20 on one "
                    "GPU (52 428 chains: 92 %% of the one wave of 148 SMs x 12 warps x 32 threads the generated kernel keeps resident; "
                    "32 768 chains x 32 points fill 58 %% of it and take 0.62 instead of 0.47 ms), 20 / N (at least 4) on N GPUs: a "
                    "chain is a sequential warm-started sweep, so the fixed total only scales with a full wave of chains per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances of the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs (profiling runs)")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value for the headline config")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 5
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.dc_points <= 0:
        args.dc_points = max(4, 20 // max(world, 1))
    head = "c4" if args.workload == "all" else args.workload
    wl = make_workload(head, args.scale, world, args.full_span, args.tran_stop, args.dc_points)

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist
    from ftspice_b200 import capi
    from ftspice_b200.engine import BatchEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = pin_to_gpu_numa_node(torch, local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")  # CPU-side waits (no NCCL kernel spinning on a GPU another process uses)
    cx = Ctx()
    cx.torch, cx.dist, cx.capi, cx.BatchEngine = torch, dist, capi, BatchEngine
    cx.rank, cx.world, cx.local_rank = rank, world, local_rank
    cx.dev = torch.device("cuda", local_rank)
    cx.stream = torch.cuda.Stream(device=cx.dev)
    cx.scale = args.scale
    with torch.cuda.stream(cx.stream):
        cx.flush = torch.empty(256 << 20, dtype=torch.uint8, device=cx.dev)  # > 126 MB L2
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    cx.hbm_peak = peaks.get("hbm_gbs", 6650.0)
    cx.hbm_peak_source = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"
    if "fp64_tflops" in peaks:
        cx.fp64_peak = float(peaks["fp64_tflops"])
        cx.fp64_peak_source = "MEASURED_PEAKS.json fp64_tflops"
    else:
        cx.fp64_peak = float(capi.lib().ftsb_measure_fp64_tflops(local_rank))
        cx.fp64_peak_source = ("DFMA micro-benchmark in this run (ftsb_measure_fp64_tflops: 8 independent FMA chains per thread, "
                               "8 blocks of 256 threads per SM; its SASS is profiles/r02_fp64_peak_sass.txt); MEASURED_PEAKS.json has "
                               "no fp64 entry; 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2 nominal. The solution path runs without "
                               "FMA contraction (parity), so DMUL+DADD pairs cap it at 0.5 of this FMA peak")

    tran_like = wl["kind"] == "tran"
    out = run_config(cx, wl, args.steps, args.warmup, max(2, min(args.steps, 5)) if tran_like else args.steps, args.opt,
                     not args.no_cpu, sampler=ClockSampler(local_rank), cpu_sample=args.cpu_sample)
    subs = {}
    if args.workload == "all":
        plan = [("c2", max(args.steps, 20), args.warmup, max(args.steps, 20)),
                ("c3", max(args.steps, 20), args.warmup, max(args.steps, 20)),
                ("c5", 2, 1, 1)]
        for name, k, w, ke in plan:
            swl = make_workload(name, args.scale, world, args.full_span, None, args.dc_points)
            try:
                r = run_config(cx, swl, k, w, ke, [], not args.no_cpu)
            except Exception as e:  # a failing sub-config must not take the headline with it
                r = {"error": f"{type(e).__name__}: {e}"} if rank == 0 else None
            if rank == 0:
                subs[name] = r
            if name == "c2" and world > 1:
                torch.cuda.synchronize()
                dist.barrier(group=host_group)
                if rank == 0:
                    try:
                        subs["c2"]["e2e_multi_api"] = multi_api_leg(cx, swl, 20)
                    except Exception as e:
                        subs["c2"]["e2e_multi_api"] = {"error": f"{type(e).__name__}: {e}"}
                dist.barrier(group=host_group)
    if rank == 0:
        out["config"]["numa_pinning"] = numa
        if subs:
            out["configs"] = subs
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
