#!/usr/bin/env python
"""bench.py — batched solve-points/s of the B200 circuit-solve engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic circuit instances.  Default workload =
BASELINE.json configs[1] (C2): the proj1-style nonlinear netlist (diode + NPN), .op, 65 536 instances per GPU,
+-10 % Monte-Carlo element values.  `value` = solve-points/s with the parameters already resident in HBM
(device pointers through the C-ABI); `e2e` = the same through the C-ABI with pinned HOST buffers, host->device
and device->host copies inside the timed region.  One process per GPU (torchrun for N > 1); instances are
sharded with no data-path collective (weak scaling); times are CUDA-event times, max over ranks.

`--impl reference` times the reference's algorithm on the host cores (the C oracle port — the Rust reference
cannot be built in this image) on a bounded sample of the same workload, all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from ftspice_b200 import netlist as nl  # noqa: E402


# ----------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------
def make_workload(name: str, scale: float):
    if name == "c2":
        circ = nl.load(nl.PROJ1_NL)
        return dict(kind="op", circ=circ, batch=max(1, int(65536 * scale)),
                    label="C2 proj1-nl (diode+NPN, n=14, m=4) .op x 65536 instances/GPU, +-10% Monte Carlo")
    if name == "c4":
        circ = nl.load(nl.ladder_netlist(64, "SIN( 0.0 1.0 10M )", stop="5.4u", step="1n"))
        return dict(kind="tran", circ=circ, batch=max(1, int(65536 * scale)), stop=5.4e-6, step=1e-9,
                    label="C4 64-node RC/RL ladder (n=65) .tran 5.4u 1n (~10k steps) x 65536 instances/GPU, SIN drive")
    if name == "c5":
        circ = nl.load(nl.inverter_chain_netlist(256, stop="168n", step="50p"))
        return dict(kind="tran", circ=circ, batch=max(1, int(16384 * scale)), stop=168e-9, step=50e-12,
                    label="C5 256-node NMOS inverter chain (n=258) .tran 168n 50p (~5k steps) x 16384 instances")
    raise SystemExit(f"unknown workload {name}")


NMOS_TEST = """* NMOS with pull up resistor (reference test/nmos_test.sp shape)
V01 1 0 0.8V
V02 2 0 5V
R23 2 3 R=1000
M310 3 1 0 0 t_model
.OP
.END
"""


def make_workload_c3(scale: float, world: int, points: int = 32):
    """1M sweep points cut into chains of `points` points; each chain is one reference `.dc` run (cold start at the chain
    head, warm starts inside, src/engine.rs:116-127).  Shorter chains = more parallel work for the same point count."""
    circ = nl.load(NMOS_TEST)
    chains_total = max(world, int((1 << 20) // points * scale))
    return dict(kind="dc", circ=circ, batch=chains_total // world, points=points, sweep="V01", lo=0.0, hi=3.4,
                chains_total=chains_total,
                label=f"C3 nmos_test .dc V01 0..3.4 as {chains_total} chains x {points} points (1M sweep points at scale 1), "
                      f"sharded over the GPUs")


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
    # executed residual: every kernel forms A*x from the flat stamp map (one multiply-add per linear / companion
    # stamp term), not from a dense n x n product
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
# clocks
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


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the ORACLE, executed as the checker-side baseline only)
# ----------------------------------------------------------------------------------------------------------
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
        st, done, it, x = o.batch_dc(vals, None, circ.elem_index(wl["sweep"]), start, stop, step, wl["points"],
                                     threads=threads)
        dt = time.perf_counter() - t0
        return int(done.sum()), dt
    vals, fns = nl.monte_carlo(circ, sample, first=first)
    t0 = time.perf_counter()
    st, nrec, *_ = o.batch_tran(vals, fns, wl["stop"], wl["step"], 0, threads=threads, want_wave=False)
    dt = time.perf_counter() - t0
    return int(nrec.sum()), dt


def dc_ranges(wl, count: int, first: int):
    width = (wl["hi"] - wl["lo"]) / wl["chains_total"]
    idx = np.arange(first, first + count, dtype=np.float64)
    start = wl["lo"] + idx * width
    step = np.full(count, width / wl["points"])
    stop = start + width - 0.25 * step
    return start, stop, step


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 50 for c2/c3, 3 for the .tran workloads)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0, help="batch-size multiplier (1.0 = the named config)")
    ap.add_argument("--tran-stop", type=float, default=None, help="override .tran stop time (c4/c5 shortening)")
    ap.add_argument("--dc-points", type=int, default=32, help="points per .dc chain of workload c3 (1M points in total)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances of the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (e.g. lanes=2, force_generic=1)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 50 if args.workload in ("c2", "c3") else 3
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    wl = make_workload_c3(args.scale, world, args.dc_points) if args.workload == "c3" else make_workload(args.workload, args.scale)
    if args.tran_stop and wl["kind"] == "tran":
        wl["stop"] = args.tran_stop
        wl["label"] += f" [stop overridden to {args.tran_stop:g}s]"
    circ = wl["circ"]
    unit = "solve-points/s"
    metric = f"batched solve-points/sec ({'.' + wl['kind']})"

    # ------------------------------------------------------------------ reference arm (host cores only)
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        # bounded sample of the same workload per step: the full batch for .op (a few hundred ms on a multi-core host)
        sample = args.cpu_sample or {"op": wl["batch"], "dc": min(wl["batch"], 1024), "tran": 16}[wl["kind"]]
        if wl["kind"] == "tran" and not args.tran_stop:
            wl["stop"] = wl["stop"] / 20.0  # bounded: 1/20 of the time span per step (about 500 accepted steps)
        for _ in range(max(args.warmup, 0)):
            cpu_run(wl, max(threads, sample // 8), threads)
        pts, secs = 0, 0.0
        for _ in range(max(args.steps, 1)):
            p, s = cpu_run(wl, sample, threads)
            pts += p
            secs += s
        v = pts / secs
        samp = f"{sample} instances/step of the same workload" + (f", .tran stop {wl['stop']:g}s" if wl["kind"] == "tran" else "")
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["label"], "note": "reference algorithm on host cores: C oracle port of src/engine "
                       "(the Rust reference cannot be built here: no cargo); pthreads over instances"},
            "cpu_baseline": {"value": v, "unit": unit, "cores": threads, "kind": "port", "sample": samp},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from ftspice_b200 import capi
    from ftspice_b200.engine import BatchEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    B = wl["batch"]
    first = rank * B
    be = BatchEngine(circ, device=local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        be.set_option(k, int(v))
    stream = torch.cuda.Stream(device=dev)
    capi.check(capi.lib().ftsb_set_stream(be.h, stream.cuda_stream), be.h)
    n_out = circ.n_start if wl["kind"] == "op" else circ.n_run

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
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
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
    import ctypes as ct
    vp = lambda t: ct.c_void_p(t.data_ptr())
    L = capi.lib()

    def step_host():
        # the call a user makes with host buffers: parameters in, results out, blocking (ftsb_solve_op / ftsb_solve_dc
        # cut the batch into chunks so that the copies of one chunk overlap the solve of the others)
        fnp = vp(h_fns) if be.n_fn else None
        if wl["kind"] == "op":
            capi.check(L.ftsb_solve_op(be.h, B, vp(h_vals), fnp, capi.LAYOUT_INSTANCE_MAJOR, vp(ho_x), vp(ho_it), vp(ho_st)), be.h)
        elif wl["kind"] == "dc":
            capi.check(L.ftsb_solve_dc(be.h, B, vp(h_vals), fnp, capi.LAYOUT_INSTANCE_MAJOR, circ.elem_index(wl["sweep"]),
                                       vp(h_rng[0]), vp(h_rng[1]), vp(h_rng[2]), P, vp(ho_x), vp(ho_it), vp(ho_done),
                                       vp(ho_st)), be.h)
        else:
            capi.check(L.ftsb_set_params(be.h, B, vp(h_vals), fnp, capi.LAYOUT_INSTANCE_MAJOR, capi.MEM_HOST), be.h)
            opts = capi.FtsbTranOpts(0, 1, 0)
            capi.check(L.ftsb_run_tran(be.h, wl["stop"], wl["step"], ct.byref(opts), None, None, None, vp(ho_nrec),
                                       vp(ho_x), vp(ho_st), capi.MEM_HOST), be.h)

    h2d = h_vals.numel() * 8 + (h_fns.numel() * 8 if be.n_fn else 0) + (3 * B * 8 if wl["kind"] == "dc" else 0)
    d2h = ho_x.numel() * 8 + ho_st.numel() * 4 + (ho_it.numel() * 4 if wl["kind"] != "tran" else 0) + \
        (B * 8 if wl["kind"] != "op" else 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps (untimed)."""
        evs = []
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fn()
            e1.record(stream)
            evs.append((e0, e1))
        stream.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    # warm-up
    for _ in range(args.warmup):
        step_device()
    be.sync()
    for _ in range(min(args.warmup, 2)):
        step_host()

    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = be.launch_count
    t_wall0 = time.perf_counter()
    ms_dev = timed(step_device, args.steps)
    launches = be.launch_count - launches0
    dev_variant = be.variant
    barrier()
    t_wall = time.perf_counter() - t_wall0
    cnt = be.counters()
    ms_e2e = timed(step_host, args.steps)
    barrier()
    e2e_variant = be.variant
    # the host-buffer path must have produced what the device path produced (same instances, same kernel arithmetic)
    if not (torch.equal(ho_x.view(torch.int64), d_x.cpu().view(torch.int64)) and torch.equal(ho_st, d_st.cpu())):
        raise RuntimeError("bench: host-buffer (e2e) results differ from the device-buffer results")
    clocks = sampler.stop() if rank == 0 else {}

    # solve points of one step on this rank (device results of the last device step)
    if wl["kind"] == "op":
        pts_rank = int((d_st == 0).sum().item())
    elif wl["kind"] == "dc":
        pts_rank = int(d_done.sum().item())
    else:
        pts_rank = int(d_nrec.sum().item())
    ok_rank = int((d_st == 0).sum().item())

    tot = torch.tensor([sum(ms_dev), sum(ms_e2e), float(pts_rank), float(ok_rank)], dtype=torch.float64, device=dev)
    mx = tot.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        pts_all, ok_all = sm[2].item(), sm[3].item()
        # the only NVLink traffic of the job: gather the operating points / final states on every rank (untimed)
        g0 = time.perf_counter()
        gathered = torch.empty((world * B, n_out), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, d_x.reshape(B, -1)[:, :n_out].contiguous())
        torch.cuda.synchronize()
        gather_ms = 1e3 * (time.perf_counter() - g0)
    else:
        pts_all, ok_all = float(pts_rank), float(ok_rank)
        gather_ms = 0.0
    dev_ms_total, e2e_ms_total = mx[0].item(), mx[1].item()

    if rank == 0:
        value = pts_all * args.steps / (dev_ms_total * 1e-3)
        e2e_v = pts_all * args.steps / (e2e_ms_total * 1e-3)
        fm = flop_model(circ, wl["kind"] if wl["kind"] == "op" else "run")
        ref_flops = cnt["newton_iters"] * (fm["factor"] + fm["solve"] + fm["rest"])
        sparse = dev_variant.startswith(("spw", "jit", "plan"))  # kernels that skip structural zeros: device / plan flop counts
        if sparse:  # structure-exploiting kernel: flops it actually executed (device counter) + the per-iteration rest
            exec_flops = cnt["sparse_flops"] + cnt["newton_iters"] * fm["rest_exec"]
        else:
            exec_flops = cnt["lu_factor"] * fm["factor"] + cnt["lu_solve"] * fm["solve"] + cnt["newton_iters"] * fm["rest_exec"]
        kern_ms = float(np.mean(ms_dev))  # events bracket the step's launches (one kernel for .op/.dc; start-up + run for .tran)
        v0 = dev_variant.split("/")[0]
        kname = ("ftsb_jit (NVRTC, per topology)" if v0.startswith("jit") else "k_plan" if v0.startswith("plan") else
                 "k_spw" if v0.startswith("spw") else "k_sync" if v0.startswith(("sync", "zskip")) else
                 "k_lanes" if v0.startswith("lanes") else "k_warp" if "r_n" in v0 else
                 "k_cta" if v0.startswith("cta") else "k_subwarp")
        ncu_extra = {}
        traffic = None  # dram bytes per launch of that kernel from the committed ncu capture (profiles/traffic.json)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj.get(args.workload, {})
            if ent.get("variant_prefix") and dev_variant.startswith(ent["variant_prefix"]) and args.scale == 1.0:
                traffic = ent["dram_bytes_per_launch"]
                ncu_extra = {k: ent[k] for k in ("fp64_pipe_pct", "issue_active_pct", "warp_instructions_per_launch") if k in ent}
        except Exception:
            pass
        fp64_peak = float(L.ftsb_measure_fp64_tflops(local_rank))
        achieved = exec_flops / (kern_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        n_par = int(np.isin(circ.kind, (nl.R, nl.C, nl.L, nl.V, nl.I)).sum()) + 7 * be.n_fn
        pts_per_inst = max(pts_rank / max(B, 1), 1e-9)
        alg_bytes = 8.0 * (n_par / pts_per_inst + fm["n"] + 2) * pts_rank
        out = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_total / args.steps, "higher_is_better": True, "scaling": "weak" if wl["kind"] != "dc" else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["label"], "instances_per_gpu": B, "n_unknowns": fm["n"], "variant": dev_variant,
                       "l2": "flushed (256 MiB memset) between timed steps", "converged_instances": int(ok_all),
                       "solve_points_per_step": int(pts_all), "arithmetic": "fp64, no FMA contraction on the solution path",
                       "wall_s_timed_region": t_wall, "final_gather_ms_untimed": gather_ms},
            "clocks": clocks,
            "e2e": {"value": e2e_v, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms_total / args.steps, "variant": e2e_variant},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak > 0 else None, "traffic": traffic,
                         "kernel": f"{kname}<{wl['kind']}> ({dev_variant})",
                         "kernel_ms": kern_ms, "launches_per_step": launches / max(args.steps, 1),
                         "note": "per step = per batch: the .op step of a nonlinear small circuit is two launches of the same "
                                 "generated kernel (pass A: iterations 0..7, pass B: the rest) + the dense redo pass; flops and "
                                 "kernel_ms cover the whole step, profiles/traffic.json holds the ncu figures of pass B",
                         "executed_flops_per_launch": exec_flops,
                         "reference_equivalent_flops_per_launch": ref_flops,
                         "peak_source": "DFMA micro-benchmark in this run (ftsb_measure_fp64_tflops); MEASURED_PEAKS.json has no "
                                        "fp64 entry. The solution path runs without FMA contraction (parity), so DMUL+DADD pairs "
                                        "cap it at 0.5 of this FMA peak",
                         "newton_iters_per_launch": cnt["newton_iters"], "lu_factor_per_launch": cnt["lu_factor"],
                         "lu_solve_per_launch": cnt["lu_solve"], "redo_instances": cnt.get("redo", 0),
                         "ncu": ncu_extra or None,
                         "hbm": {"achieved": alg_bytes / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak,
                                 "algorithmic_bytes_per_launch": alg_bytes,
                                 "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"}},
        }
        if not args.no_cpu and world == 1:
            # bounded sample: a few seconds of single-thread CPU work
            sample = args.cpu_sample or {"op": min(wl["batch"], 131072), "dc": min(wl["batch"], (1 << 20) // wl.get("points", 1)),
                                         "tran": min(wl["batch"], 8)}[wl["kind"]]
            wl_cpu = dict(wl)
            samp = f"{sample} instances of the same workload (same seed, instances 0..{sample - 1})"
            if wl["kind"] == "tran":
                wl_cpu["stop"] = wl["stop"] / 20.0
                samp += f", .tran stop shortened to {wl_cpu['stop']:g}s"
            p, s = cpu_run(wl_cpu, sample, 1)
            out["cpu_baseline"] = {"value": p / s, "unit": unit, "cores": 1, "kind": "port", "sample": samp,
                                   "note": "C oracle restatement of the reference's src/engine, single thread like the "
                                           "reference; not the Rust binary (no cargo in this image)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
