// api.cu — the C-ABI of include/ftsb.h: handle management, parameter upload, kernel dispatch.
// No CPU fallback: every entry point that computes needs a CUDA device and fails loudly without one.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ftsb.h"
#include "compile.h"
#include "jit.h"
#include "solver_lanes.cuh"
#include "solver_plan.cuh"
#include "solver_sparse.cuh"
#include "solver_team.cuh"
#include "solver_warp.cuh"

using namespace ftsb;

namespace {

thread_local std::string g_last_error;

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per (function, device), not per handle: two handles on ONE device
// driven by two host threads (ftsb_multi with a device listed twice) would change it under each other's launches, so the
// plan-and-launch sequence of a run is serialised per device.  Launches are asynchronous: the lock is held for microseconds.
std::mutex &device_launch_mutex(int device)
{
    static std::mutex mu[64];
    return mu[device >= 0 && device < 64 ? device : 0];
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max(bytes, (size_t)256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

} // namespace

struct ftsb_circuit {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    size_t smem_optin = 0;
    Compiled comp;
    DevBuf ibuf;
    DevProg prog{};
    long long batch = 0;
    long long vs_inst = 0, vs_elem = 1, fs_inst = 0, fs_elem = 1; // layout of the device copies of the parameters (RunArgs)
    DevBuf vals, fn, tmp_in, work, counters, scratch;
    // staging of outputs when the caller's buffers are on the host
    DevBuf o_x, o_it, o_status, o_done, o_t, o_nrec, o_xf, i_start, i_stop, i_step;
    DevBuf s_x, s_it, s_st; // start-up solutions of .tran (small-circuit kernels)
    DevBuf redo;            // instances the sparse kernels hand to the dense kernels
    // learned elimination plans (plan.h) per analysis mode (.op, .dc): built by a pilot run on the first call
    PlanHost plan_host[2];
    DevBuf plan_buf[2], trace;
    int plan_state[2] = {0, 0}; // 0 = not tried, 1 = plans available, -1 = no usable plan
    std::map<unsigned long long, long long> plan_hist[2]; // traced pivot sequences and how often each occurred
    int relearn_left[2] = {0, 0};                          // how many more times redo traces may extend the plans
    bool redo_traced[2] = {false, false};                  // the last run's redo pass wrote a trace
    int last_redo_an = -1;                                 // analysis whose redo count sits in the counters
    int plan_opt = 1;           // 1: use the plan-interpreting kernels for n <= 16 (.op / .dc)
    int plan_lanes = 0;         // lanes per instance of those kernels (0 = automatic)
    int plan_max = PLAN_MAX;    // plans kept (most frequent pivot sequences first); set before the first run
    // per-topology specialised kernels (jit.h), one per analysis mode, generated from the plans
    JitKernel jit[2];
    int jit_state[2] = {0, 0}; // 0 = not tried, 1 = loaded, -1 = unavailable (NVRTC / driver missing, compile error)
    int jit_occ[2] = {0, 0};
    int jit_opt = 1;
    int jit_log1p = 1;      // 1: generated kernels use fts_log1p.h like the built kernels; 0: CUDA's log1p (A/B measurements)
    int jit_two_pass = 8;   // k > 0: .op of nonlinear circuits in passes of k Newton iterations (launch_jit)
    int jit_passes = 2;     // number of passes (the last one runs to the end)
    int jit_two_pass_min_plans = 3; // ... for circuits whose pivot sequences needed at least this many plans
    int jit_max_regs = 0;   // > 0: __maxnreg__ of the generated kernel (instead of jit_min_blocks).  Measured: 224 / 216 registers
                            // still give 2 warps per scheduler (16 K registers each): a third needs <= 168 and spills
    int jit_min_blocks = 1; // __launch_bounds__(32, this): register cap of the generated kernel
    int last_plan_an = -1;  // analysis mode of the last run when it used the plan / JIT kernels (flop accounting)
    std::string jit_msg;
    std::string err;
    std::string variant = "none";
    long long launches = 0;
    int skip_refactor = 1;
    int force_cta = 0;
    int force_generic = 0; // 1: never use the small-circuit lane kernels (A/B testing)
    int lanes = 0;         // >0: lanes per instance requested for the dense small-circuit kernels (if built)
    int sparse_warp = 1;   // 1 (default): structure-exploiting warp-per-instance kernels for n > 16
    int hist_global_opt = 0; // 1: mid-size .tran keeps the per-step history in global scratch (no gain at n = 65: registers bound the occupancy)
    int windowed = 1;      // 1: the sparse kernel eliminates windows of independent columns at once when the predicted
                           // structure allows it; 2: always try; 0: never
    int slots_global = 1;  // 1: large circuits keep the slot table and the solution vectors in L2-resident global scratch
    int fill_cap = 0;      // >0: run-time fill-in pool entries per instance of those kernels (0 = automatic)
    // pipelined host-buffer entry points (ftsb_solve_op / ftsb_solve_dc): one stream and one work counter per chunk
    static constexpr int PIPE_MAX = 8;
    cudaStream_t pipe[PIPE_MAX] = {};
    cudaEvent_t pipe_ev[PIPE_MAX] = {}, pipe_ev0 = nullptr;
    bool pipe_ready = false;
    int pipe_chunks = 4; // chunks per call (option "pipe_chunks"; 1 = serial copies)
    DevBuf pipe_work;
    static constexpr int PASS_MAX = 8;
    unsigned long long pipe_begin[PASS_MAX * PIPE_MAX] = {};
    // per-team .tran kernels for n > 16 (solver_team.cuh): static elimination plan along the predicted pivot sequence
    TeamPlanHost team_plan;
    DevBuf team_buf, team_scratch, redo2;
    int team_state = 0;    // 0 = not tried, 1 = plan on the device, -1 = no usable plan
    int team_opt = 1;      // option "team": 1 (default) use them, 0 = the warp-per-instance kernels as before
    int team_g = 0;        // option "team_lanes": lanes per instance (0 = automatic)
    int team_levels = 1;   // option "team_levels": 1 (default) level-scheduled elimination / substitutions for factors in shared memory, 0 = one lane walks them
    int team_rows = 1;     // option "team_rows": 1 (default) linear circuits use the kernels with compile-time rows per lane when one fits
    int team_place = -1;   // option "team_place": 0 slot table and factors in global scratch, 3 in shared memory (-1 = automatic)
    int l2_persist = 1;    // option "l2_persist": keep the per-warp global scratch rows resident in L2 (access policy window)
    size_t l2_set_aside = 0, l2_window_max = 0;
    bool l2_queried = false;
    int sparse = 0;        // 1: zero-skipping thread-per-instance kernels for n <= 16; 0 (default): dense elimination,
                           // G lanes per instance -- measured faster on B200 (the thread-per-instance mapping fits only
                           // 2-3 warps per SM in shared memory and is latency bound)
};

#define FTSB_CUDA(c, call)                                                                        \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            (c)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                       \
            return FTSB_E_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

namespace {

int fail(ftsb_circuit *c, int code, const std::string &msg)
{
    if (c)
        c->err = msg;
    else
        g_last_error = msg;
    return code;
}

Gather dev_gather(const int *base, const GatherOff &o)
{
    Gather g;
    g.dst = base + o.dst;
    g.ptr = base + o.ptr;
    g.con = base + o.con;
    g.n_dst = o.n_dst;
    g.n_con = o.n_con;
    return g;
}

ModeProg dev_mode(const int *base, const ModeOff &o)
{
    ModeProg m;
    m.n = o.n;
    m.nv = o.nv;
    m.ni = o.ni;
    m.n_nl = o.n_nl;
    m.n_gfun = o.n_gfun;
    m.pad = 0;
    m.cls = base + o.cls;
    m.nl = base + o.nl;
    m.A_lin = dev_gather(base, o.A_lin);
    m.A_dyn = dev_gather(base, o.A_dyn);
    m.J_nl = dev_gather(base, o.J_nl);
    m.b_all = dev_gather(base, o.b_all);
    m.r_nl = dev_gather(base, o.r_nl);
    m.f_nl = dev_gather(base, o.f_nl);
    m.J_full = dev_gather(base, o.J_full);
    m.Ax = dev_gather(base, o.Ax);
    m.sp.nnz = o.sp.nnz;
    m.sp.n_pred = o.sp.n_pred;
    m.sp.fcap = 0;
    m.sp.hist_global = 0;
    m.sp.windowed = 0;
    m.sp.pad3 = 0;
    m.sp.cptr = base + o.sp.cptr;
    m.sp.erow = base + o.sp.erow;
    m.sp.rptr = base + o.sp.rptr;
    m.sp.rcol = base + o.sp.rcol;
    m.sp.rent = base + o.sp.rent;
    m.sp.n_long = o.sp.n_long;
    m.sp.pad2 = 0;
    m.sp.rlong = base + o.sp.rlong;
    m.sp.lmap = base + o.sp.lmap;
    m.sp.S = dev_gather(base, o.sp.S);
    return m;
}

// ---- kernel variants -------------------------------------------------------------------------------------

template <int AN>
KernelFn subwarp_kernel(int g)
{
    switch (g) {
    case 4:
        return (KernelFn)k_subwarp<4, AN>;
    case 8:
        return (KernelFn)k_subwarp<8, AN>;
    case 16:
        return (KernelFn)k_subwarp<16, AN>;
    default:
        return (KernelFn)k_subwarp<32, AN>;
    }
}

KernelFn pick_kernel(int an, int g /*0 = cta*/)
{
    if (g == 0) return an == AN_OP ? (KernelFn)k_cta<AN_OP> : an == AN_DC ? (KernelFn)k_cta<AN_DC> : (KernelFn)k_cta<AN_TRAN>;
    return an == AN_OP ? subwarp_kernel<AN_OP>(g) : an == AN_DC ? subwarp_kernel<AN_DC>(g) : subwarp_kernel<AN_TRAN>(g);
}

struct LaunchPlan {
    KernelFn fn = nullptr;
    int g = 0; // sub-warp group size, 0 = CTA team
    int block = 128;
    int grid = 1;
    size_t smem = 0;
    Layout lay;
    size_t scratch_bytes = 0;
    bool lanes = false; // small-circuit kernel: .tran needs the start-up .op as a separate launch
};

// small-circuit kernels (solver_lanes.cuh): n <= 16.  sp = 1: zero-skipping elimination, one thread per instance;
// sp = 0: dense elimination, G lanes per instance.
KernelFn find_lanes_kernel(int g, int nc, int an, int sp)
{
    KernelFn (*const tab[])(int, int, int, int) = {lanes_kernel_a,  lanes_kernel_b,  lanes_kernel_c,  lanes_kernel_d, lanes_kernel_e,
                                                    lanes_kernel_s1, lanes_kernel_s2, lanes_kernel_s3, lanes_kernel_s4};
    for (auto f : tab)
        if (KernelFn fn = f(g, nc, an, sp)) return fn;
    return nullptr;
}

// Returns 0 when a small-circuit kernel was chosen.
int plan_lanes(ftsb_circuit *c, int an, long long batch, LaunchPlan &pl)
{
    const int n_start = std::max(c->comp.n_start, 1);
    const int n = an == AN_OP ? n_start : c->comp.n_run; // exact unknown count of the analysis mode
    if (n_start > 16 || n < 1) return 1;
    int g = 1, sp = 0;
    KernelFn fn = nullptr;
    if (c->sparse && c->lanes <= 0) { // opt-in: zero-skipping, thread per instance
        sp = 1;
        fn = find_lanes_kernel(1, n, an, 1);
    }
    if (!fn) {
        sp = 0;
        g = c->lanes > 0 ? c->lanes : (n <= 4 ? 1 : n <= 8 ? 2 : 4);
        fn = find_lanes_kernel(g, n, an, 0);
        if (!fn && c->lanes > 0) { // requested lane count not built: fall back to the default
            g = n <= 4 ? 1 : n <= 8 ? 2 : 4;
            fn = find_lanes_kernel(g, n, an, 0);
        }
    }
    if (!fn) return 1;
    const LaneLayout T = lane_layout(n, c->comp.n_slots, c->comp.n_dyn, c->comp.n_src, an == AN_TRAN, sp != 0);
    const size_t smem = (size_t)T.doubles * (32 / g) * 8;
    if (smem > c->smem_optin) return 1;
    if (cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, 32, smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        return 1;
    }
    pl.fn = fn;
    pl.g = g;
    pl.block = 32;
    pl.smem = smem;
    pl.lay = c->comp.lay;
    pl.lay.mat_in_smem = 1;
    pl.lay.bytes = (int)smem;
    pl.scratch_bytes = 0;
    pl.lanes = true;
    const int per_warp = 32 / g;
    long long need = (batch + per_warp - 1) / per_warp;
    pl.grid = (int)std::max(1LL, std::min(need, (long long)occ * c->sm_count));
    c->variant = std::string(sp ? "zskip" : an == AN_TRAN ? "lanes" : "sync") + std::to_string(g) + "n" + std::to_string(n) +
                 "/occ" + std::to_string(occ);
    return 0;
}

// warp-per-instance kernels (solver_warp.cuh): 16 < n_start <= 96.  Returns 0 when chosen.
int plan_warp(ftsb_circuit *c, int an, long long batch, LaunchPlan &pl)
{
    const int n_start = std::max(c->comp.n_start, 1);
    const int n = an == AN_OP ? n_start : c->comp.n_run;
    if (n_start > 96 || n < 1) return 1;
    const int rpt = (n + 31) / 32;
    KernelFn fn = warp_kernel(rpt, an);
    if (!fn) return 1;
    const WarpLayout T = warp_layout(n, c->comp.n_slots, c->comp.n_dyn, c->comp.n_src, an == AN_TRAN);
    const size_t smem = (T.bytes + 15) & ~(size_t)15;
    if (smem > c->smem_optin) return 1;
    if (cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, 32, smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        return 1;
    }
    pl.fn = fn;
    pl.g = 32;
    pl.block = 32;
    pl.smem = smem;
    pl.lay = c->comp.lay;
    pl.lay.mat_in_smem = 1;
    pl.lay.bytes = (int)smem;
    pl.scratch_bytes = 0;
    pl.lanes = true; // .tran takes its start-up solution from a separate .op launch
    pl.grid = (int)std::max(1LL, std::min((long long)batch, (long long)occ * c->sm_count));
    c->variant = "warp" + std::to_string(rpt) + "r_n" + std::to_string(n) + "/occ" + std::to_string(occ);
    return 0;
}

int plan_small(ftsb_circuit *c, int an, long long batch, LaunchPlan &pl)
{
    if (c->force_generic || c->force_cta) return 1;
    if (std::max(c->comp.n_start, 1) <= 16) return plan_lanes(c, an, batch, pl);
    return plan_warp(c, an, batch, pl);
}

int plan_launch(ftsb_circuit *c, int an, long long batch, LaunchPlan &pl)
{
    const int n = std::max(c->comp.n_start, 1);
    Layout L = c->comp.lay;
    const size_t vec_bytes = (size_t)L.doubles * 8;
    const size_t mat_bytes = (size_t)2 * L.mat_doubles * 8;
    const size_t int_bytes = (size_t)L.ints * 4;
    int g = n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : n <= 32 ? 32 : 0;
    if (c->force_cta) g = 0;
    if (plan_small(c, an, batch, pl) == 0) return 0;
    if (g) {
        size_t per = (vec_bytes + mat_bytes + int_bytes + 15) & ~(size_t)15;
        int teams = 128 / g;
        if (per * teams > c->smem_optin) g = 0; // too many slots for a sub-warp block: use a CTA
        else {
            L.mat_in_smem = 1;
            L.bytes = (int)per;
            pl.g = g;
            pl.block = 128;
            pl.smem = per * teams;
            pl.lay = L;
            pl.scratch_bytes = 0;
            KernelFn fn = pick_kernel(an, g);
            pl.fn = fn;
            cudaError_t e = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
            if (e != cudaSuccess) return fail(c, FTSB_E_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
            int occ = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, pl.block, pl.smem);
            if (e != cudaSuccess || occ < 1) return fail(c, FTSB_E_CUDA, "occupancy query failed for the sub-warp kernel");
            long long need = (batch + teams - 1) / teams;
            pl.grid = (int)std::max(1LL, std::min(need, (long long)occ * c->sm_count));
            c->variant = "warp" + std::to_string(g);
            return 0;
        }
    }
    // CTA team: one thread per row
    int block = ((n + 31) / 32) * 32;
    if (block < 64) block = 64;
    if (block > 1024) return fail(c, FTSB_E_UNSUPPORTED, "circuits with more than 1024 unknowns are not supported yet");
    size_t base = 96 * 8 + vec_bytes + int_bytes;
    L.mat_in_smem = (base + mat_bytes <= c->smem_optin) ? 1 : 0;
    if (base > c->smem_optin) return fail(c, FTSB_E_UNSUPPORTED, "per-instance vectors exceed shared memory");
    pl.g = 0;
    pl.block = block;
    pl.smem = (base + (L.mat_in_smem ? mat_bytes : 0) + 15) & ~(size_t)15;
    L.bytes = (int)pl.smem;
    pl.lay = L;
    KernelFn fn = pick_kernel(an, 0);
    pl.fn = fn;
    cudaError_t e = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return fail(c, FTSB_E_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, pl.block, pl.smem);
    if (e != cudaSuccess || occ < 1) return fail(c, FTSB_E_CUDA, "occupancy query failed for the CTA kernel");
    pl.grid = (int)std::max(1LL, std::min(batch, (long long)occ * c->sm_count));
    pl.scratch_bytes = L.mat_in_smem ? 0 : (size_t)pl.grid * mat_bytes;
    c->variant = std::string("cta") + std::to_string(block) + (L.mat_in_smem ? "" : "g");
    return 0;
}

// structure-exploiting warp-per-instance kernels (solver_sparse.cuh): n_start > 16.  Returns 0 when chosen.
int plan_sparse(ftsb_circuit *c, int an, long long batch, LaunchPlan &pl, int &fcap, int &hist_global)
{
    hist_global = 0;
    if (!c->sparse_warp || c->force_generic || c->force_cta) return 1;
    if (std::max(c->comp.n_start, 1) <= 16) return 1;
    const ModeOff &mo = c->comp.mode[an]; // AN_OP / AN_DC / AN_TRAN index the modes
    const int n = mo.n, nnz = mo.sp.nnz;
    if (n < 1) return 1;
    // run-time fill-in pool: partial pivoting may leave the predicted structure; an instance that exhausts the pool
    // goes to the dense kernels.  Halve until one warp fits.
    fcap = c->fill_cap > 0 ? c->fill_cap : std::max(32, nnz / 2);
    if (c->fill_cap <= 0 && an != AN_OP) fcap = std::min(fcap, 128); // companion conductances keep .tran / .dc pivots near the prediction
    fcap = (fcap + 1) & ~1;
    if (n >= 0xfff0) return 1; // u16 indices
    if (nnz + fcap >= 0xfff0) fcap = (0xfff0 - nnz - 2) & ~1;
    if (fcap < 2) return 1;
    const bool win = c->windowed == 2 || (c->windowed == 1 && mo.sp.window_hint >= 4);
    const size_t stat = sp_static_bytes(n, nnz, mo.sp.n_long);
    SpLayout T = sp_layout(n, nnz, fcap, c->comp.n_slots, c->comp.n_dyn, c->comp.n_src, an == AN_TRAN);
    if (an == AN_TRAN && c->hist_global_opt && T.bytes <= 32 * 1024 && T.bytes > 6 * 1024) {
        // mid-size .tran: the per-step history (touched once per accepted step) to global scratch: more warps per SM
        hist_global = 1;
        T = sp_layout(n, nnz, fcap, c->comp.n_slots, c->comp.n_dyn, c->comp.n_src, true, hist_global);
    }
    if (T.bytes > 32 * 1024) { // large circuit: per-step history (and optionally the slot table) to global scratch
        // 1: per-step history, 2: slot table, 4: x / x_new / x_proposed / b -> L2-resident global scratch; what stays in
        // shared memory is what the sequential elimination touches (matrix entries, rhs, structure)
        hist_global = (an == AN_TRAN ? 1 : 0) | (c->slots_global ? 6 : 0);
        T = sp_layout(n, nnz, fcap, c->comp.n_slots, c->comp.n_dyn, c->comp.n_src, an == AN_TRAN, hist_global);
    }
    while (stat + ((T.bytes + 15) & ~(size_t)15) > c->smem_optin && fcap > 64 && c->fill_cap <= 0) {
        fcap = (fcap / 2 + 1) & ~1;
        T = sp_layout(n, nnz, fcap, c->comp.n_slots, c->comp.n_dyn, c->comp.n_src, an == AN_TRAN, hist_global);
    }
    KernelFn fn = sparse_kernel(an, win, (hist_global & 6) ? 2 : (hist_global & 1) ? 1 : 0);
    const size_t per_warp = (T.bytes + 15) & ~(size_t)15;
    int best_w = 0, best_occ = 0;
    size_t best_smem = 0;
    for (int wpb = 4; wpb >= 1; wpb >>= 1) {
        const size_t smem = stat + per_warp * wpb;
        if (smem > c->smem_optin) continue;
        if (cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, 32 * wpb, smem) != cudaSuccess || occ < 1) {
            cudaGetLastError();
            continue;
        }
        if (occ * wpb > best_occ * best_w) {
            best_w = wpb;
            best_occ = occ;
            best_smem = smem;
        }
    }
    if (!best_w) return 1;
    if (cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)best_smem) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    pl.fn = fn;
    pl.g = 32;
    pl.block = 32 * best_w;
    pl.smem = best_smem;
    pl.lay = c->comp.lay;
    pl.lanes = true; // .tran takes its start-up solution from a separate .op launch
    const long long need = (batch + best_w - 1) / best_w;
    pl.grid = (int)std::max(1LL, std::min(need, (long long)best_occ * c->sm_count));
    pl.scratch_bytes = (size_t)pl.grid * best_w * T.hist_doubles * 8;
    c->variant = "spw_n" + std::to_string(n) + "/nnz" + std::to_string(nnz - mo.sp.n_pred) + "+" + std::to_string(mo.sp.n_pred) +
                 "/fill" + std::to_string(fcap) + (hist_global ? "/hg" + std::to_string(hist_global) : "") + (c->windowed && mo.sp.window_hint >= 4 ? "/w" + std::to_string(mo.sp.window_hint) : "") + "/occ" + std::to_string(best_occ * best_w);
    return 0;
}

int launch_kernel(ftsb_circuit *c, const LaunchPlan &pl, RunArgs &a)
{
    FTSB_CUDA(c, cudaMemsetAsync(c->work.p, 0, 8, c->stream));
    void *args[] = {&a};
    FTSB_CUDA(c, cudaLaunchKernel((const void *)pl.fn, dim3(pl.grid), dim3(pl.block), args, pl.smem, c->stream));
    c->launches++;
    return 0;
}

// ---- learned elimination plans for small circuits (plan.h, solver_plan.cuh) ---------------------------------
// First .op / .dc call of a handle: a pilot run of the dense kernel over the first PLAN_PILOT instances records the
// pivot sequence of every factorisation; the host turns the most frequent sequences into plans.
// histogram of traced pivot sequences -> plans on the device.  Returns true when usable plans exist.
int rebuild_plans(ftsb_circuit *c, int an, bool &ok)
{
    ok = false;
    const ModeOff &mo = c->comp.mode[an];
    const std::vector<int> &ib = c->comp.ints();
    std::vector<int> jdst(ib.begin() + mo.J_full.dst, ib.begin() + mo.J_full.dst + mo.J_full.n_dst);
    std::vector<std::pair<uint64_t, long long>> perms(c->plan_hist[an].begin(), c->plan_hist[an].end());
    PlanHost &ph = c->plan_host[an];
    if (!build_plans(mo.n, jdst, perms, ph, c->plan_max)) return 0;
    FTSB_CUDA(c, c->plan_buf[an].ensure(ph.ibuf.size() * 4));
    FTSB_CUDA(c, cudaMemcpyAsync(c->plan_buf[an].p, ph.ibuf.data(), ph.ibuf.size() * 4, cudaMemcpyHostToDevice, c->stream));
    FTSB_CUDA(c, cudaStreamSynchronize(c->stream)); // ph.ibuf is pageable
    ok = true;
    return 0;
}

int read_trace(ftsb_circuit *c, int an, long long rows)
{
    if (rows <= 0) return 0;
    std::vector<unsigned long long> tr((size_t)rows * PLAN_TRACE_MAX);
    FTSB_CUDA(c, cudaMemcpyAsync(tr.data(), c->trace.p, tr.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    FTSB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (unsigned long long v : tr)
        if (v) ++c->plan_hist[an][v];
    return 0;
}

int learn_plans(ftsb_circuit *c, int an, const RunArgs &a)
{
    c->plan_state[an] = -1;
    const long long sample = std::min<long long>(a.batch, PLAN_PILOT);
    LaunchPlan pd;
    if (plan_lanes(c, an, sample, pd) != 0) return 0;
    const size_t tb = (size_t)PLAN_TRACE_ROWS * PLAN_TRACE_MAX * 8;
    FTSB_CUDA(c, c->trace.ensure(tb));
    FTSB_CUDA(c, cudaMemsetAsync(c->trace.p, 0, tb, c->stream));
    // sample spread over the whole batch (list mode)
    std::vector<long long> lst((size_t)sample + 1);
    for (long long i = 0; i < sample; ++i) lst[(size_t)i] = i * a.batch / sample;
    unsigned long long cnt = (unsigned long long)sample;
    FTSB_CUDA(c, c->redo.ensure((size_t)std::max<long long>(a.batch, sample) * 8 + 8));
    FTSB_CUDA(c, cudaMemcpyAsync(c->redo.p, lst.data(), (size_t)sample * 8, cudaMemcpyHostToDevice, c->stream));
    FTSB_CUDA(c, cudaMemcpyAsync((unsigned long long *)c->counters.p + CN_REDO, &cnt, 8, cudaMemcpyHostToDevice, c->stream));
    FTSB_CUDA(c, cudaStreamSynchronize(c->stream)); // lst / cnt are pageable
    RunArgs ap = a;
    ap.prog.lay = pd.lay;
    ap.trace = (unsigned long long *)c->trace.p;
    ap.trace_rows = PLAN_TRACE_ROWS;
    ap.in_list = (const long long *)c->redo.p;
    ap.in_count = (const unsigned long long *)c->counters.p + CN_REDO;
    int rc = launch_kernel(c, pd, ap);
    if (rc) return rc;
    rc = read_trace(c, an, sample);
    if (rc) return rc;
    // the pilot's counters must not leak into the caller's run
    FTSB_CUDA(c, cudaMemsetAsync(c->counters.p, 0, CN_COUNT * 8, c->stream));
    bool ok = false;
    rc = rebuild_plans(c, an, ok);
    if (rc) return rc;
    if (ok) c->plan_state[an] = 1;
    c->relearn_left[an] = 3;
    return 0;
}

// Instances the previous run handed to the dense kernel were traced: add their pivot sequences to the plans.
int relearn_plans(ftsb_circuit *c, int an)
{
    if (c->plan_state[an] != 1 || c->relearn_left[an] <= 0 || !c->redo_traced[an]) return 0;
    c->redo_traced[an] = false;
    unsigned long long n_redo = 0;
    FTSB_CUDA(c, cudaMemcpyAsync(&n_redo, (unsigned long long *)c->counters.p + CN_REDO, 8, cudaMemcpyDeviceToHost, c->stream));
    FTSB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (n_redo == 0) {
        c->relearn_left[an] = 0; // the plans cover the workload
        return 0;
    }
    c->relearn_left[an]--;
    const size_t before = c->plan_hist[an].size();
    int rc = read_trace(c, an, std::min<long long>((long long)n_redo, PLAN_TRACE_ROWS));
    if (rc) return rc;
    if (c->plan_hist[an].size() == before) return 0; // nothing new (e.g. zero pivots: the dense kernel's job)
    const int k_before = c->plan_host[an].K;
    bool ok = false;
    rc = rebuild_plans(c, an, ok);
    if (rc) return rc;
    if (!ok) {
        c->plan_state[an] = -1;
        return 0;
    }
    (void)k_before; // new plans (or a new order): the specialised kernel is regenerated on demand
    FTSB_CUDA(c, cudaStreamSynchronize(c->stream));
    jit_unload(c->jit[an]);
    c->jit_state[an] = 0;
    return 0;
}

// Returns 0 when the plan kernel was chosen.
int plan_plans(ftsb_circuit *c, int an, long long batch, LaunchPlan &pl, PlanProg &pp)
{
    if (an != AN_OP && an != AN_DC) return 1;
    if (!c->plan_opt || c->force_generic || c->force_cta || c->sparse || c->lanes > 0) return 1;
    if (c->plan_state[an] != 1) return 1;
    const PlanHost &ph = c->plan_host[an];
    const ModeOff &mo = c->comp.mode[an];
    const int *base = (const int *)c->plan_buf[an].p;
    pp.K = ph.K;
    pp.S = ph.S;
    pp.nnz0 = ph.nnz0;
    pp.n = ph.n;
    pp.rec = reinterpret_cast<const int4 *>(base + ph.off_rec);
    pp.cand = base + ph.off_cand;
    pp.mult = base + ph.off_mult;
    pp.upd = base + ph.off_upd;
    pp.bu = base + ph.off_bu;
    pp.next = reinterpret_cast<const signed char *>(base + ph.off_next);
    pp.jdst = base + ph.off_jdst;
    const int g = c->plan_lanes == 1 || c->plan_lanes == 2 || c->plan_lanes == 4 ? c->plan_lanes : (mo.n <= 4 ? 1 : mo.n <= 8 ? 2 : 4);
    KernelFn fn = plan_kernel(g, an);
    const PlanLayout T = plan_layout(mo.n, ph.S, c->comp.n_slots, mo.n_nl == 0);
    const size_t smem = (size_t)T.doubles * (32 / g) * 8;
    if (smem > c->smem_optin) return 1;
    if (cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, 32, smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        return 1;
    }
    pl.fn = fn;
    pl.g = g;
    pl.block = 32;
    pl.smem = smem;
    pl.lay = c->comp.lay;
    pl.scratch_bytes = 0;
    pl.lanes = true;
    const int per_warp = 32 / g;
    pl.grid = (int)std::max(1LL, std::min((batch + per_warp - 1) / per_warp, (long long)occ * c->sm_count));
    c->variant = "plan" + std::to_string(ph.K) + "g" + std::to_string(g) + "n" + std::to_string(mo.n) + "/S" + std::to_string(ph.S) +
                 "/occ" + std::to_string(occ);
    return 0;
}

// per-topology kernel for (handle, mode): generate, compile and load once.  Returns true when it can be launched.
bool ensure_jit(ftsb_circuit *c, int an)
{
    if (!c->jit_opt || !c->skip_refactor || c->plan_state[an] != 1) return false;
    if (c->jit_state[an] != 0) return c->jit_state[an] == 1;
    c->jit_state[an] = -1;
    const PlanHost &ph = c->plan_host[an];
    const std::string src = jit_source(c->comp, an, ph, c->jit_log1p) + "// min_blocks " + std::to_string(c->jit_min_blocks) + " max_regs " + std::to_string(c->jit_max_regs) + "\n";
    std::string cubin, err;
    {
        // the same topology + plans give the same source: compile once per process
        static std::mutex mu;
        static std::map<std::string, std::string> cache;
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(src);
        if (it != cache.end())
            cubin = it->second;
        else {
            if (!jit_compile(src, cubin, err, c->jit_min_blocks, c->jit_max_regs)) {
                c->jit_msg = err;
                return false;
            }
            cache[src] = cubin;
        }
    }
    const int smem = jit_smem_bytes(ph);
    if ((size_t)smem > c->smem_optin || !jit_load(cubin, smem, c->jit[an], err)) {
        c->jit_msg = err;
        return false;
    }
    c->jit_occ[an] = jit_occupancy(c->jit[an]);
    if (c->jit_occ[an] < 1) {
        c->jit_msg = "occupancy query failed";
        jit_unload(c->jit[an]);
        return false;
    }
    c->jit_state[an] = 1;
    return true;
}

// Launches the per-topology kernel over instances [begin, end) on `stream`: `work` (already holding `begin`) hands out
// global instance indices and the kernel's batch bound is `end`, so several ranges can be in flight on different
// streams with one shared redo list and one set of counters.
int launch_jit_range(ftsb_circuit *c, int an, RunArgs &a, cudaStream_t stream, unsigned long long *work, long long begin,
                     long long end, int stop_it = 0, int resume = 0)
{
    const std::vector<int> &ib = c->comp.ints();
    int sweep_slot = an == AN_DC ? ib[c->comp.off_src + 5 * a.sweep_src + 1] : 0;
    const double *vals = a.vals, *fn = a.fn, *st = a.start, *sp = a.stop, *se = a.step;
    long long batch = end, max_points = a.max_points;
    void *params[] = {&vals, &fn, &batch, &work, &a.counters, &a.x, &a.n_iters, &a.status, &a.redo_list, &a.redo_count,
                      &st, &sp, &se, &max_points, &a.points_done, &sweep_slot, &stop_it, &resume, &a.vs_inst, &a.vs_elem,
                      &a.fs_inst, &a.fs_elem};
    // every resident slot: lanes refill as their instances finish, so 65 536 instances on 37 888 lanes take 1.73 instance times
    // (a grid of 1 024 warps = two whole rounds per lane measured 0.82 ms against 0.705 ms)
    const int grid = (int)std::max(1LL, std::min((end - begin + 31) / 32, (long long)c->jit_occ[an] * c->sm_count));
    std::string err;
    if (!jit_launch(c->jit[an], grid, (void *)stream, params, err)) return fail(c, FTSB_E_CUDA, err);
    c->launches++;
    return 0;
}

// Two-pass .op (option jit_two_pass = k > 0, nonlinear circuits): pass A runs Newton iterations 0..k-1 of every instance
// with all lanes of a warp in the same iteration — the only iterations whose pivot sequences vary, so the sibling-plan
// bodies run with many lanes instead of dragging a whole warp along for one freshly refilled lane — and pauses the
// instances (status -1, x and the iteration count in the output arrays); pass B continues them with per-lane refill.
// Same arithmetic: the continuation re-evaluates the devices at the paused iterate, exactly what the next iteration
// of a single pass does.
bool jit_two_pass_applies(const ftsb_circuit *c, int an)
{
    // worth a second launch only when several plans are live (measured: K = 6 -21 %, K = 4 -4 %, K <= 2 +2..10 %)
    return c->jit_two_pass > 0 && an == AN_OP && c->comp.mode[an].n_nl > 0 && c->plan_host[an].K >= c->jit_two_pass_min_plans;
}

int launch_jit(ftsb_circuit *c, int an, RunArgs &a)
{
    FTSB_CUDA(c, cudaMemsetAsync(c->work.p, 0, 8, c->stream));
    if (!jit_two_pass_applies(c, an)) return launch_jit_range(c, an, a, c->stream, a.work, 0, a.batch);
    for (int j = 0; j < c->jit_passes; ++j) { // pass j: iterations j k .. (j + 1) k - 1, the last pass to the end
        if (j) FTSB_CUDA(c, cudaMemsetAsync(c->work.p, 0, 8, c->stream));
        int rc = launch_jit_range(c, an, a, c->stream, a.work, 0, a.batch, j + 1 < c->jit_passes ? (j + 1) * c->jit_two_pass : 0, j > 0);
        if (rc) return rc;
    }
    return 0;
}

// The global scratch rows of the warp / team kernels (slot tables, vectors, per-step history of the resident instances) are
// re-read and re-written every Newton iteration / accepted step.  They fit in L2 (tens of MB), but ordinary lines are
// evicted by the streaming parameter and waveform traffic and written back to HBM over and over (round 1, C5: 7.3 GB of
// DRAM writes per 0.28 s launch for 3 MB of results).  An access-policy window marks the scratch as persisting in a
// set-aside part of L2; everything else streams past it.  Best effort: failures are ignored.
void l2_persist_window(ftsb_circuit *c, void *base, size_t bytes)
{
    if (!c->l2_queried) {
        c->l2_queried = true;
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, c->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
            const size_t want = (size_t)prop.persistingL2CacheMaxSize;
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
                c->l2_set_aside = want;
                c->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
            }
        }
        cudaGetLastError();
    }
    if (!c->l2_set_aside || !c->l2_window_max) return;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    if (c->l2_persist && base && bytes) {
        const size_t win = std::min(bytes, c->l2_window_max);
        attr.accessPolicyWindow.base_ptr = base;
        attr.accessPolicyWindow.num_bytes = win;
        attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)c->l2_set_aside / (double)win);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } // else: num_bytes = 0 switches the window off
    cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
}

// ---- per-team .tran kernels (solver_team.cuh) ----------------------------------------------------------------
struct TeamLaunch {
    TeamKernelFn fn = nullptr;
    int g = 0, place = 0, grid = 1, wpb = 4;
    size_t smem = 0, scratch_bytes = 0;
    TeamArgs ta{};
};

// Returns 0 when the per-team kernel was chosen for this .tran run.
int plan_team(ftsb_circuit *c, int an, long long batch, TeamLaunch &tl)
{
    if (an != AN_TRAN || !c->team_opt || !c->sparse_warp || !c->skip_refactor || c->force_generic || c->force_cta) return 1;
    if (std::max(c->comp.n_start, 1) <= 16) return 1;
    const ModeOff &mo = c->comp.mode[an];
    if (mo.n < 1 || !mo.sp.piv_valid) return 1;
    if (c->team_state == 0) {
        c->team_state = -1;
        const std::vector<int> &ib = c->comp.ints();
        std::vector<int> piv(ib.begin() + mo.sp.piv, ib.begin() + mo.sp.piv + mo.n);
        if (!build_team_plan(c->comp, an, piv, c->team_plan)) return 1;
        const TeamPlanHost &ph = c->team_plan;
        if (c->team_buf.ensure(ph.ibuf.size() * 4) != cudaSuccess) {
            cudaGetLastError();
            return 1;
        }
        if (cudaMemcpyAsync(c->team_buf.p, ph.ibuf.data(), ph.ibuf.size() * 4, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) {
            cudaGetLastError();
            return 1;
        }
        c->team_state = 1;
    }
    if (c->team_state != 1) return 1;
    const TeamPlanHost &ph = c->team_plan;
    const bool linear = mo.n_nl == 0;
    // placement: a linear circuit touches its slot table and factors once per attempt (global scratch); a nonlinear one in
    // every Newton iteration (shared memory: placement 3).  A circuit too large for that (n = 258: 54 KB per instance) runs
    // with only the factors and the vectors in shared memory (placement 6: slot table and entries of A in the scratch).  With
    // one lane walking the ~1 550 dependent ops of a Newton iteration that was slower than the warp kernel's windowed
    // elimination on the C5 chain (0.22 vs 0.28 M solve-points/s on 4 096 instances x 4 ns); with the level schedules, the
    // linear-prefix sums and the forward substitution merged into the elimination it is faster (0.39 M) and is tried
    // automatically when placement 3 does not fit
    int place = 0;
    // lanes per instance: rows are dealt round-robin, so few lanes waste little; more teams per warp amortise the
    // list-walking code over more instances but need more shared memory per warp
    int best_g = 0, best_occ = 0, best_wpb = 0, best_rows = 0;
    size_t best_smem = 0;
    TeamLayout best_lay{};
    for (int place_try : {c->team_place >= 0 ? (c->team_place & 7) : (linear ? 0 : 3), c->team_place >= 0 || linear ? -1 : 6, -1}) {
    if (place_try < 0 || best_g) break;
    place = place_try;
    for (int g : {8, 16, 32}) {
        if (c->team_g > 0) g = c->team_g;
        TeamKernelFn fn = (mo.n + g - 1) / g <= 32 ? team_kernel(g, place, ph.ellW) : nullptr; // rows per lane fit a 32-bit class mask
        // linear circuit, rows per lane among the instantiated counts: registers instead of shared-memory round trips
        int rows = 0;
        if (linear && ph.chain && g == 8 && (place & 2) == 0 && c->team_rows && (mo.n + g - 1) / g <= 32) {
            TeamKernelFn fr = team_kernel_rows(g, place, ph.ellW, (mo.n + g - 1) / g);
            if (fr) {
                fn = fr;
                rows = (mo.n + g - 1) / g;
            }
        }
        if (fn) {
            const int T = 32 / g;
            const TeamLayout lay = team_layout(mo.n, ph.nSlotA, ph.S, c->comp.n_slots, c->comp.n_dyn, c->comp.n_src, place, linear ? 0 : ph.nPre);
            int w_best = 0, occ_best = 0;
            size_t smem_best = 0;
            for (int wpb = rows ? TEAM_WARPS_ROWS : TEAM_WARPS_MAX; wpb >= 1; --wpb) { // warps per block: whatever packs most warps onto an SM
                const size_t smem = (size_t)2 * FTS_LOG1P_N * 8 + (size_t)wpb * lay.sm_doubles * T * 8;
                int occ = 0;
                if (smem <= c->smem_optin &&
                    cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, 32 * wpb, smem) == cudaSuccess && occ >= 1) {
                    if (occ * wpb > occ_best * w_best) {
                        w_best = wpb;
                        occ_best = occ;
                        smem_best = smem;
                    }
                } else
                    cudaGetLastError();
            }
            // automatic: the smallest lane count that keeps two warps per scheduler resident; a circuit too large for
            // that (n = 258: 50 KB per instance) stays with the warp kernel and its global scratch rows
            if (w_best && (c->team_g > 0 || occ_best * w_best >= 8)) {
                best_g = g;
                best_rows = rows;
                best_occ = occ_best;
                best_wpb = w_best;
                best_smem = smem_best;
                best_lay = lay;
            }
        }
        if (c->team_g > 0 || best_g) break;
    }
    }
    if (!best_g) return 1;
    const int T = 32 / best_g;
    tl.fn = best_rows ? team_kernel_rows(best_g, place, ph.ellW, best_rows) : team_kernel(best_g, place, ph.ellW);
    tl.g = best_g;
    tl.place = place;
    tl.wpb = best_wpb;
    tl.smem = best_smem;
    const long long per_block = (long long)best_wpb * T;
    tl.grid = (int)std::max(1LL, std::min((batch + per_block - 1) / per_block, (long long)best_occ * c->sm_count));
    tl.scratch_bytes = (size_t)tl.grid * best_wpb * best_lay.gl_doubles * T * 8;
    tl.ta.plan = ph.view((const int *)c->team_buf.p);
    tl.ta.plan.levels = c->team_levels;
    tl.ta.lay = best_lay;
    tl.ta.flops_factor = ph.flops_factor;
    tl.ta.flops_solve = ph.flops_solve;
    if (cudaFuncSetAttribute((const void *)tl.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.smem) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    c->variant = "team" + std::to_string(best_g) + "p" + std::to_string(place) + (best_rows ? "r" + std::to_string(best_rows) : std::string()) + "_n" + std::to_string(mo.n) + "/A" + std::to_string(ph.nA) +
                 "/S" + std::to_string(ph.S) + (ph.chain ? "/chain" : "") + "/occ" + std::to_string(best_occ) + "x" + std::to_string(best_wpb);
    return 0;
}

int launch_team(ftsb_circuit *c, TeamLaunch &tl, RunArgs &a)
{
    FTSB_CUDA(c, c->team_scratch.ensure(tl.scratch_bytes));
    tl.ta.gscratch = (double *)c->team_scratch.p;
    l2_persist_window(c, c->team_scratch.p, tl.scratch_bytes);
    FTSB_CUDA(c, cudaMemsetAsync(c->work.p, 0, 8, c->stream));
    void *args[] = {&a, &tl.ta};
    FTSB_CUDA(c, cudaLaunchKernel((const void *)tl.fn, dim3(tl.grid), dim3(32 * tl.wpb), args, tl.smem, c->stream));
    c->launches++;
    return 0;
}

// one analysis over the batch: the sparse kernel followed by the dense kernels on the instances it abandoned, or
// the dense / small-circuit kernels alone.  `dense_lanes`: the dense plan takes an external start-up solution.
int run_phase(ftsb_circuit *c, int an, RunArgs a, std::string &variant)
{
    LaunchPlan ps, pd;
    int fcap = 0, hist_global = 0;
    bool sparse = plan_sparse(c, an, a.batch, ps, fcap, hist_global) == 0;
    bool planned = false, use_jit = false;
    if (!sparse && (an == AN_OP || an == AN_DC) && std::max(c->comp.n_start, 1) <= 16 && c->plan_opt && !c->force_generic &&
        !c->force_cta && !c->sparse && c->lanes <= 0 && a.batch >= 64) {
        if (c->plan_state[an] == 0) {
            int rcl = learn_plans(c, an, a);
            if (rcl) return rcl;
        }
        planned = plan_plans(c, an, a.batch, ps, a.plan) == 0;
        sparse = planned; // same launch protocol: first kernel + redo list + dense pass
        if (planned) c->last_plan_an = an;
        if (planned && ensure_jit(c, an)) {
            use_jit = true;
            const PlanHost &ph = c->plan_host[an];
            c->variant = "jit" + std::to_string(ph.K) + "n" + std::to_string(ph.n) + "/S" + std::to_string(ph.S) + "/r" +
                         std::to_string(c->jit[an].regs) + "/occ" + std::to_string(c->jit_occ[an]);
        }
    }
    const std::string sv = c->variant;
    int rc = plan_launch(c, an, a.batch, pd);
    if (rc) return rc;
    const std::string dv = c->variant;
    {
        const size_t need = std::max(pd.scratch_bytes, sparse ? ps.scratch_bytes : (size_t)0);
        if (need) FTSB_CUDA(c, c->scratch.ensure(need));
    }
    a.scratch = (double *)c->scratch.p;
    std::string tv;
    if (sparse && !planned) { // .tran, n > 16: the per-team kernel first; what leaves its plan goes down the chain below
        TeamLaunch tl;
        if (plan_team(c, an, a.batch, tl) == 0) {
            tv = c->variant;
            FTSB_CUDA(c, c->redo2.ensure((size_t)a.batch * 8));
            RunArgs at = a;
            at.redo_list = (long long *)c->redo2.p;
            at.redo_count = (unsigned long long *)c->counters.p + CN_REDO2;
            FTSB_CUDA(c, cudaMemsetAsync(at.redo_count, 0, 8, c->stream));
            rc = launch_team(c, tl, at);
            if (rc) return rc;
            a.in_list = at.redo_list;
            a.in_count = at.redo_count;
        }
    }
    if (sparse) {
        FTSB_CUDA(c, c->redo.ensure((size_t)a.batch * 8));
        RunArgs as = a;
        as.prog.lay = ps.lay;
        as.prog.op.sp.fcap = as.prog.dc.sp.fcap = as.prog.tran.sp.fcap = fcap;
        as.prog.op.sp.hist_global = as.prog.dc.sp.hist_global = as.prog.tran.sp.hist_global = hist_global;
        // windows pay when the predicted structure lets several columns go at once (a rail-coupled chain: 32; a ladder: 1)
        as.prog.op.sp.windowed = as.prog.dc.sp.windowed = as.prog.tran.sp.windowed =
            (c->windowed == 2 || (c->windowed == 1 && c->comp.mode[an].sp.window_hint >= 4)) ? 1 : 0;
        as.redo_list = (long long *)c->redo.p;
        as.redo_count = (unsigned long long *)c->counters.p + CN_REDO;
        FTSB_CUDA(c, cudaMemsetAsync(as.redo_count, 0, 8, c->stream));
        if (c->last_redo_an >= 0 && c->last_redo_an != an && c->last_redo_an < 2)
            c->redo_traced[c->last_redo_an] = false; // its redo count is gone: do not relearn from a stale counter
        if (!planned && ps.scratch_bytes) l2_persist_window(c, c->scratch.p, ps.scratch_bytes);
        rc = use_jit ? launch_jit(c, an, as) : launch_kernel(c, ps, as);
        if (rc) return rc;
        if (!planned && ps.scratch_bytes) l2_persist_window(c, nullptr, 0);
        a.in_list = as.redo_list;
        a.in_count = as.redo_count;
        if (planned && c->relearn_left[an] > 0) { // the dense redo pass records what the plans did not cover
            FTSB_CUDA(c, cudaMemsetAsync(c->trace.p, 0, (size_t)PLAN_TRACE_ROWS * PLAN_TRACE_MAX * 8, c->stream));
            a.trace = (unsigned long long *)c->trace.p;
            a.trace_rows = PLAN_TRACE_ROWS;
            c->redo_traced[an] = true;
        }
        c->last_redo_an = an;
    }
    a.prog.lay = pd.lay;
    rc = launch_kernel(c, pd, a);
    if (rc) return rc;
    variant = sparse ? sv + "+redo:" + dv : dv;
    if (!tv.empty()) variant = tv + "+redo:" + variant;
    return 0;
}

int launch(ftsb_circuit *c, int an, RunArgs &a)
{
    std::lock_guard<std::mutex> device_lock(device_launch_mutex(c->device));
    a.prog = c->prog;
    a.vs_inst = c->vs_inst;
    a.vs_elem = c->vs_elem;
    a.fs_inst = c->fs_inst;
    a.fs_elem = c->fs_elem;
    a.work = (unsigned long long *)c->work.p;
    a.counters = (unsigned long long *)c->counters.p;
    a.skip_refactor = c->skip_refactor;
    if ((an == AN_OP || an == AN_DC) && c->last_redo_an == an) { // before the counters are cleared: did the plans miss instances?
        int rcl = relearn_plans(c, an);
        if (rcl) return rcl;
    }
    c->last_redo_an = -1;
    FTSB_CUDA(c, cudaMemsetAsync(c->counters.p, 0, CN_COUNT * 8, c->stream));
    std::string v_op, v_main;
    bool ext_start = false;
    c->last_plan_an = -1;
    if (an == AN_TRAN) { // does the .tran kernel take its start-up solution (engine.rs:158-161) from an .op launch?
        LaunchPlan t;
        int fc = 0, hg = 0;
        if (plan_sparse(c, an, a.batch, t, fc, hg) == 0)
            ext_start = true;
        else {
            int rc = plan_launch(c, an, a.batch, t);
            if (rc) return rc;
            ext_start = t.lanes;
        }
    }
    if (ext_start) {
        const int ns = std::max(c->comp.n_start, 1);
        FTSB_CUDA(c, c->s_x.ensure((size_t)a.batch * ns * 8));
        FTSB_CUDA(c, c->s_it.ensure((size_t)a.batch * 4));
        FTSB_CUDA(c, c->s_st.ensure((size_t)a.batch * 4));
        RunArgs ao = a;
        ao.x = (double *)c->s_x.p;
        ao.n_iters = (int *)c->s_it.p;
        ao.status = (int *)c->s_st.p;
        int rc = run_phase(c, AN_OP, ao, v_op);
        if (rc) return rc;
        // the start-up solve is not an output row of .tran
        FTSB_CUDA(c, cudaMemsetAsync((unsigned long long *)c->counters.p + CN_POINTS, 0, 8, c->stream));
        a.x_start = (const double *)c->s_x.p;
        a.start_status = (const int *)c->s_st.p;
        a.n_start = ns;
    }
    int rc = run_phase(c, an, a, v_main);
    if (rc) return rc;
    c->variant = ext_start ? v_main + "+op[" + v_op + "]" : v_main;
    return 0;
}

// the generated kernels depend on an option that changed: wait for work in flight (an asynchronous FTSB_MEM_DEVICE run may
// still execute the module), unload, regenerate on demand
void jit_reset(ftsb_circuit *c)
{
    if (cudaSetDevice(c->device) == cudaSuccess && c->stream) cudaStreamSynchronize(c->stream);
    for (int an = 0; an < 2; ++an) {
        jit_unload(c->jit[an]);
        c->jit_state[an] = 0;
    }
}

int check_ready(ftsb_circuit *c)
{
    if (!c) return fail(nullptr, FTSB_E_ARG, "null handle");
    if (c->batch <= 0) return fail(c, FTSB_E_STATE, "ftsb_set_params must be called before a run");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(c, FTSB_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return 0;
}

} // namespace

// ---------------------------------------------------------------------------------------------------------
extern "C" {

int ftsb_abi_version(void) { return FTSB_ABI_VERSION; }

int ftsb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *ftsb_last_error(const ftsb_circuit *c) { return c ? c->err.c_str() : g_last_error.c_str(); }

int ftsb_compile(const ftsb_elem *elems, int32_t n_elems, int32_t n_run, int32_t n_start, const int32_t *row_class,
                 int device, ftsb_circuit **out)
{
    if (!out) return fail(nullptr, FTSB_E_ARG, "ftsb_compile: out is NULL");
    *out = nullptr;
    if (n_elems > 0 && !elems) return fail(nullptr, FTSB_E_ARG, "ftsb_compile: elems is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, FTSB_E_CUDA,
                    std::string("ftsb_compile: no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                        "); this library has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, FTSB_E_ARG, "ftsb_compile: bad device ordinal");
    static_assert(sizeof(ftsb_elem) == sizeof(ElemDesc), "ABI element layout");
    ftsb_circuit *c = new ftsb_circuit();
    c->device = device;
    std::string err;
    const int ld_min = 0;
    if (!compile_circuit(reinterpret_cast<const ElemDesc *>(elems), n_elems, n_run, n_start, row_class, ld_min, c->comp, err)) {
        delete c;
        return fail(nullptr, FTSB_E_ARG, err);
    }
    auto bail = [&](const std::string &m, int code) {
        g_last_error = m;
        ftsb_destroy(c);
        return code;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(std::string("cudaSetDevice: ") + cudaGetErrorString(e), FTSB_E_CUDA);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return bail(std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e), FTSB_E_CUDA);
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(std::string("cudaStreamCreate: ") + cudaGetErrorString(e), FTSB_E_CUDA);
    c->own_stream = true;
    const std::vector<int> &ib = c->comp.ints();
    if ((e = c->ibuf.ensure(std::max<size_t>(ib.size(), 2) * 4)) != cudaSuccess)
        return bail(std::string("cudaMalloc: ") + cudaGetErrorString(e), FTSB_E_CUDA);
    if (!ib.empty() && (e = cudaMemcpy(c->ibuf.p, ib.data(), ib.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
        return bail(std::string("cudaMemcpy: ") + cudaGetErrorString(e), FTSB_E_CUDA);
    if ((e = c->work.ensure(8)) != cudaSuccess || (e = c->counters.ensure(CN_COUNT * 8)) != cudaSuccess)
        return bail(std::string("cudaMalloc: ") + cudaGetErrorString(e), FTSB_E_CUDA);
    cudaMemset(c->counters.p, 0, CN_COUNT * 8);
    const int *base = (const int *)c->ibuf.p;
    DevProg &P = c->prog;
    P.n_elems = c->comp.n_elems;
    P.n_slots = c->comp.n_slots;
    P.n_fn = c->comp.n_fn;
    P.n_res = c->comp.n_res;
    P.n_src = c->comp.n_src;
    P.n_dyn = c->comp.n_dyn;
    P.res = base + c->comp.off_res;
    P.src = base + c->comp.off_src;
    P.dyn = base + c->comp.off_dyn;
    P.src_of_elem = base + c->comp.off_src_of_elem;
    P.op = dev_mode(base, c->comp.mode[0]);
    P.dc = dev_mode(base, c->comp.mode[1]);
    P.tran = dev_mode(base, c->comp.mode[2]);
    P.lay = c->comp.lay;
    *out = c;
    return 0;
}

void ftsb_destroy(ftsb_circuit *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream && c->own_stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    if (c->pipe_ready) {
        for (int k = 0; k < ftsb_circuit::PIPE_MAX; ++k) {
            cudaStreamSynchronize(c->pipe[k]);
            cudaStreamDestroy(c->pipe[k]);
            cudaEventDestroy(c->pipe_ev[k]);
        }
        cudaEventDestroy(c->pipe_ev0);
    }
    c->pipe_work.release();
    DevBuf *bufs[] = {&c->ibuf, &c->vals, &c->fn, &c->tmp_in, &c->work, &c->counters, &c->scratch, &c->o_x, &c->o_it,
                      &c->o_status, &c->o_done, &c->o_t, &c->o_nrec, &c->o_xf, &c->i_start, &c->i_stop, &c->i_step, &c->s_x, &c->s_it,
                      &c->s_st, &c->redo, &c->plan_buf[0], &c->plan_buf[1], &c->trace, &c->team_buf, &c->team_scratch,
                      &c->redo2};
    for (DevBuf *b : bufs) b->release();
    jit_unload(c->jit[0]);
    jit_unload(c->jit[1]);
    delete c;
}

int32_t ftsb_n_elems(const ftsb_circuit *c) { return c ? c->comp.n_elems : -1; }
int32_t ftsb_n_run(const ftsb_circuit *c) { return c ? c->comp.n_run : -1; }
int32_t ftsb_n_start(const ftsb_circuit *c) { return c ? c->comp.n_start : -1; }
int32_t ftsb_n_fn(const ftsb_circuit *c) { return c ? c->comp.n_fn : -1; }

int ftsb_set_option(ftsb_circuit *c, const char *name, int64_t value)
{
    if (!c || !name) return fail(c, FTSB_E_ARG, "ftsb_set_option: null argument");
    if (!strcmp(name, "skip_refactor"))
        c->skip_refactor = value ? 1 : 0;
    else if (!strcmp(name, "force_cta"))
        c->force_cta = value ? 1 : 0;
    else if (!strcmp(name, "force_generic"))
        c->force_generic = value ? 1 : 0;
    else if (!strcmp(name, "lanes"))
        c->lanes = (int)value;
    else if (!strcmp(name, "sparse"))
        c->sparse = value ? 1 : 0;
    else if (!strcmp(name, "plan"))
        c->plan_opt = value ? 1 : 0;
    else if (!strcmp(name, "plan_lanes"))
        c->plan_lanes = (int)value;
    else if (!strcmp(name, "jit"))
        c->jit_opt = value ? 1 : 0;
    else if (!strcmp(name, "jit_min_blocks")) {
        c->jit_min_blocks = (int)value;
        jit_reset(c);
    }
    else if (!strcmp(name, "sparse_warp"))
        c->sparse_warp = value ? 1 : 0;
    else if (!strcmp(name, "l2_persist"))
        c->l2_persist = value ? 1 : 0;
    else if (!strcmp(name, "team"))
        c->team_opt = value ? 1 : 0;
    else if (!strcmp(name, "team_lanes"))
        c->team_g = (int)value;
    else if (!strcmp(name, "team_levels"))
        c->team_levels = value ? 1 : 0;
    else if (!strcmp(name, "team_rows"))
        c->team_rows = value ? 1 : 0;
    else if (!strcmp(name, "team_place"))
        c->team_place = (int)value;
    else if (!strcmp(name, "hist_global"))
        c->hist_global_opt = value ? 1 : 0;
    else if (!strcmp(name, "windowed"))
        c->windowed = (int)value;
    else if (!strcmp(name, "slots_global"))
        c->slots_global = value ? 1 : 0;
    else if (!strcmp(name, "fill_cap"))
        c->fill_cap = (int)value;
    else if (!strcmp(name, "jit_two_pass"))
        c->jit_two_pass = (int)std::max<int64_t>(0, std::min<int64_t>(value, 99));
    else if (!strcmp(name, "jit_two_pass_min_plans"))
        c->jit_two_pass_min_plans = (int)std::max<int64_t>(1, value);
    else if (!strcmp(name, "jit_passes"))
        c->jit_passes = (int)std::max<int64_t>(2, std::min<int64_t>(value, ftsb_circuit::PASS_MAX));
    else if (!strcmp(name, "jit_max_regs")) {
        c->jit_max_regs = (int)std::max<int64_t>(0, std::min<int64_t>(value, 255));
        jit_reset(c);
    }
    else if (!strcmp(name, "plan_max"))
        c->plan_max = (int)std::max<int64_t>(1, std::min<int64_t>(value, PLAN_MAX));
    else if (!strcmp(name, "jit_log1p")) {
        c->jit_log1p = (int)std::max<int64_t>(0, std::min<int64_t>(value, 3));
        jit_reset(c);
    } else if (!strcmp(name, "pipe_chunks"))
        c->pipe_chunks = (int)std::max<int64_t>(1, std::min<int64_t>(value, ftsb_circuit::PIPE_MAX));
    else
        return fail(c, FTSB_E_ARG, std::string("ftsb_set_option: unknown option ") + name);
    return 0;
}

int ftsb_set_params(ftsb_circuit *c, int64_t batch, const double *vals, const double *fn, int layout, int mem)
{
    if (!c) return fail(nullptr, FTSB_E_ARG, "null handle");
    if (batch <= 0) return fail(c, FTSB_E_ARG, "ftsb_set_params: batch must be > 0");
    const int ne = c->comp.n_elems, nf = c->comp.n_fn;
    if (ne > 0 && !vals) return fail(c, FTSB_E_ARG, "ftsb_set_params: vals is NULL");
    if (nf > 0 && !fn) return fail(c, FTSB_E_ARG, "ftsb_set_params: fn is NULL but the circuit has waveform sources");
    if (layout != FTSB_LAYOUT_INSTANCE_MAJOR && layout != FTSB_LAYOUT_PARAM_MAJOR)
        return fail(c, FTSB_E_ARG, "ftsb_set_params: bad layout");
    FTSB_CUDA(c, cudaSetDevice(c->device));
    const size_t vb = (size_t)batch * std::max(ne, 1) * 8, fb = (size_t)batch * std::max(nf, 1) * 7 * 8;
    FTSB_CUDA(c, c->vals.ensure(vb));
    FTSB_CUDA(c, c->fn.ensure(fb));
    const cudaMemcpyKind kind = mem == FTSB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    // either layout is kept as it comes: the kernels address element e of instance i as p[i * s_inst + e * s_elem]
    auto put = [&](DevBuf &dst, const double *src, long long inner) -> int {
        if (inner == 0 || !src) return 0;
        FTSB_CUDA(c, cudaMemcpyAsync(dst.p, src, (size_t)batch * inner * 8, kind, c->stream));
        return 0;
    };
    int rc = put(c->vals, vals, ne);
    if (rc) return rc;
    rc = put(c->fn, fn, (long long)nf * 7);
    if (rc) return rc;
    if (mem == FTSB_MEM_HOST) FTSB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->batch = batch;
    if (layout == FTSB_LAYOUT_INSTANCE_MAJOR) {
        c->vs_inst = ne;
        c->vs_elem = 1;
        c->fs_inst = (long long)nf * 7;
        c->fs_elem = 1;
    } else {
        c->vs_inst = 1;
        c->vs_elem = batch;
        c->fs_inst = 1;
        c->fs_elem = batch;
    }
    return 0;
}

int ftsb_run_op(ftsb_circuit *c, double *x, int32_t *n_iters, int32_t *status, int mem)
{
    int rc = check_ready(c);
    if (rc) return rc;
    if (!x || !n_iters || !status) return fail(c, FTSB_E_ARG, "ftsb_run_op: null output");
    const long long B = c->batch;
    const int n = c->comp.n_start;
    RunArgs a{};
    a.batch = B;
    a.vals = (const double *)c->vals.p;
    a.fn = (const double *)c->fn.p;
    if (mem == FTSB_MEM_HOST) {
        FTSB_CUDA(c, c->o_x.ensure((size_t)B * std::max(n, 1) * 8));
        FTSB_CUDA(c, c->o_it.ensure((size_t)B * 4));
        FTSB_CUDA(c, c->o_status.ensure((size_t)B * 4));
        a.x = (double *)c->o_x.p;
        a.n_iters = (int *)c->o_it.p;
        a.status = (int *)c->o_status.p;
    } else {
        a.x = x;
        a.n_iters = n_iters;
        a.status = status;
    }
    rc = launch(c, AN_OP, a);
    if (rc) return rc;
    if (mem == FTSB_MEM_HOST) {
        FTSB_CUDA(c, cudaMemcpyAsync(x, a.x, (size_t)B * n * 8, cudaMemcpyDeviceToHost, c->stream));
        FTSB_CUDA(c, cudaMemcpyAsync(n_iters, a.n_iters, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream));
        FTSB_CUDA(c, cudaMemcpyAsync(status, a.status, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream));
        FTSB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int ftsb_run_dc(ftsb_circuit *c, int32_t sweep_elem, const double *start, const double *stop, const double *step,
                int64_t max_points, double *x, int32_t *n_iters, int64_t *points_done, int32_t *status, int mem)
{
    int rc = check_ready(c);
    if (rc) return rc;
    if (!start || !stop || !step || !x || !n_iters || !points_done || !status)
        return fail(c, FTSB_E_ARG, "ftsb_run_dc: null argument");
    if (sweep_elem < 0 || sweep_elem >= c->comp.n_elems || c->comp.src_of_elem[sweep_elem] < 0)
        return fail(c, FTSB_E_ARG, "ftsb_run_dc: sweep_elem must be a V or I source (src/spice.pest:6)");
    if (max_points <= 0) return fail(c, FTSB_E_ARG, "ftsb_run_dc: max_points must be > 0");
    const long long B = c->batch;
    const int n = c->comp.n_run;
    RunArgs a{};
    a.batch = B;
    a.vals = (const double *)c->vals.p;
    a.fn = (const double *)c->fn.p;
    a.sweep_src = c->comp.src_of_elem[sweep_elem];
    a.max_points = max_points;
    if (mem == FTSB_MEM_HOST) {
        for (long long i = 0; i < B; ++i) {
            double cnt = std::ceil((stop[i] - start[i]) / step[i]);
            if (cnt > (double)max_points)
                return fail(c, FTSB_E_ARG, "ftsb_run_dc: instance " + std::to_string(i) + " has more than max_points points");
        }
        FTSB_CUDA(c, c->i_start.ensure((size_t)B * 8));
        FTSB_CUDA(c, c->i_stop.ensure((size_t)B * 8));
        FTSB_CUDA(c, c->i_step.ensure((size_t)B * 8));
        FTSB_CUDA(c, cudaMemcpyAsync(c->i_start.p, start, (size_t)B * 8, cudaMemcpyHostToDevice, c->stream));
        FTSB_CUDA(c, cudaMemcpyAsync(c->i_stop.p, stop, (size_t)B * 8, cudaMemcpyHostToDevice, c->stream));
        FTSB_CUDA(c, cudaMemcpyAsync(c->i_step.p, step, (size_t)B * 8, cudaMemcpyHostToDevice, c->stream));
        FTSB_CUDA(c, c->o_x.ensure((size_t)B * max_points * std::max(n, 1) * 8));
        FTSB_CUDA(c, c->o_it.ensure((size_t)B * max_points * 4));
        FTSB_CUDA(c, c->o_done.ensure((size_t)B * 8));
        FTSB_CUDA(c, c->o_status.ensure((size_t)B * 4));
        a.start = (const double *)c->i_start.p;
        a.stop = (const double *)c->i_stop.p;
        a.step = (const double *)c->i_step.p;
        a.x = (double *)c->o_x.p;
        a.n_iters = (int *)c->o_it.p;
        a.points_done = (long long *)c->o_done.p;
        a.status = (int *)c->o_status.p;
        FTSB_CUDA(c, cudaMemsetAsync(a.n_iters, 0, (size_t)B * max_points * 4, c->stream));
        FTSB_CUDA(c, cudaMemsetAsync(a.x, 0, (size_t)B * max_points * n * 8, c->stream));
    } else {
        a.start = start;
        a.stop = stop;
        a.step = step;
        a.x = x;
        a.n_iters = n_iters;
        a.points_done = (long long *)points_done;
        a.status = status;
    }
    rc = launch(c, AN_DC, a);
    if (rc) return rc;
    if (mem == FTSB_MEM_HOST) {
        FTSB_CUDA(c, cudaMemcpyAsync(x, a.x, (size_t)B * max_points * n * 8, cudaMemcpyDeviceToHost, c->stream));
        FTSB_CUDA(c, cudaMemcpyAsync(n_iters, a.n_iters, (size_t)B * max_points * 4, cudaMemcpyDeviceToHost, c->stream));
        FTSB_CUDA(c, cudaMemcpyAsync(points_done, a.points_done, (size_t)B * 8, cudaMemcpyDeviceToHost, c->stream));
        FTSB_CUDA(c, cudaMemcpyAsync(status, a.status, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream));
        FTSB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// ---- pipelined host-buffer entry points ------------------------------------------------------------------
// ftsb_set_params + ftsb_run_op / ftsb_run_dc in one blocking call.  When the handle is in its steady state on the
// per-topology kernel (plans learned, kernel loaded, nothing left to relearn) the batch is cut into chunks, each on
// its own stream: host->device copy of the chunk's parameters, the kernel over the chunk's instance range (global
// indices, one shared redo list), device->host copy of the chunk's results.  Chunk kernels co-reside, so the SMs stay
// as full as with one launch while the copies of the neighbouring chunks run on the copy engines.  Instances the
// kernel abandoned are redone by the dense kernel afterwards and the outputs copied again (rare).  In every other
// state (first calls of a handle, other kernel families, parameter-major input) the two serial calls run instead.
} // extern "C"

namespace {

bool jit_steady(const ftsb_circuit *c, int an, long long batch)
{
    return (an == AN_OP || an == AN_DC) && c->plan_state[an] == 1 && c->jit_state[an] == 1 && c->relearn_left[an] == 0 &&
           c->jit_opt && c->skip_refactor && c->plan_opt && !c->force_generic && !c->force_cta && !c->sparse && c->lanes <= 0 &&
           std::max(c->comp.n_start, 1) <= 16 && batch >= 4096 && c->pipe_chunks > 1;
}

int pipe_init(ftsb_circuit *c)
{
    if (c->pipe_ready) return 0;
    for (int k = 0; k < ftsb_circuit::PIPE_MAX; ++k) {
        FTSB_CUDA(c, cudaStreamCreateWithFlags(&c->pipe[k], cudaStreamNonBlocking));
        FTSB_CUDA(c, cudaEventCreateWithFlags(&c->pipe_ev[k], cudaEventDisableTiming));
    }
    FTSB_CUDA(c, cudaEventCreateWithFlags(&c->pipe_ev0, cudaEventDisableTiming));
    FTSB_CUDA(c, c->pipe_work.ensure(ftsb_circuit::PASS_MAX * ftsb_circuit::PIPE_MAX * 8));
    c->pipe_ready = true;
    return 0;
}

struct PipeIo { // host buffers of one call
    const double *vals, *fn, *start, *stop, *step;
    double *x;
    int32_t *n_iters, *status;
    int64_t *points_done;
    long long P; // points per instance (.op: 1)
};

// an error path of the pipelined call must not return while chunk streams still copy into the caller's buffers
void pipe_drain(ftsb_circuit *c)
{
    const std::string keep = c->err;
    if (c->pipe_ready)
        for (int k = 0; k < ftsb_circuit::PIPE_MAX; ++k) cudaStreamSynchronize(c->pipe[k]);
    cudaStreamSynchronize(c->stream);
    cudaGetLastError();
    c->err = keep;
}

int solve_pipelined_impl(ftsb_circuit *c, int an, long long B, const PipeIo &io, int sweep_src);

int solve_pipelined(ftsb_circuit *c, int an, long long B, const PipeIo &io, int sweep_src)
{
    const int rc = solve_pipelined_impl(c, an, B, io, sweep_src);
    if (rc) pipe_drain(c);
    return rc;
}

int solve_pipelined_impl(ftsb_circuit *c, int an, long long B, const PipeIo &io, int sweep_src)
{
    int rc = pipe_init(c);
    if (rc) return rc;
    std::unique_lock<std::mutex> device_lock(device_launch_mutex(c->device));
    const int ne = c->comp.n_elems, nf = c->comp.n_fn;
    const int n = an == AN_OP ? c->comp.n_start : c->comp.n_run;
    const long long P = io.P;
    FTSB_CUDA(c, c->vals.ensure((size_t)B * std::max(ne, 1) * 8));
    FTSB_CUDA(c, c->fn.ensure((size_t)B * std::max(nf, 1) * 7 * 8));
    FTSB_CUDA(c, c->o_x.ensure((size_t)B * P * std::max(n, 1) * 8));
    FTSB_CUDA(c, c->o_it.ensure((size_t)B * P * 4));
    FTSB_CUDA(c, c->o_status.ensure((size_t)B * 4));
    FTSB_CUDA(c, c->redo.ensure((size_t)B * 8));
    if (an == AN_DC) {
        FTSB_CUDA(c, c->i_start.ensure((size_t)B * 8));
        FTSB_CUDA(c, c->i_stop.ensure((size_t)B * 8));
        FTSB_CUDA(c, c->i_step.ensure((size_t)B * 8));
        FTSB_CUDA(c, c->o_done.ensure((size_t)B * 8));
    }
    c->batch = B;
    c->vs_inst = ne;
    c->vs_elem = 1;
    c->fs_inst = (long long)nf * 7;
    c->fs_elem = 1;
    RunArgs a{};
    a.prog = c->prog;
    a.vs_inst = c->vs_inst;
    a.vs_elem = c->vs_elem;
    a.fs_inst = c->fs_inst;
    a.fs_elem = c->fs_elem;
    a.batch = B;
    a.vals = (const double *)c->vals.p;
    a.fn = (const double *)c->fn.p;
    a.work = (unsigned long long *)c->work.p;
    a.counters = (unsigned long long *)c->counters.p;
    a.skip_refactor = c->skip_refactor;
    a.x = (double *)c->o_x.p;
    a.n_iters = (int *)c->o_it.p;
    a.status = (int *)c->o_status.p;
    a.sweep_src = sweep_src;
    a.max_points = P;
    a.start = (const double *)c->i_start.p;
    a.stop = (const double *)c->i_stop.p;
    a.step = (const double *)c->i_step.p;
    a.points_done = (long long *)c->o_done.p;
    a.redo_list = (long long *)c->redo.p;
    a.redo_count = (unsigned long long *)c->counters.p + CN_REDO;
    // the dense kernel that takes the redo list
    LaunchPlan pd;
    rc = plan_launch(c, an, B, pd);
    if (rc) return rc;
    const std::string dv = c->variant;
    if (pd.scratch_bytes) FTSB_CUDA(c, c->scratch.ensure(pd.scratch_bytes));
    a.scratch = (double *)c->scratch.p;

    // chunks: equal, multiples of 32 instances
    int K = c->pipe_chunks;
    long long chunk = ((B + K - 1) / K + 31) / 32 * 32;
    chunk = std::max(chunk, 2048LL);
    K = (int)((B + chunk - 1) / chunk);
    for (int j = 0; j < ftsb_circuit::PASS_MAX; ++j)
        for (int k = 0; k < K; ++k) c->pipe_begin[j * ftsb_circuit::PIPE_MAX + k] = (unsigned long long)(k * chunk);
    FTSB_CUDA(c, cudaMemsetAsync(c->counters.p, 0, CN_COUNT * 8, c->stream));
    FTSB_CUDA(c, cudaMemcpyAsync(c->pipe_work.p, c->pipe_begin, sizeof c->pipe_begin, cudaMemcpyHostToDevice, c->stream));
    FTSB_CUDA(c, cudaEventRecord(c->pipe_ev0, c->stream));
    auto d2h = [&](cudaStream_t s, long long b0, long long cnt) -> int {
        FTSB_CUDA(c, cudaMemcpyAsync(io.x + b0 * P * n, a.x + b0 * P * n, (size_t)cnt * P * n * 8, cudaMemcpyDeviceToHost, s));
        FTSB_CUDA(c, cudaMemcpyAsync(io.n_iters + b0 * P, a.n_iters + b0 * P, (size_t)cnt * P * 4, cudaMemcpyDeviceToHost, s));
        FTSB_CUDA(c, cudaMemcpyAsync(io.status + b0, a.status + b0, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s));
        if (an == AN_DC)
            FTSB_CUDA(c, cudaMemcpyAsync(io.points_done + b0, a.points_done + b0, (size_t)cnt * 8, cudaMemcpyDeviceToHost, s));
        return 0;
    };
    for (int k = 0; k < K; ++k) {
        cudaStream_t s = c->pipe[k];
        const long long b0 = k * chunk, b1 = std::min(B, b0 + chunk), cnt = b1 - b0;
        FTSB_CUDA(c, cudaStreamWaitEvent(s, c->pipe_ev0, 0));
        if (ne > 0)
            FTSB_CUDA(c, cudaMemcpyAsync((double *)c->vals.p + b0 * ne, io.vals + b0 * ne, (size_t)cnt * ne * 8, cudaMemcpyHostToDevice, s));
        if (nf > 0)
            FTSB_CUDA(c, cudaMemcpyAsync((double *)c->fn.p + b0 * nf * 7, io.fn + b0 * nf * 7, (size_t)cnt * nf * 56, cudaMemcpyHostToDevice, s));
        if (an == AN_DC) {
            FTSB_CUDA(c, cudaMemcpyAsync((double *)c->i_start.p + b0, io.start + b0, (size_t)cnt * 8, cudaMemcpyHostToDevice, s));
            FTSB_CUDA(c, cudaMemcpyAsync((double *)c->i_stop.p + b0, io.stop + b0, (size_t)cnt * 8, cudaMemcpyHostToDevice, s));
            FTSB_CUDA(c, cudaMemcpyAsync((double *)c->i_step.p + b0, io.step + b0, (size_t)cnt * 8, cudaMemcpyHostToDevice, s));
            FTSB_CUDA(c, cudaMemsetAsync(a.n_iters + b0 * P, 0, (size_t)cnt * P * 4, s));
            FTSB_CUDA(c, cudaMemsetAsync(a.x + b0 * P * n, 0, (size_t)cnt * P * n * 8, s));
        }
        if (jit_two_pass_applies(c, an)) {
            for (int j = 0; j < c->jit_passes && !rc; ++j)
                rc = launch_jit_range(c, an, a, s, (unsigned long long *)c->pipe_work.p + j * ftsb_circuit::PIPE_MAX + k, b0, b1,
                                      j + 1 < c->jit_passes ? (j + 1) * c->jit_two_pass : 0, j > 0);
        } else
            rc = launch_jit_range(c, an, a, s, (unsigned long long *)c->pipe_work.p + k, b0, b1);
        if (rc) return rc;
        rc = d2h(s, b0, cnt);
        if (rc) return rc;
        FTSB_CUDA(c, cudaEventRecord(c->pipe_ev[k], s));
    }
    for (int k = 0; k < K; ++k) FTSB_CUDA(c, cudaStreamWaitEvent(c->stream, c->pipe_ev[k], 0));
    // dense redo pass over whatever the chunk kernels abandoned (an empty list almost always)
    RunArgs ar = a;
    ar.in_list = a.redo_list;
    ar.in_count = a.redo_count;
    ar.prog.lay = pd.lay;
    rc = launch_kernel(c, pd, ar);
    if (rc) return rc;
    unsigned long long n_redo = 0;
    FTSB_CUDA(c, cudaMemcpyAsync(&n_redo, a.redo_count, 8, cudaMemcpyDeviceToHost, c->stream));
    device_lock.unlock(); // everything is enqueued: other handles of this device may plan and launch while this one waits
    FTSB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (n_redo) {
        rc = d2h(c->stream, 0, B);
        if (rc) return rc;
        FTSB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    const PlanHost &ph = c->plan_host[an];
    c->last_plan_an = an;
    c->last_redo_an = -1;
    c->variant = "jit" + std::to_string(ph.K) + "n" + std::to_string(ph.n) + "/S" + std::to_string(ph.S) + "/r" +
                 std::to_string(c->jit[an].regs) + "/occ" + std::to_string(c->jit_occ[an]) + "/pipe" + std::to_string(K) + "+redo:" + dv;
    return 0;
}

} // namespace

extern "C" {

int ftsb_solve_op(ftsb_circuit *c, int64_t batch, const double *vals, const double *fn, int layout, double *x,
                  int32_t *n_iters, int32_t *status)
{
    if (!c) return fail(nullptr, FTSB_E_ARG, "null handle");
    if (!x || !n_iters || !status) return fail(c, FTSB_E_ARG, "ftsb_solve_op: null output");
    if (batch > 0 && layout == FTSB_LAYOUT_INSTANCE_MAJOR && (c->comp.n_elems == 0 || vals) && (c->comp.n_fn == 0 || fn) &&
        jit_steady(c, AN_OP, batch)) {
        FTSB_CUDA(c, cudaSetDevice(c->device));
        PipeIo io{vals, fn, nullptr, nullptr, nullptr, x, n_iters, status, nullptr, 1};
        return solve_pipelined(c, AN_OP, batch, io, 0);
    }
    int rc = ftsb_set_params(c, batch, vals, fn, layout, FTSB_MEM_HOST);
    if (rc) return rc;
    return ftsb_run_op(c, x, n_iters, status, FTSB_MEM_HOST);
}

int ftsb_solve_dc(ftsb_circuit *c, int64_t batch, const double *vals, const double *fn, int layout, int32_t sweep_elem,
                  const double *start, const double *stop, const double *step, int64_t max_points, double *x,
                  int32_t *n_iters, int64_t *points_done, int32_t *status)
{
    if (!c) return fail(nullptr, FTSB_E_ARG, "null handle");
    if (batch > 0 && layout == FTSB_LAYOUT_INSTANCE_MAJOR && (c->comp.n_elems == 0 || vals) && (c->comp.n_fn == 0 || fn) &&
        start && stop && step && x && n_iters && points_done && status && max_points > 0 && sweep_elem >= 0 &&
        sweep_elem < c->comp.n_elems && c->comp.src_of_elem[sweep_elem] >= 0 && jit_steady(c, AN_DC, batch)) {
        for (long long i = 0; i < batch; ++i) {
            double cnt = std::ceil((stop[i] - start[i]) / step[i]);
            if (cnt > (double)max_points)
                return fail(c, FTSB_E_ARG, "ftsb_solve_dc: instance " + std::to_string(i) + " has more than max_points points");
        }
        FTSB_CUDA(c, cudaSetDevice(c->device));
        PipeIo io{vals, fn, start, stop, step, x, n_iters, status, points_done, (long long)max_points};
        return solve_pipelined(c, AN_DC, batch, io, c->comp.src_of_elem[sweep_elem]);
    }
    int rc = ftsb_set_params(c, batch, vals, fn, layout, FTSB_MEM_HOST);
    if (rc) return rc;
    return ftsb_run_dc(c, sweep_elem, start, stop, step, max_points, x, n_iters, points_done, status, FTSB_MEM_HOST);
}

int ftsb_run_tran(ftsb_circuit *c, double stop, double step_max, const ftsb_tran_opts *opts, double *t, double *x,
                  int32_t *n_iters, int64_t *n_rec, double *x_final, int32_t *status, int mem)
{
    int rc = check_ready(c);
    if (rc) return rc;
    if (!n_rec || !status) return fail(c, FTSB_E_ARG, "ftsb_run_tran: n_rec and status are required");
    const long long B = c->batch;
    const int n = c->comp.n_run;
    const long long cap = opts ? opts->capacity : 0;
    if (cap < 0) return fail(c, FTSB_E_ARG, "ftsb_run_tran: negative capacity");
    if (cap == 0) t = nullptr, x = nullptr, n_iters = nullptr;
    RunArgs a{};
    a.batch = B;
    a.vals = (const double *)c->vals.p;
    a.fn = (const double *)c->fn.p;
    a.t_stop = stop;
    a.step_max = step_max;
    a.cap = cap;
    a.rec_stride = opts ? opts->rec_stride : 1;
    if (mem == FTSB_MEM_HOST) {
        if (x) {
            FTSB_CUDA(c, c->o_x.ensure((size_t)B * cap * std::max(n, 1) * 8));
            FTSB_CUDA(c, cudaMemsetAsync(c->o_x.p, 0, (size_t)B * cap * n * 8, c->stream));
        }
        if (t) {
            FTSB_CUDA(c, c->o_t.ensure((size_t)B * cap * 8));
            FTSB_CUDA(c, cudaMemsetAsync(c->o_t.p, 0, (size_t)B * cap * 8, c->stream));
        }
        if (n_iters) {
            FTSB_CUDA(c, c->o_it.ensure((size_t)B * cap * 4));
            FTSB_CUDA(c, cudaMemsetAsync(c->o_it.p, 0, (size_t)B * cap * 4, c->stream));
        }
        FTSB_CUDA(c, c->o_nrec.ensure((size_t)B * 8));
        FTSB_CUDA(c, c->o_status.ensure((size_t)B * 4));
        if (x_final) FTSB_CUDA(c, c->o_xf.ensure((size_t)B * std::max(n, 1) * 8));
        a.x = x ? (double *)c->o_x.p : nullptr;
        a.t_out = t ? (double *)c->o_t.p : nullptr;
        a.n_iters = n_iters ? (int *)c->o_it.p : nullptr;
        a.n_rec = (long long *)c->o_nrec.p;
        a.status = (int *)c->o_status.p;
        a.x_final = x_final ? (double *)c->o_xf.p : nullptr;
    } else {
        a.x = x;
        a.t_out = t;
        a.n_iters = n_iters;
        a.n_rec = (long long *)n_rec;
        a.status = status;
        a.x_final = x_final;
    }
    rc = launch(c, AN_TRAN, a);
    if (rc) return rc;
    if (mem == FTSB_MEM_HOST) {
        if (x) FTSB_CUDA(c, cudaMemcpyAsync(x, a.x, (size_t)B * cap * n * 8, cudaMemcpyDeviceToHost, c->stream));
        if (t) FTSB_CUDA(c, cudaMemcpyAsync(t, a.t_out, (size_t)B * cap * 8, cudaMemcpyDeviceToHost, c->stream));
        if (n_iters) FTSB_CUDA(c, cudaMemcpyAsync(n_iters, a.n_iters, (size_t)B * cap * 4, cudaMemcpyDeviceToHost, c->stream));
        FTSB_CUDA(c, cudaMemcpyAsync(n_rec, a.n_rec, (size_t)B * 8, cudaMemcpyDeviceToHost, c->stream));
        FTSB_CUDA(c, cudaMemcpyAsync(status, a.status, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream));
        if (x_final) FTSB_CUDA(c, cudaMemcpyAsync(x_final, a.x_final, (size_t)B * n * 8, cudaMemcpyDeviceToHost, c->stream));
        FTSB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int ftsb_sync(ftsb_circuit *c)
{
    if (!c) return fail(nullptr, FTSB_E_ARG, "null handle");
    FTSB_CUDA(c, cudaSetDevice(c->device));
    FTSB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

void *ftsb_stream(ftsb_circuit *c) { return c ? (void *)c->stream : nullptr; }

int ftsb_set_stream(ftsb_circuit *c, void *cuda_stream)
{
    if (!c) return fail(nullptr, FTSB_E_ARG, "null handle");
    FTSB_CUDA(c, cudaSetDevice(c->device));
    if (c->stream && c->own_stream) {
        FTSB_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaStreamDestroy(c->stream);
    }
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return 0;
}

int ftsb_get_counters(ftsb_circuit *c, ftsb_counters *out)
{
    if (!c || !out) return fail(c, FTSB_E_ARG, "ftsb_get_counters: null argument");
    FTSB_CUDA(c, cudaSetDevice(c->device));
    unsigned long long h[CN_COUNT];
    FTSB_CUDA(c, cudaMemcpyAsync(h, c->counters.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    FTSB_CUDA(c, cudaStreamSynchronize(c->stream));
    memset(out, 0, sizeof *out);
    out->solve_points = (int64_t)h[CN_POINTS];
    out->newton_iters = (int64_t)h[CN_NEWTON];
    out->lu_factor = (int64_t)h[CN_FACTOR];
    out->lu_solve = (int64_t)h[CN_SOLVE];
    out->rej_newton = (int64_t)h[CN_REJ_NEWTON];
    out->rej_plte = (int64_t)h[CN_REJ_PLTE];
    out->redo = (int64_t)h[CN_REDO];
    out->sparse_flops = (int64_t)h[CN_FLOPS];
    if (c->last_plan_an >= 0) { // plan / JIT kernels: flops of the factorisations and substitutions along the main plan
        const PlanHost &ph = c->plan_host[c->last_plan_an];
        out->sparse_flops = out->lu_factor * ph.flops_factor + out->lu_solve * ph.flops_solve;
    }
    return 0;
}

// diagnostics (not part of include/ftsb.h): why the structure-exploiting warp kernel handed instances to the dense kernels in the
// last launch: out[0] pivot column with no usable candidate (zero-only / non-finite), out[1] fill-in pool exhausted, out[2]
// non-finite solution component
extern "C" int ftsb_debug_why(ftsb_circuit *c, int64_t *out)
{
    if (!c || !out) return -1;
    if (cudaSetDevice(c->device) != cudaSuccess) return -2;
    unsigned long long h[CN_COUNT];
    if (cudaMemcpyAsync(h, c->counters.p, sizeof h, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess)
        return -2;
    for (int k = 0; k < 3; ++k) out[k] = (int64_t)((h[CN_WHY] >> (20 * k)) & 0xfffffull);
    return 0;
}

int64_t ftsb_launch_count(const ftsb_circuit *c) { return c ? c->launches : 0; }
const char *ftsb_last_variant(const ftsb_circuit *c) { return c ? c->variant.c_str() : ""; }

} // extern "C"

// ---------------------------------------------------------------------------------------------------------
// standalone batched LU solve (KATs of gauss_lu.rs:60-109) and the fp64 micro-benchmark
// ---------------------------------------------------------------------------------------------------------
namespace {

template <class Team>
__device__ void lu_one(const Team &tm, int n, const double *a, const double *b, double *x, double *J, double *y,
                       double *xs, int *piv)
{
    for (int c = 0; c < n; ++c)
        for (int i = tm.rank; i < n; i += tm.size()) J[c * n + i] = a[i * n + c]; // row-major in, column-major J
    for (int i = tm.rank; i < n; i += tm.size()) y[i] = b[i];
    tm.sync();
    int mypos = 0;
    lu_factor(tm, n, n, J, piv, mypos);
    lu_solve(tm, n, n, J, piv, mypos, y, xs);
    for (int i = tm.rank; i < n; i += tm.size()) x[i] = xs[i];
    tm.sync();
}

__global__ void __launch_bounds__(32) k_lu_warp(int n, long long batch, const double *a, const double *b, double *x)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *J = reinterpret_cast<double *>(smem_raw);
    double *y = J + n * n;
    double *xs = y + n;
    int *piv = reinterpret_cast<int *>(xs + n);
    SubWarp<32> tm;
    for (long long s = blockIdx.x; s < batch; s += gridDim.x)
        lu_one(tm, n, a + s * n * n, b + s * n, x + s * n, J, y, xs, piv);
}

__global__ void k_lu_cta(int n, long long batch, const double *a, const double *b, double *x, double *scratch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *red = reinterpret_cast<double *>(smem_raw);
    double *y = red + 96;
    double *xs = y + n;
    int *piv = reinterpret_cast<int *>(xs + n);
    double *J = scratch + (size_t)blockIdx.x * n * n;
    CtaTeam tm(red);
    for (long long s = blockIdx.x; s < batch; s += gridDim.x)
        lu_one(tm, n, a + s * n * n, b + s * n, x + s * n, J, y, xs, piv);
}

__global__ void k_fp64_fma(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0,
           a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c);
        a1 = fma(a1, m, c);
        a2 = fma(a2, m, c);
        a3 = fma(a3, m, c);
        a4 = fma(a4, m, c);
        a5 = fma(a5, m, c);
        a6 = fma(a6, m, c);
        a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

} // namespace

extern "C" int ftsb_lu_solve(int device, int32_t n, int64_t batch, const double *a, const double *b, double *x, int mem)
{
    if (n <= 0 || batch <= 0 || !a || !b || !x) return fail(nullptr, FTSB_E_ARG, "ftsb_lu_solve: bad argument");
    if (n > 1024) return fail(nullptr, FTSB_E_UNSUPPORTED, "ftsb_lu_solve: n > 1024");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(nullptr, FTSB_E_CUDA, "ftsb_lu_solve: no usable CUDA device; this library has no CPU fallback");
    }
#define LU_CUDA(call)                                                                               \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            g_last_error = std::string(#call) + ": " + cudaGetErrorString(e__);                     \
            return FTSB_E_CUDA;                                                                     \
        }                                                                                           \
    } while (0)
    LU_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LU_CUDA(cudaGetDeviceProperties(&prop, device));
    const size_t ab = (size_t)batch * n * n * 8, vb = (size_t)batch * n * 8;
    DevBuf da, db, dx, scr;
    const double *pa = a, *pb = b;
    double *px = x;
    if (mem == FTSB_MEM_HOST) {
        LU_CUDA(da.ensure(ab));
        LU_CUDA(db.ensure(vb));
        LU_CUDA(dx.ensure(vb));
        LU_CUDA(cudaMemcpy(da.p, a, ab, cudaMemcpyHostToDevice));
        LU_CUDA(cudaMemcpy(db.p, b, vb, cudaMemcpyHostToDevice));
        pa = (const double *)da.p;
        pb = (const double *)db.p;
        px = (double *)dx.p;
    }
    int grid = (int)std::min<long long>(batch, (long long)prop.multiProcessorCount * 8);
    if (n <= 32) {
        size_t smem = ((size_t)n * n + 2 * n) * 8 + (size_t)n * 4 + 16;
        k_lu_warp<<<grid, 32, smem>>>(n, batch, pa, pb, px);
    } else {
        int block = ((n + 31) / 32) * 32;
        grid = (int)std::min<long long>(batch, (long long)prop.multiProcessorCount * 2);
        LU_CUDA(scr.ensure((size_t)grid * n * n * 8));
        size_t smem = (96 + 2 * (size_t)n) * 8 + (2 * (((size_t)n + 3) & ~(size_t)3) + 8) * 4 + 16; // piv + column list
        k_lu_cta<<<grid, block, smem>>>(n, batch, pa, pb, px, (double *)scr.p);
    }
    LU_CUDA(cudaGetLastError());
    LU_CUDA(cudaDeviceSynchronize());
    if (mem == FTSB_MEM_HOST) LU_CUDA(cudaMemcpy(x, px, vb, cudaMemcpyDeviceToHost));
    da.release();
    db.release();
    dx.release();
    scr.release();
    return 0;
#undef LU_CUDA
}

extern "C" double ftsb_measure_fp64_tflops(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return -1.0;
    }
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    const int block = 256, grid = prop.multiProcessorCount * 8, iters = 1 << 16;
    double *out = nullptr;
    if (cudaMalloc(&out, (size_t)grid * block * 8) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_fp64_fma<<<grid, block>>>(out, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8.0 * (double)iters * (double)grid * block;
        if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}


// ---------------------------------------------------------------------------------------------------------
// diagnostics (not part of include/ftsb.h): the damping's ln_1p (fts_log1p.h) evaluated on the device and on the host
// ---------------------------------------------------------------------------------------------------------
namespace {
__global__ void k_debug_log1p(const double *in, double *out, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = damp_log1p(in[i]);
}
} // namespace

extern "C" int ftsb_debug_log1p(int device, int64_t n, const double *in, double *out)
{
    if (device < 0) { // host evaluation of the same header
        alignas(16) static const unsigned long long tab[2 * FTS_LOG1P_N] = {FTS_LOG1P_TABLE};
        for (int64_t i = 0; i < n; ++i) out[i] = fts_log1p_pos(in[i], reinterpret_cast<const double *>(tab));
        return 0;
    }
    double *d = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&d, (size_t)n * 16) != cudaSuccess) return FTSB_E_CUDA;
    cudaMemcpy(d, in, (size_t)n * 8, cudaMemcpyHostToDevice);
    k_debug_log1p<<<592, 256>>>(d, d + n, (long long)n);
    cudaError_t e = cudaMemcpy(out, d + n, (size_t)n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? 0 : FTSB_E_CUDA;
}

// ---------------------------------------------------------------------------------------------------------
// diagnostics (not part of include/ftsb.h): generate + compile the per-topology kernel without a GPU
// ---------------------------------------------------------------------------------------------------------
extern "C" int ftsb_debug_jit(const ftsb_elem *elems, int32_t n_elems, int32_t n_run, int32_t n_start, const int32_t *row_class,
                              int32_t an, const uint64_t *perms, const int64_t *counts, int32_t n_perms, const char *src_path,
                              char *log, int32_t log_cap)
{
    Compiled comp;
    std::string err;
    auto say = [&](const std::string &m) {
        if (log && log_cap > 0) snprintf(log, (size_t)log_cap, "%s", m.c_str());
    };
    if (!compile_circuit(reinterpret_cast<const ElemDesc *>(elems), n_elems, n_run, n_start, row_class, 0, comp, err)) {
        say(err);
        return -1;
    }
    const ModeOff &mo = comp.mode[an];
    const std::vector<int> &ib = comp.ints();
    std::vector<int> jdst(ib.begin() + mo.J_full.dst, ib.begin() + mo.J_full.dst + mo.J_full.n_dst);
    std::vector<std::pair<uint64_t, long long>> pv;
    for (int i = 0; i < n_perms; ++i) pv.push_back({perms[i], (long long)counts[i]});
    PlanHost ph;
    if (!build_plans(mo.n, jdst, pv, ph)) {
        say("build_plans failed");
        return -2;
    }
    const std::string src = jit_source(comp, an, ph);
    if (src_path) {
        if (FILE *f = fopen(src_path, "w")) {
            fwrite(src.data(), 1, src.size(), f);
            fclose(f);
        }
    }
    std::string cubin;
    if (!jit_compile(src, cubin, err)) {
        say(err);
        return -3;
    }
    if (src_path) {
        if (FILE *f = fopen((std::string(src_path) + ".cubin").c_str(), "wb")) {
            fwrite(cubin.data(), 1, cubin.size(), f);
            fclose(f);
        }
    }
    say("ok K=" + std::to_string(ph.K) + " S=" + std::to_string(ph.S) + " cubin " + std::to_string(cubin.size()) + " bytes; " + err);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// diagnostics (not part of include/ftsb.h): the static elimination plan of the per-team kernels, built without a GPU.
// meta[0..6] = n, nA, S, nJ, nT, nU, nB; meta[7..23] = the 17 list offsets in TeamPlanHost order; meta[24] = ints used;
// meta[25..40] = ellW, nSlotA, chain, nFop, nBop, nFchunk, nBchunk and the offsets of fop, fvidx, fchunk, bop, bvidx, bchunk,
// ellcol, ovfptr, ovfcol.
// jdst[nJ] = row | col << 16 of the static entries, piv[n] = the pivot sequence the plan follows.
// ---------------------------------------------------------------------------------------------------------
extern "C" int ftsb_debug_team_plan(const ftsb_elem *elems, int32_t n_elems, int32_t n_run, int32_t n_start, const int32_t *row_class,
                                    int32_t an, int32_t *ibuf, int32_t cap, int32_t *meta, int32_t *jdst, int32_t *piv)
{
    Compiled comp;
    std::string err;
    if (!compile_circuit(reinterpret_cast<const ElemDesc *>(elems), n_elems, n_run, n_start, row_class, 0, comp, err)) return -1;
    const ModeOff &mo = comp.mode[an];
    if (!mo.sp.piv_valid) return -2;
    const std::vector<int> &ib = comp.ints();
    std::vector<int> pv(ib.begin() + mo.sp.piv, ib.begin() + mo.sp.piv + mo.n);
    TeamPlanHost ph;
    if (!build_team_plan(comp, an, pv, ph)) return -3;
    if ((int)ph.ibuf.size() > cap) return -4;
    std::copy(ph.ibuf.begin(), ph.ibuf.end(), ibuf);
    const int m[] = {ph.n, ph.nA, ph.S, ph.nJ, ph.nT, ph.nU, ph.nB, ph.o_arow, ph.o_acol, ph.o_aptr, ph.o_acon, ph.o_sptr, ph.o_scon,
                     ph.o_pivrow, ph.o_pive, ph.o_tptr, ph.o_tgt, ph.o_tgtr, ph.o_uptr, ph.o_updm, ph.o_updds, ph.o_bptr, ph.o_bwde,
                     ph.o_bwdr, (int)ph.ibuf.size()};
    std::copy(m, m + 25, meta);
    const int m2[] = {ph.ellW, ph.nSlotA, ph.chain, ph.nFop, ph.nBop, ph.nFchunk, ph.nBchunk, ph.o_fop, ph.o_fvidx, ph.o_fchunk,
                      ph.o_bop, ph.o_bvidx, ph.o_bchunk, ph.o_ellcol, ph.o_ovfptr, ph.o_ovfcol};
    std::copy(m2, m2 + 16, meta + 25); // meta[25..40]
    const int m3[] = {ph.nEarly, ph.nXop, ph.nDFop, ph.nDBop, ph.o_early, ph.o_xop, ph.o_dfop, ph.o_dbop};
    std::copy(m3, m3 + 8, meta + 41); // meta[41..48]: the direct streams
    const int m4[] = {ph.nXlev, ph.nFlev, ph.nBlev, ph.o_xlptr, ph.o_xlop, ph.o_flptr, ph.o_flop, ph.o_blptr, ph.o_blop};
    std::copy(m4, m4 + 9, meta + 49); // meta[49..57]: the level schedules
    meta[58] = ph.o_spre;
    const int m5[] = {ph.nXseg, ph.o_xseg, ph.o_xsop, ph.nBseg, ph.o_bseg, ph.o_bsop};
    std::copy(m5, m5 + 6, meta + 59); // meta[59..64]: the segmented form of the level schedules (the caller's meta holds >= 96 ints)
    for (int d = 0; d < ph.nJ; ++d) jdst[d] = ib[mo.J_full.dst + d];
    std::copy(pv.begin(), pv.end(), piv);
    return 0;
}
