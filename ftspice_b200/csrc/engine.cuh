// engine.cuh — device engine: one *team* of threads (a sub-warp group of 4/8/16/32 lanes, or a whole CTA)
// owns one circuit instance for the whole analysis; the Newton loop, the sweep chain and the time loop stay
// resident on the SM.  Replaces, for a batch of instances (all citations relative to /root/reference/):
//   MNA::get_err                       src/engine/mna.rs:25-29
//   newtons_method::solve / dampen_step / converged     src/engine/newtons_method.rs:19-89
//   gauss_lu::solve                    src/engine/gauss_lu.rs:3-54
//   NodeVecNorm::new                   src/engine/node_vec_norm.rs:12-33
//   transient::step, plte_is_too_big, plte_can_grow      src/engine/transient.rs:18-105
//   StateHistory::plte / divided_diff  src/engine/transient/state_history.rs:40-56
//   SpiceFn::eval                      src/spice_fn.rs:8-124
//   device models and stamps           src/device/*.rs, src/device/*/model.rs
//   Engine::run_op / run_dc / run_tran src/engine.rs:62-206
#pragma once
#include <cuda_runtime.h>

#include <climits>

#include "plan.h"
#include "prog.h"

#include "log1p_tab.h"
#include "fts_log1p.h"

namespace ftsb {

// ln_1p of the damped Newton step (newtons_method.rs:73-77) for its non-negative argument: fts_log1p.h
static __device__ __align__(16) const unsigned long long k_log1p_tab[2 * FTS_LOG1P_N] = {FTS_LOG1P_TABLE};
__device__ __forceinline__ double damp_log1p(double u) { return fts_log1p_pos(u, reinterpret_cast<const double *>(k_log1p_tab)); }


// ---------------------------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------------------------
// CN_FLOPS: floating-point operations the structure-exploiting solver executed in factorisations and substitutions
// (the dense kernels' flops follow from CN_FACTOR / CN_SOLVE and n)
enum : int { CN_POINTS = 0, CN_NEWTON = 1, CN_FACTOR = 2, CN_SOLVE = 3, CN_REJ_NEWTON = 4, CN_REJ_PLTE = 5, CN_FLOPS = 6, CN_LOCAL = 7,
             CN_REDO = 7, CN_WHY = 8, CN_REDO2 = 9, CN_COUNT = 12 };

struct RunArgs {
    DevProg prog;
    long long batch;
    const double *vals; // element e of instance i at vals[i * vs_inst + e * vs_elem]: [batch][n_elems] or [n_elems][batch]
    const double *fn;   // word k (= fn_slot * 7 + field) of instance i at fn[i * fs_inst + k * fs_elem]
    long long vs_inst, vs_elem, fs_inst, fs_elem; // no transposition on the way in: the kernels read either layout in place
    unsigned long long *work; // atomic work counter
    unsigned long long *counters; // [CN_COUNT]
    double *scratch;    // global A/J storage when the matrices do not fit in shared memory
    int skip_refactor;
    int pad0;
    // outputs common
    double *x;
    int *n_iters;
    int *status;
    // dc
    int sweep_src; // index into prog.src
    int pad1;
    const double *start, *stop, *step;
    long long max_points;
    long long *points_done;
    // tran
    double t_stop, step_max;
    long long cap;
    int rec_stride;
    int pad2;
    double *t_out;
    long long *n_rec;
    double *x_final;
    // .tran of the small-circuit kernels: start-up solution of a previous .op launch
    const double *x_start;   // [batch][n_start]
    const int *start_status; // [batch]
    int n_start;
    int pad3;
    // sparse kernels: instances they abandon (zero / non-finite pivot, fill-in pool exhausted) ...
    long long *redo_list;
    unsigned long long *redo_count;
    // ... and the dense kernels' list mode: run only the instances in_list[0 .. *in_count)
    const long long *in_list;
    const unsigned long long *in_count;
    // learned elimination plans (solver_plan.cuh) and, in the pilot run that learns them, the pivot-sequence trace
    PlanProg plan;
    unsigned long long *trace; // [trace_rows][PLAN_TRACE_MAX] nibble-packed position -> row of every factorisation (ring)
    long long trace_rows;      // rows available: one per work-list position
};

// next work item of a persistent kernel: an instance index, or -1 when the work is exhausted
__device__ __forceinline__ long long next_instance(const RunArgs &a)
{
    const unsigned long long idx = atomicAdd(a.work, 1ULL);
    if (a.in_list) return idx < *a.in_count ? a.in_list[idx] : -1;
    return idx < (unsigned long long)a.batch ? (long long)idx : -1;
}

// ---------------------------------------------------------------------------------------------------------
// teams
// ---------------------------------------------------------------------------------------------------------
template <int G>
struct SubWarp {
    static constexpr bool kIsCta = false;
    int rank;
    unsigned mask;
    __device__ SubWarp()
    {
        int lane = threadIdx.x & 31;
        rank = lane % G;
        mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - rank));
    }
    __device__ int size() const { return G; }
    __device__ void sync() const { __syncwarp(mask); }
    __device__ double sum(double v) const
    {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
        return v;
    }
    __device__ void sum2(double &a, double &b) const
    {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            a += __shfl_xor_sync(mask, a, o);
            b += __shfl_xor_sync(mask, b, o);
        }
    }
    __device__ int sumi(int v) const
    {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
        return v;
    }
    // (largest v, then smallest pos) over the team; every lane gets the same answer (total order, no NaN)
    __device__ void argmax(double &v, int &pos) const
    {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            double v2 = __shfl_xor_sync(mask, v, o);
            int p2 = __shfl_xor_sync(mask, pos, o);
            if (v2 > v || (v2 == v && p2 < pos)) {
                v = v2;
                pos = p2;
            }
        }
    }
    __device__ long long bcast_ll(long long v) const
    {
        return __shfl_sync(mask, v, (threadIdx.x & 31) - rank);
    }
};

struct CtaTeam {
    static constexpr bool kIsCta = true;
    int rank;
    double *red; // shared scratch: [64] doubles + [32] ints
    __device__ explicit CtaTeam(double *scratch) : rank(threadIdx.x), red(scratch) {}
    __device__ int size() const { return blockDim.x; }
    __device__ void sync() const { __syncthreads(); }
    __device__ void sum2(double &a, double &b) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) {
            red[2 * w] = a;
            red[2 * w + 1] = b;
        }
        __syncthreads();
        a = 0.0;
        b = 0.0;
        for (int i = 0; i < nw; ++i) {
            a += red[2 * i];
            b += red[2 * i + 1];
        }
    }
    __device__ double sum(double v) const
    {
        double z = 0.0;
        sum2(v, z);
        return v;
    }
    __device__ int sumi(int v) const
    {
        double a = (double)v, z = 0.0;
        sum2(a, z);
        return (int)a;
    }
    __device__ void argmax(double &v, int &pos) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double v2 = __shfl_xor_sync(0xffffffffu, v, o);
            int p2 = __shfl_xor_sync(0xffffffffu, pos, o);
            if (v2 > v || (v2 == v && p2 < pos)) {
                v = v2;
                pos = p2;
            }
        }
        int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        int *redi = reinterpret_cast<int *>(red + 64);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) {
            red[w] = v;
            redi[w] = pos;
        }
        __syncthreads();
        v = red[0];
        pos = redi[0];
        for (int i = 1; i < nw; ++i) {
            double v2 = red[i];
            int p2 = redi[i];
            if (v2 > v || (v2 == v && p2 < pos)) {
                v = v2;
                pos = p2;
            }
        }
    }
    __device__ long long bcast_ll(long long v) const
    {
        long long *r = reinterpret_cast<long long *>(red + 32);
        __syncthreads();
        if (threadIdx.x == 0) *r = v;
        __syncthreads();
        return *r;
    }
};

// per-instance working set.  Vec<S> is a strided view: S = 1 for team kernels (an instance's data is contiguous),
// S = 32 for the thread-per-instance kernels (element e of lane t lives at word e*32 + t: every warp access is
// bank-conflict free whatever per-thread element each lane touches).
template <int S>
struct Vec {
    double *p;
    __device__ __forceinline__ double &operator[](int i) const { return p[i * S]; }
    __device__ __forceinline__ Vec operator+(int off) const { return Vec{p + off * S}; }
    __device__ __forceinline__ explicit operator bool() const { return p != nullptr; }
};
template <int S>
struct IVec {
    int *p;
    __device__ __forceinline__ int &operator[](int i) const { return p[i * S]; }
};
template <int S>
struct WsT {
    Vec<S> A, J, b, r, x, xn, xp, val, state, ring, prev;
    IVec<S> piv;
    int ld;
    const int *aux; // zero-skipping lane kernels: the static structure words (shared by the block)
};
using Ws = WsT<1>;

struct Counters {
    unsigned long long c[CN_LOCAL];
    __device__ Counters()
    {
        for (int i = 0; i < CN_LOCAL; ++i) c[i] = 0;
    }
};

// ---------------------------------------------------------------------------------------------------------
// gather-based assembly (no atomics: one thread owns one destination entry)
// ---------------------------------------------------------------------------------------------------------
// ld > 0: the destinations are matrix entries (row | col << 16) of a column-major matrix with that leading dimension
template <class Team, class V>
__device__ __forceinline__ void gather(const Team &tm, const Gather &g, V out, V base, V val, int ld = 0)
{
    for (int d = tm.rank; d < g.n_dst; d += tm.size()) {
        int off = __ldg(&g.dst[d]);
        if (ld) off = (off >> 16) * ld + (off & 0xffff);
        double acc = base ? base[off] : 0.0;
        int c0 = __ldg(&g.ptr[d]), c1 = __ldg(&g.ptr[d + 1]);
        for (int c = c0; c < c1; ++c) {
            unsigned w = (unsigned)__ldg(&g.con[c]);
            double v = val[w & ~CON_NEG];
            acc += (w & CON_NEG) ? -v : v;
        }
        out[off] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------
// source waveforms  (src/spice_fn.rs:8-124)
// ---------------------------------------------------------------------------------------------------------
// strided read-only view of one instance's parameters (instance-major: stride 1; parameter-major: stride = batch)
struct ParamView {
    const double *p;
    long long se;
    __device__ __forceinline__ double operator[](int e) const { return p[e * se]; }
    __device__ __forceinline__ ParamView operator+(int off) const { return ParamView{p + off * se, se}; }
};
__device__ __forceinline__ ParamView vals_of(const RunArgs &a, long long inst) { return ParamView{a.vals + inst * a.vs_inst, a.vs_elem}; }
__device__ __forceinline__ ParamView fns_of(const RunArgs &a, long long inst) { return ParamView{a.fn + inst * a.fs_inst, a.fs_elem}; }

template <class PV>
__device__ __forceinline__ double spice_fn_eval(int kind, PV p, double t)
{
    if (kind == FN_SIN) {
        return p[0] + p[1] * sin(2.0 * 3.14159265358979323846 * p[2] * t);
    } else if (kind == FN_PULSE) {
        double v1 = p[0], v2 = p[1], delay = p[2], t_rise = p[3], t_fall = p[4], pw = p[5], period = p[6];
        double tn = fabs(fmod(t, period));
        if (tn < delay) return v1;
        if (tn < delay + t_rise) return v1 + (tn - delay) / t_rise * (v2 - v1);
        if (tn < delay + pw - t_fall) return v2;
        if (tn < delay + pw) return v2 + (tn - (delay + pw - t_fall)) / t_fall * (v1 - v2);
        return v1;
    } else { // FN_EXP
        double v1 = p[0], v2 = p[1], rd = p[2], rtau = p[3], fd = p[4], ftau = p[5];
        if (t < rd) return v1;
        double r = v1 + (v1 - v2) * expm1(-(t - rd) / rtau);
        if (t < fd) return r;
        return r + (v2 - v1) * expm1(-(t - fd) / ftau);
    }
}

// ---------------------------------------------------------------------------------------------------------
// nonlinear devices: one thread evaluates one device at x and fills its slots (conductances, equivalent
// currents, terminal currents).  Returns per-thread (#non-finite g functions) | (panic << 20).
// ---------------------------------------------------------------------------------------------------------
template <class V>
__device__ __forceinline__ double xat(V x, int i)
{
    return i >= 0 ? x[i] : 0.0;
}

template <class V>
__device__ __forceinline__ int eval_device(const int *d, V x, V val)
{
    const int kind = d[0];
    const V s = val + d[4];
    if (kind == K_D) { // src/device/diode/model.rs:7-22
        const double ISAT = 1.0e-12, NVT = 1.0 * 26e-3;
        double vd = xat(x, d[1]) - xat(x, d[2]);
        double arg = vd / NVT;
        double i = ISAT * expm1(arg);
        double g = ISAT / NVT * exp(arg);
        s[D_G] = g;
        s[D_IEQ] = i - g * vd;
        s[D_I] = i;
        return isfinite(i) ? 0 : 1;
    } else if (kind == K_Q) { // src/device/npn/model.rs:8-54
        const double IES = 2e-14, ICS = 99e-14, VTE = 26e-3, VTC = 26e-3, AR = 0.02, AF = 0.99;
        double vc = xat(x, d[1]), vb = xat(x, d[2]), ve = xat(x, d[3]);
        double vbe = vb - ve, vbc = vb - vc;
        double ae = vbe / VTE, ac = vbc / VTC;
        double em1e = expm1(ae), em1c = expm1(ac), ee = exp(ae), ec = exp(ac);
        double ie = -IES * em1e + AR * ICS * em1c;
        double ic = AF * IES * em1e - ICS * em1c;
        double ib = -(ie + ic);
        double gee = IES / VTE * ee, gec = AR * ICS / VTC * ec, gce = AF * IES / VTE * ee, gcc = ICS / VTC * ec;
        double ie_eq = ie + gee * vbe - gec * vbc;
        double ic_eq = ic - gce * vbe + gcc * vbc;
        s[Q_CC] = gcc;
        s[Q_BB] = gcc + gee - gce - gec;
        s[Q_EE] = gee;
        s[Q_EC] = gec;
        s[Q_CE] = gce;
        s[Q_EB] = gec - gee;
        s[Q_BE] = gce - gee;
        s[Q_CB] = gce - gcc;
        s[Q_BC] = gec - gcc;
        s[Q_RC] = ic_eq;
        s[Q_RB] = ie_eq + ic_eq;
        s[Q_RE] = ie_eq;
        s[Q_IC] = ic;
        s[Q_IB] = ib;
        s[Q_IE] = ie;
        return (isfinite(ic) ? 0 : 1) + (isfinite(ib) ? 0 : 1) + (isfinite(ie) ? 0 : 1);
    } else { // K_M: src/device/nmos/model.rs:15-79, src/device/nmos.rs:90-134
        const double BETA = 0.5e-3, VT = 0.6, LAMBDA = 0.01;
        double vd = xat(x, d[1]), vg = xat(x, d[2]), vs = xat(x, d[3]);
        bool sw = vs > vd; // nmos.rs:105-108 / :69-71
        if (sw) {
            double t = vd;
            vd = vs;
            vs = t;
        }
        double vgs = vg - vs, vds = vd - vs;
        double id, gds, gm;
        int panic = 0;
        if (vgs <= VT) {
            id = 0.0;
            gds = 0.0;
            gm = 0.0;
        } else if (0.0 <= vds && vds <= vgs - VT) {
            id = BETA * ((vgs - VT) * vds - 0.5 * (vds * vds));
            gds = BETA * (vgs - VT - vds);
            gm = BETA * vds;
        } else if (0.0 <= vgs - VT && vgs - VT <= vds) {
            id = 0.5 * BETA * (vgs - VT) * (vgs - VT) * (1.0 + LAMBDA * vds);
            gds = 0.5 * BETA * LAMBDA * ((vgs - VT) * (vgs - VT));
            gm = BETA * (vgs - VT) * (1.0 + LAMBDA * vds);
        } else { // unreachable!() in the reference (NaN voltages)
            id = gds = gm = __longlong_as_double(0x7ff8000000000000LL);
            panic = 1;
        }
        double ieq = id - gds * vds - gm * vgs;
        if (!sw) {
            s[M_DD] = gds;
            s[M_SS] = gds + gm;
            s[M_DS] = -(gds + gm);
            s[M_SD] = -gds;
            s[M_DG] = gm;
            s[M_SG] = -gm;
            s[M_RD] = -ieq;
            s[M_RS] = ieq;
        } else { // indices swapped: (d',s') = (s,d)
            s[M_SS] = gds;
            s[M_DD] = gds + gm;
            s[M_SD] = -(gds + gm);
            s[M_DS] = -gds;
            s[M_SG] = gm;
            s[M_DG] = -gm;
            s[M_RS] = -ieq;
            s[M_RD] = ieq;
        }
        s[M_ID] = id; // residual rows use the netlist d/s (closures swap voltages only, nmos.rs:65-73)
        return (isfinite(id) ? 0 : 2) | (panic << 20);
    }
}

// evaluates every nonlinear device at x; returns team-wide nbad | panic<<20
template <class Team, class V>
__device__ __forceinline__ int eval_nl(const Team &tm, const ModeProg &M, V x, V val)
{
    if (M.n_nl == 0) return 0;
    int acc = 0;
    for (int d = tm.rank; d < M.n_nl; d += tm.size()) {
        int desc[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) desc[k] = __ldg(&M.nl[d * 5 + k]);
        acc += eval_device(desc, x, val);
    }
    acc = tm.sumi(acc);
    tm.sync();
    return acc;
}

// ---------------------------------------------------------------------------------------------------------
// residual f = A x + H g(x) - b and its class norms sqrt(sum f^2)/len  (mna.rs:25-29, node_vec_norm.rs:12-33)
// ---------------------------------------------------------------------------------------------------------
template <class Team, class W, class V>
__device__ __forceinline__ void resid_norms(const Team &tm, const ModeProg &M, const W &w, V x, int nbad, double &nv,
                                            double &ni)
{
    const int n = M.n, ld = w.ld;
    double sv = 0.0, si = 0.0;
    for (int i = tm.rank; i < n; i += tm.size()) {
        double ax = 0.0;
        for (int j = 0; j < n; ++j) ax = ax + w.A[j * ld + i] * x[j]; // separate multiply and add, like the reference's dot (mna.rs:28)
        double hg = 0.0;
        if (M.n_gfun > 0) {
            int own_bad = 0;
            int c0 = __ldg(&M.f_nl.ptr[i]), c1 = __ldg(&M.f_nl.ptr[i + 1]);
            for (int c = c0; c < c1; ++c) {
                unsigned wd = (unsigned)__ldg(&M.f_nl.con[c]);
                double v = w.val[wd & ~CON_NEG];
                own_bad += isfinite(v) ? 0 : 1;
                hg += (wd & CON_NEG) ? -v : v;
            }
            // dense H g: a non-finite g_k multiplies the zeros of every row it does not touch (0*inf = NaN)
            if (nbad > own_bad) hg = __longlong_as_double(0x7ff8000000000000LL);
        }
        double f = ax + hg - w.b[i];
        double f2 = f * f;
        if (__ldg(&M.cls[i]))
            si += f2;
        else
            sv += f2;
    }
    tm.sum2(sv, si);
    nv = sqrt(sv) / (double)M.nv;
    ni = sqrt(si) / (double)M.ni;
}

template <class Team, class V>
__device__ __forceinline__ void vec_norms(const Team &tm, const ModeProg &M, V v, double &nv, double &ni)
{
    double sv = 0.0, si = 0.0;
    for (int i = tm.rank; i < M.n; i += tm.size()) {
        double f2 = v[i] * v[i];
        if (__ldg(&M.cls[i]))
            si += f2;
        else
            sv += f2;
    }
    tm.sum2(sv, si);
    nv = sqrt(sv) / (double)M.nv;
    ni = sqrt(si) / (double)M.ni;
}

// ---------------------------------------------------------------------------------------------------------
// dense LU with partial pivoting (gauss_lu.rs:3-54).  Column-major J, thread `rank` owns row `rank`.
// Rows are not moved: each row carries its current *position*; the pivot is the candidate with the largest
// |a| and, among equals, the smallest position — the reference's "first strictly larger" rule.
// ---------------------------------------------------------------------------------------------------------
template <class Team, class V, class IV>
__device__ __forceinline__ void lu_factor(const Team &tm, int n, int ld, V J, IV piv, int &mypos)
{
    const int r = tm.rank;
    const bool have = r < n;
    mypos = r;
    // CTA teams (large n, matrix possibly in global memory): skip the updates `a - m*p` whose product is an exact
    // zero.  Guarded so that the result is the dense one bit for bit: a row skips only if its multiplier is finite,
    // and the pivot row's zero columns are skipped only if every entry of the pivot row is finite (0*inf = NaN must
    // still poison).  clist / cnt live behind piv[] (Layout::ints).
    int *clist = nullptr, *cnt = nullptr;
    if constexpr (Team::kIsCta) {
        clist = &piv[(n + 3) & ~3];
        cnt = clist + ((n + 3) & ~3);
    }
    for (int k = 0; k < n; ++k) {
        tm.sync();
        if constexpr (Team::kIsCta) { // between two barriers: after step k-1's readers, before step k's writers
            if (r == 0) {
                cnt[0] = 0; // non-zero columns of the pivot row
                cnt[1] = 0; // a non-finite entry in the pivot row
            }
        }
        double v = -1.0;
        int p = INT_MAX;
        if (have && mypos >= k) {
            v = fabs(J[k * ld + r]);
            p = mypos;
            if (v != v) v = (mypos == k) ? __longlong_as_double(0x7ff0000000000000LL) : -0.5;
        }
        tm.argmax(v, p);
        const bool iam = have && mypos == p;
        const bool atk = have && mypos == k;
        if (iam) {
            mypos = k;
            piv[k] = r;
        } else if (atk) {
            mypos = p;
        }
        tm.sync();
        const int P = piv[k];
        if constexpr (Team::kIsCta) {
            for (int j = k + 1 + r; j < n; j += tm.size()) {
                const double pv = J[j * ld + P];
                if (pv != 0.0) clist[atomicAdd(&cnt[0], 1)] = j; // NaN != 0 too
                if (!(fabs(pv) < __longlong_as_double(0x7ff0000000000000LL))) cnt[1] = 1;
            }
            tm.sync();
            if (have && mypos > k) {
                const double alpha = 1.0 / J[k * ld + P];
                const double m = alpha * J[k * ld + r];
                J[k * ld + r] = m;
                const bool mfin = fabs(m) < __longlong_as_double(0x7ff0000000000000LL);
                if (cnt[1] || !mfin) {
                    for (int j = k + 1; j < n; ++j) J[j * ld + r] -= m * J[j * ld + P];
                } else if (m != 0.0) {
                    const int nc = cnt[0];
                    for (int i = 0; i < nc; ++i) {
                        const int j = clist[i];
                        J[j * ld + r] -= m * J[j * ld + P];
                    }
                }
            }
        } else {
            if (have && mypos > k) {
                const double alpha = 1.0 / J[k * ld + P];
                const double m = alpha * J[k * ld + r];
                J[k * ld + r] = m;
                for (int j = k + 1; j < n; ++j) J[j * ld + r] -= m * J[j * ld + P];
            }
        }
    }
    tm.sync();
}

// forward + backward substitution with the stored multipliers; rhs in w (by row, destroyed), solution in xout
template <class Team, class V, class IV>
__device__ __forceinline__ void lu_solve(const Team &tm, int n, int ld, V J, IV piv, int mypos, V y, V xout)
{
    const int r = tm.rank;
    const bool have = r < n;
    for (int k = 0; k < n - 1; ++k) {
        tm.sync();
        const double yP = y[piv[k]];
        if (have && mypos > k) y[r] -= J[k * ld + r] * yP;
    }
    tm.sync();
    for (int i = n - 1; i >= 0; --i) {
        if (have && mypos == i) xout[i] = y[r] / J[i * ld + r];
        tm.sync();
        if (have && mypos < i) y[r] -= xout[i] * J[i * ld + r];
    }
    tm.sync();
}

// ---------------------------------------------------------------------------------------------------------
// damped Newton (newtons_method.rs:19-71).  Returns n_iters >= 0, or -status.
// A, b hold the linear part (incl. companions); val holds device values at x on entry? No: evaluated here.
// ---------------------------------------------------------------------------------------------------------
// A *solver* policy provides State (what persists between Newton calls: the cached factorisation of a linear
// circuit) and newton().  SmemSolver: generic, matrix in shared (or global) memory, one thread per row.
template <class Team>
struct SmemSolver {
    struct State {
        bool jvalid; // J holds a valid factorisation of the current A (linear circuits only)
        int mypos;   // this thread's row position in that factorisation
        __device__ State() : jvalid(false), mypos(0) {}
    };
    // called by the analysis drivers whenever the matrix A has been rebuilt
    static constexpr bool kHasA = true; // keeps the assembled linear part A in memory
    static constexpr bool kHasB = true; // ... and its right-hand side b
    static constexpr bool kExternalStartup = false; // .tran runs its start-up .op itself
    template <class W>
    static __device__ __forceinline__ void matrix_changed(const Team &, const ModeProg &, const W &, State &ns)
    {
        ns.jvalid = false;
    }
    template <class W>
    static __device__ int newton(const Team &tm, const ModeProg &M, const W &w, State &ns, Counters &cn, int skip);
};

template <class Team>
template <class W>
__device__ int SmemSolver<Team>::newton(const Team &tm, const ModeProg &M, const W &w, State &ns, Counters &cn, int skip)
{
    const int n = M.n, ld = w.ld;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const bool linear = (M.n_nl == 0);

    int flags = eval_nl(tm, M, w.x, w.val);
    if (flags >> 20) return -ST_NMOS_UNREACHABLE;
    double eov, eoi;
    resid_norms(tm, M, w, w.x, flags & 0xfffff, eov, eoi);
    double sov = eov, soi = eoi;
    double ev = INF, ei = INF, sv = INF, si = INF;
    bool xp_ready = false;

    int it = 0;
    while (it < 100 && !(sv < 0.001 * sov + 1e-6 && si < 0.001 * soi + 1e-9 && ev < 0.001 * eov + 1e-6 &&
                         ei < 0.001 * eoi + 1e-9)) {
        if (!(linear && skip && ns.jvalid)) {
            // J = A + nonlinear stamps at x
            for (int c = 0; c < n; ++c)
                for (int i = tm.rank; i < n; i += tm.size()) w.J[c * ld + i] = w.A[c * ld + i];
            tm.sync();
            gather(tm, M.J_nl, w.J, w.J, w.val, ld);
            lu_factor(tm, n, ld, w.J, w.piv, ns.mypos);
            ns.jvalid = linear;
            cn.c[CN_FACTOR]++;
        }
        if (!(linear && skip && xp_ready)) {
            for (int i = tm.rank; i < n; i += tm.size()) { // r = b + nonlinear rhs (row-indexed map)
                double acc = w.b[i];
                int c0 = __ldg(&M.r_nl.ptr[i]), c1 = __ldg(&M.r_nl.ptr[i + 1]);
                for (int c = c0; c < c1; ++c) {
                    unsigned wd = (unsigned)__ldg(&M.r_nl.con[c]);
                    double v = w.val[wd & ~CON_NEG];
                    acc += (wd & CON_NEG) ? -v : v;
                }
                w.r[i] = acc;
            }
            lu_solve(tm, n, ld, w.J, w.piv, ns.mypos, w.r, w.xp);
            xp_ready = true;
            cn.c[CN_SOLVE]++;
        }
        // damped step (newtons_method.rs:73-77) and its norms
        double qv = 0.0, qi = 0.0;
        for (int i = tm.rank; i < n; i += tm.size()) {
            double d = w.xp[i] - w.x[i];
            double t = (1.3 / 16.0) * copysign(1.0, d) * damp_log1p(16.0 * fabs(d));
            if (d != d) t = d;
            w.xn[i] = w.x[i] + t;
            double t2 = t * t;
            if (__ldg(&M.cls[i]))
                qi += t2;
            else
                qv += t2;
        }
        tm.sum2(qv, qi);
        sv = sqrt(qv) / (double)M.nv;
        si = sqrt(qi) / (double)M.ni;
        tm.sync();

        flags = eval_nl(tm, M, w.xn, w.val);
        if (flags >> 20) return -ST_NMOS_UNREACHABLE;
        resid_norms(tm, M, w, w.xn, flags & 0xfffff, ev, ei);
        if (isinf(ev) || isinf(ei)) return -ST_DIVERGED;

        tm.sync();
        for (int i = tm.rank; i < n; i += tm.size()) w.x[i] = w.xn[i];
        tm.sync();
        ++it;
        cn.c[CN_NEWTON]++;
        eov = ev;
        eoi = ei;
        sov = sv;
        soi = si;
    }
    return it < 100 ? it : -ST_NOT_CONVERGED;
}

// ---------------------------------------------------------------------------------------------------------
// per-instance set-up shared by the three analyses: constant slot, resistor conductances, source values at
// t = 0 (Engine::new evaluates waveforms at 0, src/engine.rs:47-51)
// ---------------------------------------------------------------------------------------------------------
template <class Team, class W>
__device__ __forceinline__ void load_instance(const Team &tm, const RunArgs &a, const W &w, long long inst)
{
    const DevProg &P = a.prog;
    const ParamView vals = vals_of(a, inst), fn = fns_of(a, inst);
    if (tm.rank == 0) w.val[SLOT_ONE] = 1.0;
    for (int i = tm.rank; i < P.n_res; i += tm.size()) {
        int e = __ldg(&P.res[2 * i]), s = __ldg(&P.res[2 * i + 1]);
        w.val[s] = 1.0 / vals[e];
    }
    for (int i = tm.rank; i < P.n_src; i += tm.size()) {
        const int *sd = &P.src[5 * i];
        int e = __ldg(&sd[0]), s = __ldg(&sd[1]), fk = __ldg(&sd[3]), fs = __ldg(&sd[4]);
        double v = fk != FN_NONE ? spice_fn_eval(fk, fn + fs * 7, 0.0) : vals[e];
        w.val[s] = v;
        if (w.prev) { // source history, .tran only
            w.prev[3 * i] = v;     // value at t = 0
            w.prev[3 * i + 1] = v; // value at the element's last evaluation before this step
            w.prev[3 * i + 2] = v; // value at the current attempt
        }
    }
    tm.sync();
}

template <class Team, class W>
__device__ __forceinline__ void build_linear(const Team &tm, const ModeProg &M, const W &w, bool has_a, bool has_b)
{
    const int n = M.n, ld = w.ld;
    using V = decltype(w.b);
    if (has_a)
        for (int c = 0; c < n; ++c)
            for (int i = tm.rank; i < n; i += tm.size()) w.A[c * ld + i] = 0.0;
    tm.sync();
    if (has_a) gather(tm, M.A_lin, w.A, V{nullptr}, w.val, ld);
    if (has_b) gather(tm, M.b_all, w.b, V{nullptr}, w.val); // row-indexed: writes every row
    tm.sync();
}

template <class Team>
__device__ void flush_counters(const Team &tm, const RunArgs &a, const Counters &cn)
{
    if (tm.rank == 0) {
        for (int i = 0; i < CN_LOCAL; ++i)
            if (cn.c[i]) atomicAdd(&a.counters[i], cn.c[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Engine::run_op  (src/engine.rs:62-90) for one instance.  Leaves x (start-up numbering) in w.x.
// ---------------------------------------------------------------------------------------------------------
template <class Team, class Solver, class W>
__device__ __forceinline__ int solve_op(const Team &tm, const RunArgs &a, const W &w, Counters &cn)
{
    const ModeProg &M = a.prog.op;
    build_linear(tm, M, w, Solver::kHasA, Solver::kHasB);
    for (int i = tm.rank; i < M.n; i += tm.size()) w.x[i] = 0.0;
    tm.sync();
    typename Solver::State ns;
    Solver::matrix_changed(tm, M, w, ns);
    return Solver::newton(tm, M, w, ns, cn, a.skip_refactor);
}

template <class Team, class Solver, class W>
__device__ __forceinline__ void run_op_instance(const Team &tm, const RunArgs &a, const W &w, long long inst, Counters &cn)
{
    load_instance(tm, a, w, inst);
    int r = solve_op<Team, Solver, W>(tm, a, w, cn);
    const int n = a.prog.op.n;
    for (int i = tm.rank; i < n; i += tm.size()) a.x[inst * n + i] = w.x[i];
    if (tm.rank == 0) {
        a.n_iters[inst] = r >= 0 ? r : 0;
        a.status[inst] = r >= 0 ? ST_OK : -r;
    }
    if (r >= 0) cn.c[CN_POINTS]++;
}

// ---------------------------------------------------------------------------------------------------------
// Engine::run_dc  (src/engine.rs:92-140): one warm-started chain per instance
// ---------------------------------------------------------------------------------------------------------
template <class Team, class Solver, class W>
__device__ __forceinline__ void run_dc_instance(const Team &tm, const RunArgs &a, const W &w, long long inst, Counters &cn)
{
    const ModeProg &M = a.prog.dc;
    const int n = M.n;
    load_instance(tm, a, w, inst);
    build_linear(tm, M, w, Solver::kHasA, Solver::kHasB);
    for (int i = tm.rank; i < n; i += tm.size()) w.x[i] = 0.0;
    tm.sync();
    const double start = a.start[inst], stop = a.stop[inst], step = a.step[inst];
    const double cnt_d = ceil((stop - start) / step); // ndarray Array::range (engine.rs:120)
    long long count = cnt_d >= 0.0 ? (long long)cnt_d : 0;
    if (count > a.max_points) count = a.max_points;
    const int sweep_slot = __ldg(&a.prog.src[5 * a.sweep_src + 1]);
    typename Solver::State ns;
    Solver::matrix_changed(tm, M, w, ns);
    int status = ST_OK;
    long long k = 0;
    for (; k < count; ++k) {
        const double sv = start + step * (double)k;
        if (tm.rank == 0) w.val[sweep_slot] = sv;
        tm.sync();
        if (Solver::kHasB) gather(tm, M.b_all, w.b, decltype(w.b){nullptr}, w.val);
        tm.sync();
        int r = Solver::newton(tm, M, w, ns, cn, a.skip_refactor);
        if (r < 0) { // `?` aborts the whole sweep (engine.rs:127)
            status = -r;
            break;
        }
        double *xo = a.x + (inst * a.max_points + k) * n;
        for (int i = tm.rank; i < n; i += tm.size()) xo[i] = w.x[i];
        if (tm.rank == 0) a.n_iters[inst * a.max_points + k] = r;
        cn.c[CN_POINTS]++;
    }
    if (tm.rank == 0) {
        a.points_done[inst] = k;
        a.status[inst] = status;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Engine::run_tran  (src/engine.rs:142-206) + transient::step (src/engine/transient.rs:18-97), stateless
// form of SURVEY.md App. D: companions frozen at the first attempt's h (Q10), x never restored (Q11),
// current sources drift (Q9), PLTE over a ring of the last three accepted points.
// ---------------------------------------------------------------------------------------------------------
template <class Team, class Solver, class W>
__device__ __forceinline__ void run_tran_instance(const Team &tm, const RunArgs &a, const W &w, long long inst, Counters &cn)
{
    const DevProg &P = a.prog;
    const ModeProg &M = P.tran;
    const int n = M.n, ld = w.ld;
    const ParamView vals = vals_of(a, inst), fn = fns_of(a, inst);

    load_instance(tm, a, w, inst);
    long long nrec = 0, stored = 0;
    int status = ST_OK;

    // start-up solution: a full .op in the start-up numbering (engine.rs:158-161, Q13)
    int r0 = 0;
    if constexpr (Solver::kExternalStartup) {
        const int st0 = a.start_status[inst];
        r0 = st0 == ST_OK ? 0 : -st0;
        for (int i = tm.rank; i < n; i += tm.size()) w.x[i] = a.x_start[inst * a.n_start + i];
        tm.sync();
    } else {
        r0 = solve_op<Team, Solver, W>(tm, a, w, cn);
    }
    if (r0 < 0) {
        status = -r0;
    } else {
        // init_state (cap.rs:38-47, ind.rs:42-49)
        for (int i = tm.rank; i < P.n_dyn; i += tm.size()) {
            const int *dd = &P.dyn[6 * i];
            int kind = __ldg(&dd[2]), vp = __ldg(&dd[3]), vn = __ldg(&dd[4]), br = __ldg(&dd[5]);
            if (kind == K_C) {
                w.state[2 * i] = xat(w.x, vp) - xat(w.x, vn);
                w.state[2 * i + 1] = 0.0;
            } else {
                w.state[2 * i] = 0.0;
                if constexpr (Solver::kExternalStartup)
                    w.state[2 * i + 1] = a.x_start[inst * a.n_start + br];
                else
                    w.state[2 * i + 1] = w.x[br];
            }
        }
        tm.sync();
        // x (run numbering) is the prefix of the start-up x: nothing to move.  Linear part for .tran:
        build_linear(tm, M, w, Solver::kHasA, Solver::kHasB);

        double t = 0.0, h = 1e-18;
        double tr0 = 0.0, tr1 = 0.0, tr2 = 0.0; // times of the last three accepted records (oldest first)
        double hA = -1.0;                        // h the companions in A were built with
        typename Solver::State ns;
        const double c3 = -1.0 / 12.0;

        while (t < a.t_stop && status == ST_OK) {
            const double h_first = h;
            // companion models at the first attempt's h (cap/model.rs:10-17, ind/model.rs:10-17)
            for (int i = tm.rank; i < P.n_dyn; i += tm.size()) {
                const int *dd = &P.dyn[6 * i];
                int e = __ldg(&dd[0]), s = __ldg(&dd[1]), kind = __ldg(&dd[2]);
                double u = w.state[2 * i], ic = w.state[2 * i + 1];
                if (kind == K_C) {
                    double g = 2.0 * vals[e] / h_first;
                    w.val[s] = g;
                    w.val[s + 1] = -(ic + g * u);
                } else {
                    double g = h_first / (2.0 * vals[e]);
                    w.val[s] = g;
                    w.val[s + 1] = ic + g * u;
                }
            }
            tm.sync();
            if (h_first != hA) {
                if (Solver::kHasA) gather(tm, M.A_dyn, w.A, decltype(w.A){nullptr}, w.val, ld);
                hA = h_first;
                tm.sync();
                Solver::matrix_changed(tm, M, w, ns);
            }
            double next_h = h;
            bool accepted = false;
            int nit = 0;
            while (!accepted) {
                const double te = t + h;
                // time-varying sources at t + h (transient.rs:36-42)
                for (int i = tm.rank; i < P.n_src; i += tm.size()) {
                    const int *sd = &P.src[5 * i];
                    int s = __ldg(&sd[1]), kind = __ldg(&sd[2]), fk = __ldg(&sd[3]), fs = __ldg(&sd[4]);
                    if (fk != FN_NONE) {
                        double v = spice_fn_eval(fk, fn + fs * 7, te);
                        w.prev[3 * i + 2] = v;
                        // V: assignment (Q8).  I: b keeps I(0) - I(t_entry) + I(t+h) (Q9)
                        w.val[s] = kind == K_V ? v : (w.prev[3 * i] - w.prev[3 * i + 1]) + v;
                    }
                }
                tm.sync();
                if (Solver::kHasB) gather(tm, M.b_all, w.b, decltype(w.b){nullptr}, w.val);
                tm.sync();
                int r = Solver::newton(tm, M, w, ns, cn, a.skip_refactor);
                if (r == -ST_NOT_CONVERGED) {
                    h *= 0.5;
                    cn.c[CN_REJ_NEWTON]++;
                } else if (r < 0) {
                    status = -r;
                    break;
                } else if (nrec < 3) {
                    accepted = true;
                    nit = r;
                } else {
                    // PLTE = -0.5 h^3 dd3 over (ring0, ring1, ring2, candidate)  (state_history.rs:40-56)
                    const double hn = te - tr2;
                    const double alpha = 6.0 * c3 * (hn * hn * hn);
                    const double a01 = 1.0 / (tr1 - tr0), a12 = 1.0 / (tr2 - tr1), a23 = 1.0 / (te - tr2);
                    const double a02 = 1.0 / (tr2 - tr0), a13 = 1.0 / (te - tr1), a03 = 1.0 / (te - tr0);
                    const int o0 = (int)(nrec % 3), o1 = (int)((nrec + 1) % 3), o2 = (int)((nrec + 2) % 3);
                    double pv = 0.0, pi = 0.0, xv = 0.0, xi = 0.0;
                    for (int i = tm.rank; i < n; i += tm.size()) {
                        double x0 = w.ring[o0 * ld + i], x1 = w.ring[o1 * ld + i], x2 = w.ring[o2 * ld + i];
                        double x3 = w.x[i];
                        double d01 = a01 * (x1 - x0), d12 = a12 * (x2 - x1), d23 = a23 * (x3 - x2);
                        double d02 = a02 * (d12 - d01), d13 = a13 * (d23 - d12);
                        double d03 = a03 * (d13 - d02);
                        double pl = alpha * d03;
                        if (__ldg(&M.cls[i])) {
                            pi += pl * pl;
                            xi += x3 * x3;
                        } else {
                            pv += pl * pl;
                            xv += x3 * x3;
                        }
                    }
                    tm.sum2(pv, pi);
                    tm.sum2(xv, xi);
                    const double plv = sqrt(pv) / (double)M.nv, pli = sqrt(pi) / (double)M.ni;
                    const double xnv = sqrt(xv) / (double)M.nv, xni = sqrt(xi) / (double)M.ni;
                    const bool too_big = plv > xnv * 0.001 + 1e-3 || pli > xni * 0.001 + 1e-6; // transient.rs:99-101
                    if (too_big) {
                        h *= 0.5;
                        cn.c[CN_REJ_PLTE]++;
                    } else {
                        accepted = true;
                        nit = r;
                        const bool can_grow = plv < 0.1 * 1e-3 && pli < 0.1 * 1e-6; // :103-105
                        next_h = (can_grow && h <= a.step_max / 2.0) ? h * 2.0 : h;
                    }
                }
                if (!accepted && h < 1e-18) { // transient.rs:88-90
                    status = ST_TRAN_H_UNDERFLOW;
                    break;
                }
            }
            if (status != ST_OK) break;
            if (nrec < 3) next_h = h;

            // accept: push the record (state_history.rs:24-30) and write it out per the recording policy
            const double te = t + h;
            {
                const int slot = (int)(nrec % 3); // overwrites the oldest
                for (int i = tm.rank; i < n; i += tm.size()) w.ring[slot * ld + i] = w.x[i];
                tr0 = tr1;
                tr1 = tr2;
                tr2 = te;
                const int stride = a.rec_stride > 1 ? a.rec_stride : 1;
                if (stored < a.cap && (nrec % stride) == 0) {
                    const long long row = inst * a.cap + stored;
                    if (a.x)
                        for (int i = tm.rank; i < n; i += tm.size()) a.x[row * n + i] = w.x[i];
                    if (tm.rank == 0) {
                        if (a.t_out) a.t_out[row] = te;
                        if (a.n_iters) a.n_iters[row] = nit;
                    }
                    ++stored;
                }
                ++nrec;
                cn.c[CN_POINTS]++;
            }
            // sources keep the value of their last evaluation (vdd.rs:40-44, idd.rs:40-44)
            for (int i = tm.rank; i < P.n_src; i += tm.size()) w.prev[3 * i + 1] = w.prev[3 * i + 2];
            t += h;
            // update_state with the accepted h (engine.rs:183-185; cap/model.rs:19-25, ind/model.rs:19-25)
            for (int i = tm.rank; i < P.n_dyn; i += tm.size()) {
                const int *dd = &P.dyn[6 * i];
                int e = __ldg(&dd[0]), kind = __ldg(&dd[2]), vp = __ldg(&dd[3]), vn = __ldg(&dd[4]);
                double u_old = w.state[2 * i], i_old = w.state[2 * i + 1];
                double u_new = xat(w.x, vp) - xat(w.x, vn);
                double i_new;
                if (kind == K_C) {
                    double g = 2.0 * vals[e] / h;
                    i_new = g * (u_new - u_old) - i_old;
                } else {
                    double g = h / (2.0 * vals[e]);
                    i_new = g * (u_new + u_old) + i_old;
                }
                w.state[2 * i] = u_new;
                w.state[2 * i + 1] = i_new;
            }
            tm.sync();
            h = next_h;
        }
    }
    if (a.x_final)
        for (int i = tm.rank; i < n; i += tm.size()) a.x_final[inst * n + i] = w.x[i];
    if (tm.rank == 0) {
        a.n_rec[inst] = nrec;
        a.status[inst] = status;
    }
}

// ---------------------------------------------------------------------------------------------------------
// kernels: persistent grid, teams pull instances from an atomic counter (data-dependent run times)
// ---------------------------------------------------------------------------------------------------------
enum : int { AN_OP = 0, AN_DC = 1, AN_TRAN = 2 };

__device__ __forceinline__ Ws make_ws(const Layout &L, double *vec, double *mat, int *ints)
{
    Ws w;
    w.ld = L.ld;
    w.b.p = vec + L.ob;
    w.r.p = vec + L.orr;
    w.x.p = vec + L.ox;
    w.xn.p = vec + L.oxn;
    w.xp.p = vec + L.oxp;
    w.val.p = vec + L.oval;
    w.state.p = vec + L.ostate;
    w.ring.p = vec + L.oring;
    w.prev.p = vec + L.oprev;
    w.A.p = mat;
    w.J.p = mat + L.mat_doubles;
    w.piv.p = ints;
    return w;
}

template <class Team, class Solver, int AN, class W>
__device__ __forceinline__ void run_instance(const Team &tm, const RunArgs &a, const W &w, long long inst, Counters &cn)
{
    if (AN == AN_OP)
        run_op_instance<Team, Solver, W>(tm, a, w, inst, cn);
    else if (AN == AN_DC)
        run_dc_instance<Team, Solver, W>(tm, a, w, inst, cn);
    else
        run_tran_instance<Team, Solver, W>(tm, a, w, inst, cn);
}

template <int G, int AN>
__global__ void __launch_bounds__(128) k_subwarp(const __grid_constant__ RunArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SubWarp<G> tm;
    const Layout &L = a.prog.lay;
    const int team_in_block = threadIdx.x / G;
    unsigned char *base = smem_raw + (size_t)team_in_block * L.bytes;
    double *vec = reinterpret_cast<double *>(base);
    double *mat = vec + L.doubles;
    int *ints = reinterpret_cast<int *>(mat + 2 * L.mat_doubles);
    Ws w = make_ws(L, vec, mat, ints);
    Counters cn;
    for (;;) {
        long long inst = 0;
        if (tm.rank == 0) inst = next_instance(a);
        inst = tm.bcast_ll(inst);
        if (inst < 0) break;
        run_instance<SubWarp<G>, SmemSolver<SubWarp<G>>, AN>(tm, a, w, inst, cn);
        tm.sync();
    }
    flush_counters(tm, a, cn);
}

template <int AN>
__global__ void k_cta(const __grid_constant__ RunArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Layout &L = a.prog.lay;
    double *red = reinterpret_cast<double *>(smem_raw); // 96 doubles of reduction scratch
    CtaTeam tm(red);
    double *vec = red + 96;
    double *mat = L.mat_in_smem ? vec + L.doubles : a.scratch + (size_t)blockIdx.x * 2 * L.mat_doubles;
    int *ints = reinterpret_cast<int *>(vec + L.doubles + (L.mat_in_smem ? 2 * L.mat_doubles : 0));
    Ws w = make_ws(L, vec, mat, ints);
    Counters cn;
    for (;;) {
        long long inst = 0;
        if (threadIdx.x == 0) inst = next_instance(a);
        inst = tm.bcast_ll(inst);
        if (inst < 0) break;
        run_instance<CtaTeam, SmemSolver<CtaTeam>, AN>(tm, a, w, inst, cn);
        tm.sync();
    }
    flush_counters(tm, a, cn);
}

} // namespace ftsb
