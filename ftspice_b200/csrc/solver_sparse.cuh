// solver_sparse.cuh — structure-exploiting elimination, one WARP per instance, any n (16 < n <= 4096).
//
// MNA matrices are very sparse (C4 ladder, n = 65: 192 structural non-zeros; C5 inverter chain, n = 258: 1 021 of
// 66 564) and the reference's dense elimination (gauss_lu.rs:29-42) spends almost all of its n^3/3 updates on
// `a - 0*p`.  Those leave `a` unchanged, so an elimination that touches only non-zero multipliers and non-zero
// pivot-row entries — with the reference's pivot rule (first position with the strictly largest |a|), reciprocal
// scaling, separate multiply / subtract and update order — returns the reference's numbers (up to the sign of a
// zero) at a cost that grows with the number of non-zeros: C5 5.7 M element updates -> ~500.
//
// Storage per instance (shared memory): one f64 per entry of the *static* structure (J_full's pattern plus the
// fill-in a symbolic elimination predicts at circuit-compile time, compile.cpp) and a pool of run-time fill-in
// entries (partial pivoting is value dependent, so instances may deviate from the prediction) chained per row and
// per column.  Rows never move: each carries its position.  The static structure itself (column lists, row
// lists) is shared by the warps of a block.
//
// An instance whose elimination meets a zero / infinite / NaN pivot candidate, a non-finite solution component, or
// runs out of fill-in pool, is abandoned (status ST_REDO) and put on a redo list: the dense kernels then run it
// from scratch with the reference's full arithmetic (0*inf = NaN poisoning and all), so observable behaviour in
// those cases is the reference's too.
#pragma once
#include "solver_warp.cuh"

namespace ftsb {

typedef unsigned short u16;
constexpr u16 SP_NONE = 0xffff;
constexpr int ST_REDO = 100; // internal: never returned to the caller

struct SpLayout {
    int osv, oy, ox, oxn, oxp, obv, oval, oprev, ostate, oring, doubles; // in doubles
    int hist_doubles; // > 0: that many doubles per resident warp live in global scratch, not here: the per-step history
                      // (state + ring; bit 0 of `glob`) and / or the slot table (bit 1)
    int val_global;   // the slot table is in the global region (offset oval there)
    int vec_global;   // x, x_new, x_proposed and b are in the global region (bit 2); tval then has its own shared array
    int otval;
    int ofr, ofc, ofnr, ofnc, orh, och, opos, opiv, ope, otcol, halfs;      // in u16 after the doubles
    int words;                                                               // fill signature: one u32 per row, after the u16s
    int fcap;
    size_t bytes;
};

__host__ __device__ inline SpLayout sp_layout(int n, int nnz, int fcap, int n_slots, int n_dyn, int n_src, bool tran,
                                              int glob = 0)
{
    SpLayout t;
    int o = 0;
    auto take = [&](int c) {
        int r = o;
        o += (c + 1) & ~1;
        return r;
    };
    t.fcap = fcap;
    t.osv = take(nnz + fcap);
    t.oy = take(n);
    int go = 0; // offsets inside the global region
    auto takeg = [&](int c) {
        int r = go;
        go += (c + 1) & ~1;
        return r;
    };
    t.val_global = (glob & 2) ? 1 : 0;
    t.vec_global = (glob & 4) ? 1 : 0;
    if (t.vec_global) { // only the rhs (read and written inside the sequential elimination) stays on chip
        t.ox = takeg(n);
        t.oxn = takeg(n);
        t.oxp = takeg(n);
        t.obv = takeg(n);
        t.otval = take(n);
    } else {
        t.ox = take(n);
        t.oxn = take(n);
        t.oxp = take(n);
        t.obv = take(n); // b of the current solve (the residual subtracts it every iteration)
        t.otval = t.oxn; // scratch during a factorisation: x_new is dead then
    }
    t.oval = t.val_global ? takeg(n_slots) : take(n_slots);
    t.oprev = take(tran ? 3 * n_src : 0);
    if (tran && (glob & 1)) {
        t.ostate = takeg(2 * n_dyn);
        t.oring = takeg(3 * n);
    } else {
        t.ostate = take(tran ? 2 * n_dyn : 0);
        t.oring = take(tran ? 3 * n : 0);
    }
    t.hist_doubles = go;
    t.doubles = o;
    int h = 0;
    auto takeh = [&](int c) {
        int r = h;
        h += c;
        return r;
    };
    t.ofr = takeh(fcap);
    t.ofc = takeh(fcap);
    t.ofnr = takeh(fcap);
    t.ofnc = takeh(fcap);
    t.orh = takeh(n);
    t.och = takeh(n);
    t.opos = takeh(n);
    t.opiv = takeh(n);
    t.ope = takeh(n);
    t.otcol = takeh(n);
    t.halfs = (h + 7) & ~7;
    t.words = (n + 3) & ~3;
    t.bytes = (size_t)t.doubles * 8 + (size_t)t.halfs * 2 + (size_t)t.words * 4;
    return t;
}

// u16 copy of the static structure, shared by the block:
// cptr[n+1] rptr[n+1] rlong[n] erow[nnz] rcol[nnz] rent[nnz] lmap[n_long][n]
__host__ __device__ inline size_t sp_static_bytes(int n, int nnz, int n_long)
{
    return (((size_t)2 * (n + 1) + n + 3 * (size_t)nnz + (size_t)n_long * n) * 2 + 15) & ~(size_t)15;
}

struct SpStatic {
    const u16 *cptr, *rptr, *rlong, *erow, *rcol, *rent, *lmap;
    int n, nnz;
};

struct SpInst {
    double *sv, *tval;
    u16 *fr, *fc, *fnr, *fnc, *rhead, *chead, *pos, *piv, *pe, *tcol;
    unsigned *sig; // bit (j & 31) of sig[r]: row r may hold a run-time fill-in entry in a column congruent to j
    int fcap;
    int windowed; // try to eliminate windows of independent columns at once
};

// WIN: compile the windowed elimination / back substitution in (separate kernels: the extra code costs registers)
template <bool WIN>
struct SparseWarpSolverT {
    using Team = WarpTeam;
    static constexpr bool kHasA = false;
    static constexpr bool kHasB = false;
    static constexpr bool kExternalStartup = true;
    struct State {
        bool jvalid;
        int nfill; // run-time fill-in entries of the current factorisation
        __device__ State() : jvalid(false), nfill(0) {}
    };
    template <class W>
    static __device__ __forceinline__ void matrix_changed(const Team &, const ModeProg &, const W &, State &s)
    {
        s.jvalid = false;
    }

    template <class W>
    static __device__ __forceinline__ void views(const ModeProg &M, const W &w, SpStatic &st, SpInst &in)
    {
        const int n = M.n, nnz = M.sp.nnz, fcap = M.sp.fcap;
        const u16 *sb = reinterpret_cast<const u16 *>(w.aux);
        st.n = n;
        st.nnz = nnz;
        st.cptr = sb;
        st.rptr = sb + (n + 1);
        st.rlong = sb + 2 * (n + 1);
        st.erow = st.rlong + n;
        st.rcol = st.erow + nnz;
        st.rent = st.rcol + nnz;
        st.lmap = st.rent + nnz;
        in.fcap = fcap;
        in.windowed = M.sp.windowed;
        in.sv = w.J.p;
        in.tval = w.A.p; // pivot-row scratch of a factorisation (aliases x_new unless the vectors live in global memory)
        u16 *hb = reinterpret_cast<u16 *>(w.piv.p);
        in.fr = hb;
        in.fc = in.fr + fcap;
        in.fnr = in.fc + fcap;
        in.fnc = in.fnr + fcap;
        in.rhead = in.fnc + fcap;
        in.chead = in.rhead + n;
        in.pos = in.chead + n;
        in.piv = in.pos + n;
        in.pe = in.piv + n;
        in.tcol = in.pe + n;
        const int halfs = (4 * fcap + 6 * n + 7) & ~7;
        in.sig = reinterpret_cast<unsigned *>(hb + halfs);
    }

    // entry index of (r, j) in the static row list or the row's fill chain; -1 when absent
    static __device__ __forceinline__ int lookup(const SpStatic &st, const SpInst &in, int r, int j)
    {
        const u16 li = st.rlong[r];
        if (li != SP_NONE) { // long row: direct map
            const u16 e = st.lmap[(int)li * st.n + j];
            if (e != SP_NONE) return (int)e;
        } else {
            int q0 = st.rptr[r], q1 = st.rptr[r + 1];
            if (q1 - q0 > 8) { // columns ascending: bisect down to a short range
                while (q1 - q0 > 4) {
                    const int mid = (q0 + q1) >> 1;
                    if ((int)st.rcol[mid] <= j)
                        q0 = mid;
                    else
                        q1 = mid;
                }
            }
            for (int t = q0; t < q1; ++t)
                if ((int)st.rcol[t] == j) return (int)st.rent[t];
        }
        if (!((in.sig[r] >> (j & 31)) & 1u)) return -1; // no fill-in entry there
        for (u16 f = in.rhead[r]; f != SP_NONE; f = in.fnr[f])
            if ((int)in.fc[f] == j) return st.nnz + (int)f;
        return -1;
    }

    // (largest v, then smallest pos) over the warp; v >= 0 or -1 (no candidate): compare the bit patterns as integers
    static __device__ __forceinline__ void warp_argmax(double &v, int &p, int &r, int &e)
    {
        const long long key = v < 0.0 ? -1LL : __double_as_longlong(v);
        const unsigned hi = (unsigned)((unsigned long long)(key + 1) >> 32), lo = (unsigned)(unsigned long long)(key + 1);
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const bool ch = hi == mh;
        const unsigned ml = __reduce_max_sync(0xffffffffu, ch ? lo : 0u);
        const bool cl = ch && lo == ml;
        const unsigned mp = __reduce_min_sync(0xffffffffu, cl ? (unsigned)p : 0xffffffffu);
        const unsigned who = __ballot_sync(0xffffffffu, cl && (unsigned)p == mp);
        const int src = __ffs((int)who) - 1;
        v = __shfl_sync(0xffffffffu, v, src);
        p = __shfl_sync(0xffffffffu, p, src);
        r = __shfl_sync(0xffffffffu, r, src);
        e = __shfl_sync(0xffffffffu, e, src);
    }

    // Up to 32 consecutive elimination steps at once.  Lane l examines column k0 + l on the matrix as it stands: pivot
    // search, multipliers, the entries its updates would touch.  That is the sequential algorithm's view of the
    // column if (a) every earlier column of the window pivots on the row already sitting at its position (no swap:
    // positions stay as they are) and (b) no earlier pivot row of the window holds an entry in this column (then no
    // update of the window writes into it).  The window is cut at the first column for which that cannot be shown —
    // or that needs a swap, fill-in, more than two targets / pivot-row entries, or meets a zero / non-finite
    // candidate — and what is left is applied in column order by one lane, so that every sum is accumulated in the
    // sequential order (the rail column of an inverter chain is a true recurrence).  Returns the columns done.
    static __device__ __forceinline__ int window(const Team &tm, const SpStatic &st, const SpInst &in, double *__restrict__ y,
                                                 int k0, unsigned long long &flops)
    {
        const int n = st.n, nnz = st.nnz, lane = tm.rank;
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        const int nw = min(32, n - k0), k = k0 + lane;
        bool valid = lane < nw;
        int P = -1, eP = -1, nt = 0, npv = 0;
        int tr[2] = {0, 0}, te[2] = {0, 0}, pe2[2] = {0, 0}, pj[2] = {0, 0}, e2[4] = {0, 0, 0, 0};
        double tmul[2] = {0.0, 0.0};
        bool tnz[2] = {false, false};
        unsigned dirty = 0;
        int c0 = 0, c1 = 0;
        if (valid) {
            c0 = st.cptr[k];
            c1 = st.cptr[k + 1];
            double bv = -1.0;
            int bp = 0x7fffffff;
            bool bad = false;
            for (int t = c0; t < c1; ++t) {
                const int r = st.erow[t], p = in.pos[r];
                if (p >= k) {
                    const double v = fabs(in.sv[t]);
                    bad |= !(v < INF);
                    if (v > bv || (v == bv && p < bp)) {
                        bv = v;
                        bp = p;
                        P = r;
                        eP = t;
                    }
                }
            }
            for (u16 f = in.chead[k]; f != SP_NONE; f = in.fnc[f]) {
                const int r = in.fr[f], p = in.pos[r];
                if (p >= k) {
                    const double v = fabs(in.sv[nnz + f]);
                    bad |= !(v < INF);
                    if (v > bv || (v == bv && p < bp)) {
                        bv = v;
                        bp = p;
                        P = r;
                        eP = nnz + (int)f;
                    }
                }
            }
            valid = !bad && bv > 0.0 && bp == k; // the pivot sits at its position already
        }
        if (valid) { // pivot row: entries with column > k
            for (int t = st.rptr[P]; t < st.rptr[P + 1]; ++t) {
                const int j = st.rcol[t];
                if (j > k) {
                    if (npv < 2) {
                        pj[npv] = j;
                        pe2[npv] = st.rent[t];
                    }
                    ++npv;
                    if (j < k0 + nw) dirty |= 1u << (j - k0);
                }
            }
            for (u16 f = in.rhead[P]; f != SP_NONE; f = in.fnr[f]) {
                const int j = in.fc[f];
                if (j > k) {
                    if (npv < 2) {
                        pj[npv] = j;
                        pe2[npv] = nnz + (int)f;
                    }
                    ++npv;
                    if (j < k0 + nw) dirty |= 1u << (j - k0);
                }
            }
            valid = npv <= 2;
        }
        if (valid) { // rows below: multipliers and the entries their updates touch
            const double alpha = 1.0 / in.sv[eP]; // (:31)
            for (int t = c0; t < c1; ++t) {
                const int r = st.erow[t];
                if ((int)in.pos[r] > k) {
                    if (nt < 2) {
                        const double a = in.sv[t];
                        tr[nt] = r;
                        te[nt] = t;
                        tmul[nt] = alpha * a;
                        tnz[nt] = a != 0.0;
                    }
                    ++nt;
                }
            }
            for (u16 f = in.chead[k]; f != SP_NONE; f = in.fnc[f]) {
                const int r = in.fr[f];
                if ((int)in.pos[r] > k) {
                    if (nt < 2) {
                        const double a = in.sv[nnz + f];
                        tr[nt] = r;
                        te[nt] = nnz + (int)f;
                        tmul[nt] = alpha * a;
                        tnz[nt] = a != 0.0;
                    }
                    ++nt;
                }
            }
            valid = nt <= 2;
            if (valid) {
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        if (t < nt && tnz[t] && q < npv) {
                            const int e = lookup(st, in, tr[t], pj[q]);
                            e2[t * 2 + q] = e;
                            valid = valid && e >= 0; // a missing entry would be fill-in: the sequential step creates it
                        }
            }
        }
        // extent of the window
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        int ext = __ffs((int)~vmask) - 1; // first lane that is not valid (32 if none)
        if (ext < 0) ext = 32;
        ext = min(ext, nw);
        const unsigned dm = __reduce_or_sync(0xffffffffu, lane < ext ? dirty : 0u);
        if (dm) ext = min(ext, __ffs((int)dm) - 1);
        if (ext <= 0) return 0;
        // multipliers (every lane its own column's entries)
        if (lane < ext) {
            in.pe[k] = (u16)eP;
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (t < nt) in.sv[te[t]] = tmul[t];
        }
        __syncwarp();
        // the updates, in column order; lane 0 does the arithmetic, the owners supply the operands
        for (int c = 0; c < ext; ++c) {
            const int Pc = __shfl_sync(0xffffffffu, P, c);
            const int ntc = __shfl_sync(0xffffffffu, nt, c), npc = __shfl_sync(0xffffffffu, npv, c);
            const double yk = y[Pc];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int r = __shfl_sync(0xffffffffu, tr[t], c);
                const double m = __shfl_sync(0xffffffffu, tmul[t], c);
                const int nz = __shfl_sync(0xffffffffu, (int)tnz[t], c);
                if (t < ntc && nz) {
                    flops += 3 + 2 * npc;
                    if (lane == 0) y[r] = y[r] - m * yk; // (:40)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int ed = __shfl_sync(0xffffffffu, e2[t * 2 + q], c);
                        const int es = __shfl_sync(0xffffffffu, pe2[q], c);
                        if (q < npc && lane == 0) in.sv[ed] = in.sv[ed] - m * in.sv[es]; // (:35-39)
                    }
                }
            }
        }
        __syncwarp();
        return ext;
    }

    // gauss_lu.rs:3-42 on the sparse storage, rhs y carried along.  Returns false -> abandon the instance (ST_REDO).
    // (0 = ok, 1 = zero / non-finite pivot candidate, 2 = fill-in pool exhausted)
    // flops: executed floating-point operations; nfill_out: fill-in entries created
    static __device__ int factor(const Team &tm, const SpStatic &st, const SpInst &in, double *__restrict__ y,
                                 unsigned long long &flops, int &nfill_out)
    {
        const int n = st.n, nnz = st.nnz, lane = tm.rank;
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        for (int i = lane; i < n; i += 32) {
            in.pos[i] = (u16)i;
            in.piv[i] = (u16)i;
            in.rhead[i] = SP_NONE;
            in.chead[i] = SP_NONE;
        }
        for (int i = lane; i < n; i += 32) in.sig[i] = 0u;
        int nfill = 0;
        __syncwarp();
        for (int k = 0; k < n; ++k) {
            if (WIN && in.windowed) { // as many independent columns as can be shown independent, at once
                const int done = window(tm, st, in, y, k, flops);
                if (done > 0) {
                    k += done - 1;
                    continue;
                }
            }
            const int c0 = st.cptr[k], c1 = st.cptr[k + 1];
            // ---- pivot search over the entries of column k held by rows at positions >= k ----
            double bv = -1.0;
            int bp = 0x7fffffff, br = -1, be = -1;
            bool bad = false;
            for (int t = c0 + lane; t < c1; t += 32) {
                const int r = st.erow[t], p = in.pos[r];
                if (p >= k) {
                    const double v = fabs(in.sv[t]);
                    bad |= !(v < INF);
                    if (v > bv || (v == bv && p < bp)) {
                        bv = v;
                        bp = p;
                        br = r;
                        be = t;
                    }
                }
            }
            {
                int i = 0;
                for (u16 f = in.chead[k]; f != SP_NONE; f = in.fnc[f], ++i) {
                    if ((i & 31) == lane) {
                        const int r = in.fr[f], p = in.pos[r];
                        if (p >= k) {
                            const double v = fabs(in.sv[nnz + f]);
                            bad |= !(v < INF);
                            if (v > bv || (v == bv && p < bp)) {
                                bv = v;
                                bp = p;
                                br = r;
                                be = nnz + (int)f;
                            }
                        }
                    }
                }
            }
            warp_argmax(bv, bp, br, be);
            if (__any_sync(0xffffffffu, bad) || !(bv > 0.0)) return 1;
            const int P = br, eP = be, ppos = bp;
            // ---- row swap (:19-27): positions only ----
            if (lane == 0) {
                const u16 rk = in.piv[k];
                in.piv[k] = (u16)P;
                in.piv[ppos] = rk;
                in.pos[P] = (u16)k;
                in.pos[rk] = (u16)ppos;
                in.pe[k] = (u16)eP;
            }
            const double alpha = 1.0 / in.sv[eP]; // (:31)
            const double yk = y[P];
            // ---- pivot row: entries with column > k -> (tcol, tval) ----
            int np = 0;
            {
                const int q0 = st.rptr[P], q1 = st.rptr[P + 1];
                for (int tb = q0; tb < q1; tb += 32) {
                    const int t = tb + lane;
                    const bool take = t < q1 && (int)st.rcol[t] > k;
                    const unsigned bal = __ballot_sync(0xffffffffu, take);
                    if (take) {
                        const int idx = np + __popc(bal & ((1u << lane) - 1u));
                        in.tcol[idx] = st.rcol[t];
                        in.tval[idx] = in.sv[st.rent[t]];
                    }
                    np += __popc(bal);
                }
                for (u16 f = in.rhead[P]; f != SP_NONE; f = in.fnr[f]) {
                    if ((int)in.fc[f] > k) {
                        if (lane == 0) {
                            in.tcol[np] = in.fc[f];
                            in.tval[np] = in.sv[nnz + f];
                        }
                        ++np;
                    }
                }
            }
            __syncwarp();
            // ---- eliminate column k from the rows below (:29-42), one row at a time, lanes over the pivot row ----
            const int nstat = c1 - c0;
            u16 f = in.chead[k];
            for (int i = 0;; ++i) {
                int r, e;
                if (i < nstat) {
                    e = c0 + i;
                    r = st.erow[e];
                } else {
                    if (f == SP_NONE) break;
                    e = nnz + (int)f;
                    r = in.fr[f];
                    f = in.fnc[f];
                }
                if ((int)in.pos[r] <= k) continue; // the pivot row itself, or a finished row (entry of U)
                const double a = in.sv[e];
                const double m = alpha * a;
                __syncwarp();
                if (lane == 0) in.sv[e] = m;
                if (a == 0.0) continue; // a - 0*p == a
                if (lane == 0) y[r] = y[r] - m * yk; // (:40)
                flops += 3 + 2 * np;
                for (int jb = 0; jb < np; jb += 32) {
                    const int ji = jb + lane;
                    bool need = false;
                    int j = 0;
                    double t = 0.0;
                    if (ji < np) {
                        j = in.tcol[ji];
                        t = m * in.tval[ji];
                        const int e2 = lookup(st, in, r, j);
                        if (e2 >= 0)
                            in.sv[e2] = in.sv[e2] - t; // (:35-39)
                        else
                            need = true;
                    }
                    unsigned fb = __ballot_sync(0xffffffffu, need);
                    if (fb) { // run-time fill-in: the dense entry was +0.0
                        if (nfill + __popc(fb) > in.fcap) return 2;
                        const int mine = nfill + __popc(fb & ((1u << lane) - 1u));
                        if (need) {
                            in.fr[mine] = (u16)r;
                            in.fc[mine] = (u16)j;
                            in.sv[nnz + mine] = 0.0 - t;
                        }
                        __syncwarp();
                        if (lane == 0) { // chain insertions in order (they share the row chain)
                            for (unsigned bb = fb; bb; bb &= bb - 1u) {
                                const int fi = nfill + __popc(fb & ((1u << (__ffs((int)bb) - 1)) - 1u));
                                const int cj = in.fc[fi];
                                in.fnr[fi] = in.rhead[r];
                                in.rhead[r] = (u16)fi;
                                in.fnc[fi] = in.chead[cj];
                                in.chead[cj] = (u16)fi;
                                in.sig[r] |= 1u << (cj & 31);
                            }
                        }
                        nfill += __popc(fb);
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
        }
        flops += n; // reciprocals
        nfill_out = nfill;
        return 0;
    }

    // a new right-hand side through the stored multipliers
    static __device__ void forward(const Team &tm, const SpStatic &st, const SpInst &in, double *__restrict__ y)
    {
        const int n = st.n, nnz = st.nnz, lane = tm.rank;
        for (int k = 0; k < n - 1; ++k) {
            const double yk = y[in.piv[k]];
            const int c0 = st.cptr[k], c1 = st.cptr[k + 1];
            __syncwarp();
            for (int t = c0 + lane; t < c1; t += 32) {
                const int r = st.erow[t];
                if ((int)in.pos[r] > k) {
                    const double m = in.sv[t];
                    if (m != 0.0) y[r] = y[r] - m * yk;
                }
            }
            int i = 0;
            for (u16 f = in.chead[k]; f != SP_NONE; f = in.fnc[f], ++i) {
                if ((i & 31) == lane) {
                    const int r = in.fr[f];
                    if ((int)in.pos[r] > k) {
                        const double m = in.sv[nnz + f];
                        if (m != 0.0) y[r] = y[r] - m * yk;
                    }
                }
            }
            __syncwarp();
        }
    }

    // gauss_lu.rs:44-53; false when a solution component is not finite
    static __device__ bool backward(const Team &tm, const SpStatic &st, const SpInst &in, double *__restrict__ y,
                                    double *__restrict__ xp)
    {
        const int n = st.n, nnz = st.nnz, lane = tm.rank;
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        bool ok = true;
        for (int i = n - 1; i >= 0; --i) {
            if (WIN && in.windowed) {
                // Window of up to 32 columns (descending).  Lane l takes column i - l: its solution component is
                // y[row] / pivot as soon as no later column of the window holds an entry in its pivot row (then nothing
                // the window does writes into that y).  All divisions of the window run in parallel; the few columns
                // that hold entries of U are then applied in descending order, as the sequential loop would.
                const int nw = min(32, i + 1), ic = i - lane;
                bool valid = lane < nw;
                int ri = 0, nU = 0;
                if (valid) {
                    ri = in.piv[ic];
                    for (int t = st.rptr[ri]; t < st.rptr[ri + 1] && valid; ++t) {
                        const int j = st.rcol[t];
                        if (j > ic && j <= i) valid = false;
                    }
                    for (u16 f = in.rhead[ri]; f != SP_NONE && valid; f = in.fnr[f]) {
                        const int j = in.fc[f];
                        if (j > ic && j <= i) valid = false;
                    }
                    for (int t = st.cptr[ic]; t < st.cptr[ic + 1]; ++t) nU += (int)in.pos[st.erow[t]] < ic ? 1 : 0;
                    for (u16 f = in.chead[ic]; f != SP_NONE; f = in.fnc[f]) nU += (int)in.pos[in.fr[f]] < ic ? 1 : 0;
                }
                const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                int ext = __ffs((int)~vmask) - 1;
                if (ext < 0) ext = 32;
                ext = min(ext, nw);
                if (ext >= 2) {
                    // a column with entries of U changes y of rows that later lanes of the window read only if that
                    // row is one of their pivot rows: excluded above for entries INSIDE the window; a U entry of
                    // column ic sits in a row pivoted before ic, i.e. the pivot row of a LATER lane -> cut there too
                    unsigned ucols = __ballot_sync(0xffffffffu, lane < ext && nU > 0);
                    // rows touched by column ic's U entries have positions < ic: lanes > l.  Keep only the first such
                    // column and end the window right after it.
                    if (ucols) ext = min(ext, __ffs((int)ucols));
                    double xi = 0.0;
                    if (lane < ext) {
                        xi = y[ri] / in.sv[in.pe[ic]];
                        xp[ic] = xi;
                    }
                    ok = ok && __all_sync(0xffffffffu, lane >= ext || fabs(xi) < INF);
                    __syncwarp();
                    ucols &= ext >= 32 ? 0xffffffffu : ((1u << ext) - 1u);
                    if (ucols) { // at most one: the last column of the window
                        const int l = __ffs((int)ucols) - 1, iu = i - l;
                        const double xu = __shfl_sync(0xffffffffu, xi, l);
                        const int c0 = st.cptr[iu], c1 = st.cptr[iu + 1];
                        for (int t = c0 + lane; t < c1; t += 32) {
                            const int r = st.erow[t];
                            if ((int)in.pos[r] < iu) y[r] = y[r] - xu * in.sv[t];
                        }
                        int q = 0;
                        for (u16 f = in.chead[iu]; f != SP_NONE; f = in.fnc[f], ++q) {
                            if ((q & 31) == lane) {
                                const int r = in.fr[f];
                                if ((int)in.pos[r] < iu) y[r] = y[r] - xu * in.sv[nnz + f];
                            }
                        }
                        __syncwarp();
                    }
                    i -= ext - 1;
                    continue;
                }
            }
            const int ri = in.piv[i];
            const double xi = y[ri] / in.sv[in.pe[i]];
            ok = ok && (fabs(xi) < INF);
            if (lane == 0) xp[i] = xi;
            const int c0 = st.cptr[i], c1 = st.cptr[i + 1];
            __syncwarp();
            for (int t = c0 + lane; t < c1; t += 32) {
                const int r = st.erow[t];
                if ((int)in.pos[r] < i) y[r] = y[r] - xi * in.sv[t];
            }
            int q = 0;
            for (u16 f = in.chead[i]; f != SP_NONE; f = in.fnc[f], ++q) {
                if ((q & 31) == lane) {
                    const int r = in.fr[f];
                    if ((int)in.pos[r] < i) y[r] = y[r] - xi * in.sv[nnz + f];
                }
            }
            __syncwarp();
        }
        return ok;
    }

    // f = A x + H g - b and its class norms (mna.rs:25-29, node_vec_norm.rs:12-33); b comes from the copy the
    // right-hand-side assembly of this Newton call left in w.b
    template <class W, class V>
    static __device__ __forceinline__ void resid(const Team &tm, const ModeProg &M, const W &w, V x, int nbad, double &nv,
                                                 double &ni)
    {
        const int n = M.n;
        double sv = 0.0, si = 0.0;
        for (int row = tm.rank; row < n; row += 32) {
            double ax = 0.0;
            {
                const int c0 = __ldg(&M.Ax.ptr[row]), c1 = __ldg(&M.Ax.ptr[row + 1]);
                for (int c = c0; c < c1; ++c) {
                    const unsigned wd = (unsigned)__ldg(&M.Ax.con[c]);
                    const double v = w.val[wd & AX_SLOT_MASK];
                    const double xc = x[(wd & ~CON_NEG) >> AX_COL_SHIFT];
                    ax = ax + ((wd & CON_NEG) ? -v : v) * xc; // separate multiply and add, like the reference's dot (mna.rs:28)
                }
            }
            double hg = 0.0;
            if (M.n_gfun > 0) {
                int own_bad = 0;
                const int c0 = __ldg(&M.f_nl.ptr[row]), c1 = __ldg(&M.f_nl.ptr[row + 1]);
                for (int c = c0; c < c1; ++c) {
                    const unsigned wd = (unsigned)__ldg(&M.f_nl.con[c]);
                    const double v = w.val[wd & ~CON_NEG];
                    own_bad += isfinite(v) ? 0 : 1;
                    hg += (wd & CON_NEG) ? -v : v;
                }
                if (nbad > own_bad) hg = __longlong_as_double(0x7ff8000000000000LL);
            }
            const double f = ax + hg - w.b[row];
            if (__ldg(&M.cls[row]))
                si += f * f;
            else
                sv += f * f;
        }
        tm.sum2(sv, si);
        nv = sqrt(sv) / (double)M.nv;
        ni = sqrt(si) / (double)M.ni;
    }

    // damped Newton (newtons_method.rs:19-71); same structure as WarpSolver::newton
    template <class W>
    static __device__ __forceinline__ int newton(const Team &tm, const ModeProg &M, const W &w, State &s, Counters &cn,
                                                 int skip)
    {
        const int n = M.n;
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        const bool linear = (M.n_nl == 0);
        SpStatic st;
        SpInst in;
        views(M, w, st, in);
        double *__restrict__ y = w.r.p;
        double *__restrict__ xp = w.xp.p;

        int flags = eval_nl(tm, M, w.x, w.val);
        if (flags >> 20) return -ST_NMOS_UNREACHABLE;
        double sov = INF, soi = INF, eov = INF, eoi = INF;
        double ev = INF, ei = INF, sv = INF, si = INF;
        bool xp_ready = false;
        int it = 0;
        while (it < 100 && !(sv < 0.001 * sov + 1e-6 && si < 0.001 * soi + 1e-9 && ev < 0.001 * eov + 1e-6 &&
                             ei < 0.001 * eoi + 1e-9)) {
            const bool do_factor = !(linear && skip && s.jvalid);
            const bool do_solve = !(linear && skip && xp_ready);
            if (do_solve) {
                for (int row = tm.rank; row < n; row += 32) { // r = b + nonlinear rhs, reference accumulation order
                    double acc = 0.0;
                    int c0 = __ldg(&M.b_all.ptr[row]), c1 = __ldg(&M.b_all.ptr[row + 1]);
                    for (int c = c0; c < c1; ++c) {
                        const unsigned wd = (unsigned)__ldg(&M.b_all.con[c]);
                        const double v = w.val[wd & ~CON_NEG];
                        acc += (wd & CON_NEG) ? -v : v;
                    }
                    w.b[row] = acc; // b of this solve: sources do not change inside a Newton call
                    c0 = __ldg(&M.r_nl.ptr[row]);
                    c1 = __ldg(&M.r_nl.ptr[row + 1]);
                    for (int c = c0; c < c1; ++c) {
                        const unsigned wd = (unsigned)__ldg(&M.r_nl.con[c]);
                        const double v = w.val[wd & ~CON_NEG];
                        acc += (wd & CON_NEG) ? -v : v;
                    }
                    y[row] = acc;
                }
                if (do_factor) {
                    // values of the static entries from the slot table (predicted fill-in: no contribution -> 0.0)
                    const Gather &g = M.sp.S;
                    for (int d = tm.rank; d < g.n_dst; d += 32) {
                        double acc = 0.0;
                        const int c0 = __ldg(&g.ptr[d]), c1 = __ldg(&g.ptr[d + 1]);
                        for (int c = c0; c < c1; ++c) {
                            const unsigned wd = (unsigned)__ldg(&g.con[c]);
                            const double v = w.val[wd & ~CON_NEG];
                            acc += (wd & CON_NEG) ? -v : v;
                        }
                        in.sv[d] = acc;
                    }
                    __syncwarp();
                    s.jvalid = false;
                    if (const int why = factor(tm, st, in, y, cn.c[CN_FLOPS], s.nfill)) {
                        cn.c[CN_REJ_NEWTON] = 0x100000000ull | (unsigned)why; // diagnostic: picked up by the kernel
                        return -ST_REDO;
                    }
                    s.jvalid = linear;
                    cn.c[CN_FACTOR]++;
                } else {
                    __syncwarp();
                    forward(tm, st, in, y);
                    cn.c[CN_FLOPS] += (unsigned long long)(st.nnz + s.nfill); // upper bound: 2 per entry of L, L ~ half
                }
                if (!backward(tm, st, in, y, xp)) {
                    cn.c[CN_REJ_NEWTON] = 0x100000000ull | 3u;
                    return -ST_REDO;
                }
                xp_ready = true;
                cn.c[CN_SOLVE]++;
                cn.c[CN_FLOPS] += (unsigned long long)(st.nnz + s.nfill + n); // back substitution: 2 per entry of U + n divisions
            }
            double qv = 0.0, qi = 0.0;
            for (int i = tm.rank; i < n; i += 32) { // damped step (newtons_method.rs:73-77)
                const double xi = w.x[i];
                const double d = xp[i] - xi;
                const double t = (1.3 / 16.0) * copysign(1.0, d) * damp_log1p(16.0 * fabs(d));
                w.xn[i] = xi + t;
                if (__ldg(&M.cls[i]))
                    qi += t * t;
                else
                    qv += t * t;
            }
            tm.sum2(qv, qi);
            sv = sqrt(qv) / (double)M.nv;
            si = sqrt(qi) / (double)M.ni;
            tm.sync();

            flags = eval_nl(tm, M, w.xn, w.val);
            if (flags >> 20) return -ST_NMOS_UNREACHABLE;
            resid(tm, M, w, w.xn, flags & 0xfffff, ev, ei);
            if (isinf(ev) || isinf(ei)) return -ST_DIVERGED;
            for (int i = tm.rank; i < n; i += 32) w.x[i] = w.xn[i];
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
};

// one warp per instance; the warps of a block share the u16 copy of the static structure
// GLOB: 0 = the whole working set is in shared memory (every pointer provably shared: LDS / STS, no generic accesses);
// 1 = the per-step history (C/L state, ring of accepted x) lives in global scratch; 2 = so may the slot table and the
// solution vectors (large circuits)
template <int AN, bool WIN, int GLOB>
// (no min-blocks argument: nvcc then settles on 128 registers, 16 warps per SM at n = 65; asking for 1 block gives 174
// registers / 8 warps and 5 blocks gives 96 registers with more spills — both measured slower)
__global__ void __launch_bounds__(128) k_spw(const __grid_constant__ RunArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevProg &P = a.prog;
    const ModeProg &M = AN == AN_OP ? P.op : AN == AN_DC ? P.dc : P.tran;
    const int n = M.n, nnz = M.sp.nnz;
    {
        u16 *sb = reinterpret_cast<u16 *>(smem_raw);
        for (int i = threadIdx.x; i <= n; i += blockDim.x) {
            sb[i] = (u16)__ldg(&M.sp.cptr[i]);
            sb[n + 1 + i] = (u16)__ldg(&M.sp.rptr[i]);
        }
        u16 *rl = sb + 2 * (n + 1);
        for (int i = threadIdx.x; i < n; i += blockDim.x) rl[i] = (u16)__ldg(&M.sp.rlong[i]); // -1 -> SP_NONE
        u16 *e0 = rl + n;
        for (int i = threadIdx.x; i < nnz; i += blockDim.x) {
            e0[i] = (u16)__ldg(&M.sp.erow[i]);
            e0[nnz + i] = (u16)__ldg(&M.sp.rcol[i]);
            e0[2 * nnz + i] = (u16)__ldg(&M.sp.rent[i]);
        }
        u16 *lm = e0 + 3 * nnz;
        for (int i = threadIdx.x; i < M.sp.n_long * n; i += blockDim.x) lm[i] = (u16)__ldg(&M.sp.lmap[i]);
    }
    __syncthreads();
    const int fcap = M.sp.fcap;
    const SpLayout T = sp_layout(n, nnz, fcap, P.n_slots, P.n_dyn, P.n_src, AN == AN_TRAN, M.sp.hist_global);
    unsigned char *wb = smem_raw + sp_static_bytes(n, nnz, M.sp.n_long) + (size_t)(threadIdx.x >> 5) * ((T.bytes + 15) & ~(size_t)15);
    double *base = reinterpret_cast<double *>(wb);
    Ws w;
    w.ld = n; // stride of the history ring
    w.J.p = base + T.osv;
    w.r.p = base + T.oy;
    w.prev.p = AN == AN_TRAN ? base + T.oprev : nullptr;
    {
        // per-step history (C/L state, last three accepted x): shared memory, or — large circuits — a global
        // scratch row per resident warp (touched once per accepted step, coalesced)
        if constexpr (GLOB == 1) {
            double *gb = a.scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * T.hist_doubles;
            w.state.p = gb + T.ostate;
            w.ring.p = gb + T.oring;
            w.val.p = base + T.oval;
            w.x.p = base + T.ox;
            w.xn.p = base + T.oxn;
            w.xp.p = base + T.oxp;
            w.b.p = base + T.obv;
        } else if constexpr (GLOB == 2) {
            double *gb = a.scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * T.hist_doubles;
            const bool hist_g = AN == AN_TRAN && (M.sp.hist_global & 1);
            w.state.p = (hist_g ? gb : base) + T.ostate;
            w.ring.p = (hist_g ? gb : base) + T.oring;
            w.val.p = (T.val_global ? gb : base) + T.oval; // large slot tables: L2-resident scratch (more warps per SM)
            double *vb = T.vec_global ? gb : base;
            w.x.p = vb + T.ox;
            w.xn.p = vb + T.oxn;
            w.xp.p = vb + T.oxp;
            w.b.p = vb + T.obv;
        } else {
            w.state.p = base + T.ostate;
            w.ring.p = base + T.oring;
            w.val.p = base + T.oval;
            w.x.p = base + T.ox;
            w.xn.p = base + T.oxn;
            w.xp.p = base + T.oxp;
            w.b.p = base + T.obv;
        }
        w.A.p = base + T.otval; // (the dense A is not kept by this solver: the field carries the pivot-row scratch)
    }
    w.piv.p = reinterpret_cast<int *>(base + T.doubles);
    w.aux = reinterpret_cast<const int *>(smem_raw);
    WarpTeam tm;
    Counters cn;
    for (;;) {
        long long inst = 0;
        if (tm.rank == 0) inst = next_instance(a);
        inst = __shfl_sync(0xffffffffu, inst, 0);
        if (inst < 0) break;
        Counters ci;
        run_instance<WarpTeam, SparseWarpSolverT<WIN>, AN>(tm, a, w, inst, ci);
        tm.sync();
        int st = 0;
        if (tm.rank == 0) st = a.status[inst];
        st = __shfl_sync(0xffffffffu, st, 0);
        if (st == ST_REDO) { // hand the instance to the dense kernels
            if (tm.rank == 0) {
                a.redo_list[atomicAdd(a.redo_count, 1ULL)] = inst;
                // why (diagnostic): 1 pivot, 2 fill pool, 3 non-finite solution; 20 bits per reason
                const int why = (int)(ci.c[CN_REJ_NEWTON] & 3ull);
                atomicAdd(&a.counters[CN_WHY], 1ULL << (20 * (why > 0 ? why - 1 : 0)));
            }
        } else {
            for (int i = 0; i < CN_LOCAL; ++i) cn.c[i] += ci.c[i];
        }
    }
    flush_counters(tm, a, cn);
}

KernelFn sparse_kernel(int an, bool windowed, int global_scratch); // kern_sparse.cu

} // namespace ftsb
