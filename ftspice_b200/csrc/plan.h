// plan.h — learned elimination plans for small circuits (solver_plan.cuh).
//
// Partial pivoting is value dependent, but over a Monte-Carlo batch / a Newton run the pivot sequence takes very few
// values (C2: 6 distinct sequences in 56 781 factorisations, the top three cover 99.6 %).  A *plan* is one pivot
// sequence with everything that follows from it worked out on the host: per elimination step the structural
// candidates of the pivot search, the multipliers, the list of `sv[d] -= sv[m] * sv[s]` updates on a compact entry
// numbering (static pattern + the fill-in of every plan), and the substitution lists.  The kernel interprets a plan
// and checks at every step that the planned pivot row IS the row the reference's rule picks (first position with the
// strictly largest |a| among the structural candidates); if another row wins it switches to the plan that shares
// the pivots so far and takes that row next, or — unknown sequence — gives the instance to the dense kernel.
#pragma once
#include <cuda_runtime.h> // int4

#include <cstdint>
#include <utility>
#include <vector>

namespace ftsb {

constexpr int PLAN_MAX = 8;         // plans kept per analysis mode
constexpr int PLAN_TRACE_MAX = 128; // pivot sequences recorded per pilot instance (ring)
constexpr int PLAN_PILOT = 1024;    // pilot instances, spread over the batch
constexpr int PLAN_TRACE_ROWS = 4096; // rows of the trace buffer (pilot sample / traced redo instances)

// device view (pointers into one int buffer)
struct PlanProg {
    int K, S, nnz0, n;
    const int4 *rec;         // [K][n] packed step records (see solver_plan.cuh)
    const int *cand;         // e | row << 8 | pos << 16
    const int *mult;         // e | row << 8
    const int *upd;          // d | m << 8 | s << 16
    const int *bu;           // e | row << 8      (entries of U in column i, for the back substitution)
    const signed char *next; // [K][n][16] plan that shares the first k pivots and takes row w at step k, or -1
    const int *jdst;         // [nnz0] compact entry index of J_full's d-th destination
};

struct PlanHost {
    int K = 0, S = 0, nnz0 = 0, n = 0;
    std::vector<int> ibuf; // rec (16-byte aligned at offset 0), then the packed lists
    int off_rec = 0, off_cand = 0, off_mult = 0, off_upd = 0, off_bu = 0, off_next = 0, off_jdst = 0;
    double coverage = 0.0; // share of the pilot's factorisations whose pivot sequence has a plan
    long long flops_factor = 0, flops_solve = 0; // floating-point operations of one factorisation / one back substitution (plan 0)
};

// jfull_dst: J_full destinations (row | col << 16) in J_full order; perms: (nibble-packed position -> row, count),
// any order.  Returns false when no usable plan exists (n < 2, n > 16, too many entries, inconsistent traces).
bool build_plans(int n, const std::vector<int> &jfull_dst, std::vector<std::pair<uint64_t, long long>> perms, PlanHost &out,
                 int max_plans = PLAN_MAX);

} // namespace ftsb
