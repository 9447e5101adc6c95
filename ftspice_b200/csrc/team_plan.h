// team_plan.h — static elimination plan of the per-team kernels (solver_team.cuh).
//
// For mid-size circuits (n > 16) the pivot sequence of the reference's partial pivoting (gauss_lu.rs:6-27) is, in
// practice, one fixed sequence per topology: companion conductances make the node rows diagonally dominant and the
// +-1 incidence entries of V-source rows beat every conductance.  Given that sequence everything else is static:
// the fill-in, the targets of every elimination step, the `sv[d] -= m * sv[s]` updates, the substitution lists.
// The host works them out once per (handle, analysis mode); the kernel walks the lists and CHECKS at every step that
// the planned row is the row the reference's rule picks (first position with the strictly largest |a|); an instance
// for which it is not — or that meets a zero / non-finite pivot candidate or solution component — is handed to the
// structure-exploiting warp kernel (solver_sparse.cuh), which pivots dynamically, and from there to the dense kernels.
#pragma once
#include <vector>

#include "compile.h"

namespace ftsb {

// device view (pointers into one int buffer)
struct TeamPlan {
    int n, nA, S, nJ;       // unknowns, entries of A (linear + companion part), entries of the factors (nJ static + fill-in)
    const int *arow;        // [n + 1] entries of A's row r: [arow[r], arow[r + 1]) (columns ascending)
    const int *acol;        // [nA]
    const int *aptr;        // [nA + 1] contributions of A's entry (slot | sign words, reference accumulation order)
    const int *acon;
    const int *sptr;        // [nJ + 1] contributions of the static entries of J (all phases); fill-in entries start at 0.0
    const int *scon;
    const int *pivrow;      // [n] pivot row of step k
    const int *pive;        // [n] its entry in column k
    const int *tptr;        // [n + 1] targets (rows eliminated) of step k
    const int *tgt;         // [nT] multiplier entry (r, k) | strict << 30: strict = the row sits BEFORE the pivot row, so the
                            //      reference's scan keeps it unless the pivot is strictly larger
    const int *tgtr;        // [nT] row r
    const int *uptr;        // [nT + 1] updates of target t
    const int *updm;        // [nU] multiplier entry
    const int *updds;       // [nU] destination entry | source entry << 16
    const int *bptr;        // [n + 1] entries of U in column i (rows at final positions < i)
    const int *bwde;        // [nB] entry
    const int *bwdr;        // [nB] row
};

struct TeamPlanHost {
    int n = 0, nA = 0, S = 0, nJ = 0, nT = 0, nU = 0, nB = 0;
    std::vector<int> ibuf;
    int o_arow = 0, o_acol = 0, o_aptr = 0, o_acon = 0, o_sptr = 0, o_scon = 0, o_pivrow = 0, o_pive = 0, o_tptr = 0, o_tgt = 0,
        o_tgtr = 0, o_uptr = 0, o_updm = 0, o_updds = 0, o_bptr = 0, o_bwde = 0, o_bwdr = 0;
    long long flops_factor = 0, flops_solve = 0; // floating-point operations of one factorisation / one pair of substitutions
    TeamPlan view(const int *base) const;
};

// pivseq[k] = pivot row of step k.  False when the sequence is not a valid elimination order of the pattern or an index
// does not fit the packed lists.
bool build_team_plan(const Compiled &comp, int an, const std::vector<int> &pivseq, TeamPlanHost &out);

} // namespace ftsb
