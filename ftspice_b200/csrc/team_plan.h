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
    const int *spre;        // [nJ] (count << 16 | index): the leading `count` contributions of the entry are linear / companion stamps
                            //      (they do not change inside a Newton call); their partial sum is kept per attempt at svp[index] and
                            //      the per-iteration gather starts behind them.  0: no kept sum (fewer than two such contributions)
    int nPre;               // kept sums
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
    // ---- residual in ELL form: row r holds its first ellW (3 or 4) entries at A slots r * ellW + w (padding: value 0.0,
    //      column r); row n is a ZERO row (lanes dealt a row >= n clamp to it); the entries of longer rows follow at slots
    //      ellW * (n + 1) + k (ovf lists).  aptr / acon are indexed by slot.
    int ellW, nSlotA;       // slots of A: ellW * (n + 1) + overflow entries
    const int *ellcol;      // [n + 1][2] columns of the row's ELL entries, 16 bits each
    const int *ovfptr;      // [n + 1]
    const int *ovfcol;      // [n_ovf]
    int longRow;            // the row with the most overflow entries if it has at least 16 (a supply rail), else -1
    // ---- substitutions as single-lane op streams (one lane of a team walks them; the values they need are staged
    //      through shared memory in chunks of TEAM_CHUNK by all lanes of the team).  Value indices address the factors
    //      array followed by the reciprocal pivots (index S + k).
    //      forward op : x = m entry value slot;  word0 = row r | pivot row << 16
    //      backward op: word1 < 0: divide — word0 = pivot row | column i << 16, two value slots (pivot, its reciprocal);
    //                   else update — word0 = row r, one value slot (entry of U); uses the x_i of the last divide
    int chain, nFop, nBop, nFchunk, nBchunk; // chain = 1: every elimination step has few targets, use the streams
    const int *fop;         // [nFop] word0
    const int *fvidx;       // [nFchunk * TEAM_CHUNK] value index of slot (padding: 0)
    const int *fchunk;      // [nFchunk + 1] ops of chunk c: [fchunk[c], fchunk[c + 1]); their value slots are consecutive from
                            //              the start of the chunk
    const int *bop;         // [nBop][2]
    const int *bvidx;       // [nBchunk * TEAM_CHUNK]
    const int *bchunk;      // [nBchunk + 1]
    // ---- direct streams, for factors that live in shared memory (nonlinear circuits: a factorisation per Newton
    //      iteration).  An elimination step whose pivot and column entries no earlier update writes is EARLY: its pivot
    //      check, reciprocal and multipliers are done for all such steps at once, lanes in parallel (an inverter chain:
    //      every stage); what is left — the other steps' checks / multipliers inline, and every `sv[d] -= sv[m] * sv[s]`
    //      update in elimination order (the rail column of a chain is a true recurrence) — is one op stream walked by one
    //      lane.  xop: word1 < 0: inline step, word0 = k; else update, word0 = dst | src << 16, word1 = multiplier entry.
    //      dfop: word0 = row r | pivot row << 16, word1 = multiplier entry.  dbop: word1 < 0: divide, word0 = pivot row |
    //      column i << 16, word1 & 0xffff = pivot entry; else update, word0 = row r, word1 = entry of U.
    int nEarly, nXop, nDFop, nDBop;
    const int *early;       // [nEarly] step indices
    const int *xop;         // [nXop][2]
    const int *dfop;        // [nDFop][2]
    const int *dbop;        // [nDBop][2]
    // ---- level schedules of the same three op sets (factors in shared memory): the ops are sorted into LEVELS such that an
    //      op's operands are produced in earlier levels and nothing it overwrites is read later than its own level by another
    //      op (read-after-write, write-after-write and write-after-read on the entries of the factors / y / x); the ops of a
    //      level are dealt to the lanes of the team, a warp barrier separates levels.  Every value sees the same operations in
    //      the same order as in the sequential streams, so the results are the same bits.  An inverter chain: 765 elimination
    //      ops in 256 levels (all 257 early heads at once, then the rail recurrence), forward 509 ops in 256 levels, backward
    //      512 ops in THREE levels (x_rail, all updates, all divides) instead of one lane walking 1 786 dependent ops.
    //      xlop: as xop, plus the FORWARD ops (word1 bit 30 set: word0 = row r | pivot row << 16, word1 & 0xffff = multiplier
    //      entry) — the right-hand side is eliminated along with the matrix; flop: a separate forward schedule (empty).  blop: divide as dbop; update: word0 = row r | column i << 16 (x_i is read from
    //      x_proposed), word1 = entry of U.
    int levels;             // 1: use them (option team_levels)
    int nXlev, nFlev, nBlev;
    const int *xlptr, *xlop; // [nXlev + 1], [nXop][2] sorted by level
    const int *flptr, *flop;
    const int *blptr, *blop;
    // ---- the same schedules as the kernel walks them: segments (count, offset) — count > 0: a run of `count` NARROW levels (at
    //      most TEAM_NARROW ops each) stored padded to TEAM_NARROW op slots per level (padding: word1 = TEAM_OP_NOP), so that a
    //      lane's op word of level j sits at offset + j * TEAM_NARROW + lane and is fetched one level ahead without pointer
    //      chasing; count < 0: one WIDE level of -count ops.
    int nXseg, nBseg;
    const int *xseg, *xsop; // [nXseg][2], [..][2]
    const int *bseg, *bsop;
};
constexpr int TEAM_NARROW = 4;
constexpr int TEAM_OP_NOP = 0x20000000, TEAM_OP_FWD = 0x40000000;

constexpr int TEAM_CHUNK = 16; // value slots per chunk: two chunks of staging per instance (256 bytes; measured on C4: 8 slots 68.4, 16 slots 72.1, 32 slots 58.2 M solve-points/s); teams of more lanes use the lane-parallel substitutions

struct TeamPlanHost {
    int n = 0, nA = 0, S = 0, nJ = 0, nT = 0, nU = 0, nB = 0, nPre = 0;
    std::vector<int> ibuf;
    int o_arow = 0, o_acol = 0, o_aptr = 0, o_acon = 0, o_sptr = 0, o_scon = 0, o_spre = 0, o_pivrow = 0, o_pive = 0, o_tptr = 0, o_tgt = 0,
        o_tgtr = 0, o_uptr = 0, o_updm = 0, o_updds = 0, o_bptr = 0, o_bwde = 0, o_bwdr = 0;
    int ellW = 0, nSlotA = 0, o_ellcol = 0, o_ovfptr = 0, o_ovfcol = 0, longRow = -1;
    int chain = 0, nFop = 0, nBop = 0, nFchunk = 0, nBchunk = 0, o_fop = 0, o_fvidx = 0, o_fchunk = 0, o_bop = 0, o_bvidx = 0,
        o_bchunk = 0;
    int nEarly = 0, nXop = 0, nDFop = 0, nDBop = 0, o_early = 0, o_xop = 0, o_dfop = 0, o_dbop = 0;
    int nXseg = 0, nBseg = 0, o_xseg = 0, o_xsop = 0, o_bseg = 0, o_bsop = 0;
    int nXlev = 0, nFlev = 0, nBlev = 0, o_xlptr = 0, o_xlop = 0, o_flptr = 0, o_flop = 0, o_blptr = 0, o_blop = 0;
    long long flops_factor = 0, flops_solve = 0; // floating-point operations of one factorisation / one pair of substitutions
    TeamPlan view(const int *base) const;
};

// pivseq[k] = pivot row of step k.  False when the sequence is not a valid elimination order of the pattern or an index
// does not fit the packed lists.
bool build_team_plan(const Compiled &comp, int an, const std::vector<int> &pivseq, TeamPlanHost &out);

} // namespace ftsb
