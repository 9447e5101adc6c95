// prog.h — the compiled circuit "program" shared by the host compiler and the device engine.
//
// A topology is compiled once into (a) a slot table: every quantity a device stamps (a conductance, an
// equivalent current, a source value, a branch current) owns one f64 slot of a per-instance value vector;
// and (b) gather maps: for every matrix / right-hand-side / residual entry that receives stamps, the ordered
// list of (slot, sign) contributions, in netlist element order and, inside an element, in the order the
// reference's stamp functions touch the entry (src/device/*.rs).  One thread owns one destination entry and
// sums its list, so assembly needs no atomics and is deterministic.
#pragma once
#include <cstdint>

namespace ftsb {

enum : int { K_R = 0, K_C = 1, K_L = 2, K_V = 3, K_I = 4, K_D = 5, K_Q = 6, K_M = 7 };
enum : int { FN_NONE = 0, FN_SIN = 1, FN_PULSE = 2, FN_EXP = 3 };
enum : int { ST_OK = 0, ST_NOT_CONVERGED = 1, ST_DIVERGED = 2, ST_TRAN_H_UNDERFLOW = 3, ST_NMOS_UNREACHABLE = 4 };

// slot 0 always holds 1.0 (the +-1 incidence entries of V sources and start-up inductors)
constexpr int SLOT_ONE = 0;
// slots per element kind
constexpr int SLOTS_R = 1;   // g = 1/R
constexpr int SLOTS_SRC = 1; // V: value ; I: effective value
constexpr int SLOTS_DYN = 2; // C, L (.tran): g_eq, i_eq
// diode: g_eq, i_eq, i
enum : int { D_G = 0, D_IEQ = 1, D_I = 2, SLOTS_D = 3 };
// NPN (npn.rs:110-133): Jacobian entries named row-col, rhs, currents
enum : int {
    Q_CC = 0, Q_BB = 1, Q_EE = 2, Q_EC = 3, Q_CE = 4, Q_EB = 5, Q_BE = 6, Q_CB = 7, Q_BC = 8,
    Q_RC = 9, Q_RB = 10, Q_RE = 11, Q_IC = 12, Q_IB = 13, Q_IE = 14, SLOTS_Q = 15
};
// NMOS (nmos.rs:116-133): entries named by the NETLIST terminals (the run-time d/s swap is folded into the
// values), rhs, drain current
enum : int { M_DD = 0, M_DS = 1, M_DG = 2, M_SD = 3, M_SS = 4, M_SG = 5, M_RD = 6, M_RS = 7, M_ID = 8, SLOTS_M = 9 };

constexpr unsigned CON_NEG = 0x80000000u; // contribution word = slot | (negative ? CON_NEG : 0)

// CSR gather map over destinations
struct Gather {
    const int *dst; // [n_dst] target: a row (vectors) or row | col << 16 (matrices; the kernel applies its own ld)
    const int *ptr; // [n_dst + 1]
    const int *con; // contribution words
    int n_dst;
    int n_con;
};

// Static structure of the Jacobian for the structure-exploiting (sparse) solver: the structurally non-zero
// entries of J_full plus the fill-in a symbolic elimination predicts, in column-major order (entry index e:
// column ascending, then row ascending).  Run-time fill-in outside this set is handled dynamically per instance.
struct SparseProg {
    int nnz;         // static entries
    int n_pred;      // of which predicted fill-in (no stamp contributes: value 0.0 until an update touches it)
    int fcap;        // run-time fill-in pool entries per instance (set by the launcher)
    int hist_global; // .tran: C/L state and the history ring live in global scratch (set by the launcher)
    const int *cptr; // [n + 1] entries of column c: e in [cptr[c], cptr[c + 1])
    const int *erow; // [nnz]   row of entry e
    const int *rptr; // [n + 1] row-ordered view: the t-th entry of row r, t in [rptr[r], rptr[r + 1]), columns ascending
    const int *rcol; // [nnz]   its column
    const int *rent; // [nnz]   its entry index e
    int windowed;    // the launcher's switch for the windowed elimination (solver_sparse.cuh)
    int pad3;
    int n_long;      // rows with many entries get a direct column -> entry map (at most SP_MAX_LONG of them)
    int pad2;
    const int *rlong; // [n]  index of the row's map, or -1
    const int *lmap;  // [n_long][n] entry index of (row, col) or -1
    Gather S;        // sv[e] = sum of the stamp contributions of entry e (dst[e] == e)
};

constexpr int SP_MAX_LONG = 4;
constexpr int SP_LONG_ROW = 12; // a row is 'long' above this many static entries

// one analysis mode: OP (start-up numbering, inductors are shorts with a current row), DC (inductors and
// capacitors open), TRAN (trapezoidal companions)
struct ModeProg {
    int n;          // unknowns
    int nv, ni;     // rows per class (Voltage / Current)
    int n_nl;       // nonlinear devices
    int n_gfun;     // nonlinear functions m (diode 1, NPN 3, NMOS 3)
    int pad;
    const int *cls; // [n] row class
    const int *nl;  // [n_nl][5] kind, node0, node1, node2, slot base
    Gather A_lin;   // A[dst] = sum of linear stamps                       (R, V, start-up L)
    Gather A_dyn;   // A[dst] = linear + companion stamps, dst touched by C/L (TRAN only)
    Gather J_nl;    // J[dst] = A[dst] + nonlinear stamps
    Gather b_all;   // b[row] = linear + companion rhs                    (row-indexed: ptr has n+1 entries)
    Gather r_nl;    // r[row] = b[row] + nonlinear rhs                      (row-indexed)
    Gather f_nl;    // (H g)[row] = sum of device currents                  (row-indexed)
    // thread-per-instance kernels keep neither A nor b in memory:
    Gather J_full;  // J[dst] = linear + companion + nonlinear stamps, every structurally non-zero entry
    Gather Ax;      // row-indexed; con = slot | sign | col << AX_COL_SHIFT: (A x)[row] = sum of +-val[slot] * x[col]
    SparseProg sp;  // structure-exploiting solver (solver_sparse.cuh)
};
constexpr int AX_COL_SHIFT = 20;
constexpr unsigned AX_SLOT_MASK = (1u << AX_COL_SHIFT) - 1u;

// per-instance working-set layout (offsets in doubles from the instance base)
struct Layout {
    int ld;       // leading dimension of the column-major matrices
    int oA, oJ;   // matrices (may live in global scratch instead: see mat_in_smem)
    int ob, orr, ox, oxn, oxp, of; // vectors [ld]
    int oval;     // [n_slots]
    int ostate;   // [2 * n_dyn]  (u, i) per C/L
    int oring;    // [3][ld] last three accepted x
    int oprev;    // [3][n_src] source value at t=0 / at its last evaluation / at this attempt (I-source drift, Q9)
    int doubles;  // vector part size (everything except A, J)
    int mat_doubles; // ld * ld
    int opiv;     // int region: piv[ld]
    int ints;
    int mat_in_smem;
    int bytes;    // shared-memory bytes per instance
};

struct DevProg {
    int n_elems, n_slots, n_fn, n_res, n_src, n_dyn;
    const int *res; // [n_res][2]  elem, slot
    const int *src; // [n_src][5]  elem, slot, kind, fn_kind, fn_slot
    const int *dyn; // [n_dyn][6]  elem, slot, kind, vpos, vneg, branch(start-up row of L or -1)
    const int *src_of_elem; // [n_elems] index into src or -1
    ModeProg op, dc, tran;
    Layout lay;
};

} // namespace ftsb
