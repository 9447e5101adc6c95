// compile.h — host-side result of compiling a topology (offsets into one int buffer; see prog.h)
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "prog.h"

namespace ftsb {

struct ElemDesc { // mirrors ftsb_elem of include/ftsb.h
    int kind;
    int node[3];
    int branch;
    int fn_kind;
};

struct GatherOff {
    int dst = 0, ptr = 0, con = 0, n_dst = 0, n_con = 0;
};

struct SparseOff {
    int nnz = 0, n_pred = 0, cptr = 0, erow = 0, rptr = 0, rcol = 0, rent = 0, n_long = 0, rlong = 0, lmap = 0, window_hint = 0;
    int piv = 0;       // [n] predicted pivot row of every elimination step (the guess the static structure was built with)
    int piv_valid = 0; // every column had a structural pivot candidate
    GatherOff S;
};

struct ModeOff {
    int n = 0, nv = 0, ni = 0, n_nl = 0, n_gfun = 0;
    int cls = 0, nl = 0;
    GatherOff A_lin, A_dyn, J_nl, b_all, r_nl, f_nl, J_full, Ax;
    SparseOff sp;
};

struct Compiled {
    struct Impl;
    std::shared_ptr<Impl> impl;
    int n_elems = 0, n_run = 0, n_start = 0, n_slots = 0, n_fn = 0, n_res = 0, n_src = 0, n_dyn = 0;
    int off_res = 0, off_src = 0, off_dyn = 0, off_src_of_elem = 0;
    ModeOff mode[3]; // op, dc, tran
    Layout lay;
    std::vector<int> slot_of_elem, src_of_elem, kinds;
    const std::vector<int> *ibuf = nullptr;
    const std::vector<int> &ints() const;
};

bool compile_circuit(const ElemDesc *elems, int n_elems, int n_run, int n_start, const int *row_class, int ld_min,
                     Compiled &out, std::string &err);

} // namespace ftsb
