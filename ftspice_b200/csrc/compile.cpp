// compile.cpp — host-side circuit compiler: element list -> slot table + gather maps (see prog.h).
//
// The emission order inside each element follows the reference's stamp functions line by line, so that a
// destination's contribution list is in the order the reference would have accumulated it:
//   Res::linear_stamp            src/device/res.rs:34-50
//   Vdd::linear_stamp            src/device/vdd.rs:46-64
//   Idd::linear_stamp            src/device/idd.rs:46-57
//   Ind::linear_startup_stamp    src/device/ind.rs:70-92      (.op only; no stamp at all in .dc)
//   Cap/Ind::dynamic_stamp       src/device/cap.rs:68-105, src/device/ind.rs:118-155   (.tran)
//   Diode/NPN/NMOS::nonlinear_stamp + nonlinear_funcs   diode.rs:39-93, npn.rs:39-134, nmos.rs:39-134
#include "compile.h"

#include <algorithm>
#include <map>
#include <string>

namespace ftsb {

namespace {

enum Phase { LIN = 0, DYN = 1, NL = 2 };

struct Con {
    int dst;
    unsigned word;
    int phase;
};

struct ModeBuild {
    int n = 0;
    int ld = 0;
    std::vector<Con> a, b, f; // in emission order
    std::vector<int> nl;      // [n_nl][5]
    int n_gfun = 0;
    void A(int row, int col, int slot, bool neg, int phase)
    {
        if (row < 0 || col < 0) return;
        a.push_back({row | (col << 16), (unsigned)slot | (neg ? CON_NEG : 0u), phase}); // ld-independent
    }
    void B(int row, int slot, bool neg, int phase)
    {
        if (row < 0) return;
        b.push_back({row, (unsigned)slot | (neg ? CON_NEG : 0u), phase});
    }
    void F(int row, int slot, bool neg)
    {
        if (row < 0) return;
        f.push_back({row, (unsigned)slot | (neg ? CON_NEG : 0u), NL});
    }
};

// CSR over destinations; `want` selects destinations by the phases present, `take` the contributions kept
struct GatherBuild {
    std::vector<int> dst, ptr, con;
};

GatherBuild make_gather(const std::vector<Con> &cons, int need_phase, unsigned take_mask)
{
    // destinations in ascending order; contributions keep emission order, lower phases first
    std::map<int, std::vector<const Con *>> by_dst;
    for (const Con &c : cons) by_dst[c.dst].push_back(&c);
    GatherBuild g;
    g.ptr.push_back(0);
    for (auto &kv : by_dst) {
        bool has = false;
        for (const Con *c : kv.second) has |= (c->phase == need_phase);
        if (!has) continue;
        g.dst.push_back(kv.first);
        for (int ph = 0; ph < 3; ++ph) {
            if (!(take_mask & (1u << ph))) continue;
            for (const Con *c : kv.second)
                if (c->phase == ph) g.con.push_back((int)c->word);
        }
        g.ptr.push_back((int)g.con.size());
    }
    return g;
}

} // namespace

struct Compiled::Impl {
    std::vector<int> ibuf;
    int push(const std::vector<int> &v)
    {
        int off = (int)ibuf.size();
        ibuf.insert(ibuf.end(), v.begin(), v.end());
        if (ibuf.size() & 1) ibuf.push_back(0);
        return off;
    }
};

static void put_gather(Compiled::Impl &im, const GatherBuild &g, GatherOff &o)
{
    o.n_dst = (int)g.dst.size();
    o.n_con = (int)g.con.size();
    o.dst = im.push(g.dst);
    o.ptr = im.push(g.ptr);
    o.con = im.push(g.con);
}

bool compile_circuit(const ElemDesc *elems, int n_elems, int n_run, int n_start, const int *row_class, int ld_min,
                     Compiled &out, std::string &err)
{
    if (n_elems < 0 || n_run < 0 || n_start < n_run) {
        err = "ftsb_compile: need 0 <= n_run <= n_start";
        return false;
    }
    if (n_start > 0 && !row_class) {
        err = "ftsb_compile: row_class is NULL";
        return false;
    }
    out = Compiled();
    out.impl = std::make_shared<Compiled::Impl>();
    Compiled::Impl &im = *out.impl;
    out.n_elems = n_elems;
    out.n_run = n_run;
    out.n_start = n_start;

    const int ld = std::max(std::max(n_start, 1), ld_min); // register kernels pad the rows to lanes x rows-per-lane
    ModeBuild mb[3]; // 0 op, 1 dc, 2 tran
    mb[0].n = n_start;
    mb[1].n = n_run;
    mb[2].n = n_run;
    for (auto &m : mb) m.ld = ld;

    std::vector<int> res, src, dyn, src_of_elem(n_elems, -1), slot_of_elem(n_elems, -1);
    int n_slots = 1; // slot 0 == 1.0
    int n_fn = 0;

    for (int ei = 0; ei < n_elems; ++ei) {
        const ElemDesc &e = elems[ei];
        auto chk = [&](int idx, int lim, const char *what) {
            if (idx < -1 || idx >= lim) {
                err = "ftsb_compile: element " + std::to_string(ei) + ": " + what + " index " + std::to_string(idx) +
                      " out of range";
                return false;
            }
            return true;
        };
        const int nn = (e.kind == K_Q || e.kind == K_M) ? 3 : 2;
        for (int j = 0; j < nn; ++j)
            if (!chk(e.node[j], n_run, "node")) return false;
        switch (e.kind) {
        case K_R: {
            int s = n_slots;
            n_slots += SLOTS_R;
            slot_of_elem[ei] = s;
            res.push_back(ei);
            res.push_back(s);
            int vneg = e.node[0], vpos = e.node[1];
            for (auto &m : mb) {
                m.A(vneg, vneg, s, false, LIN);
                m.A(vpos, vpos, s, false, LIN);
                m.A(vpos, vneg, s, true, LIN);
                m.A(vneg, vpos, s, true, LIN);
            }
            break;
        }
        case K_V: {
            if (e.branch < 0 || !chk(e.branch, n_run, "branch")) {
                if (err.empty()) err = "ftsb_compile: V source " + std::to_string(ei) + " needs a branch row";
                return false;
            }
            int s = n_slots;
            n_slots += SLOTS_SRC;
            slot_of_elem[ei] = s;
            src_of_elem[ei] = (int)src.size() / 5;
            src.insert(src.end(), {ei, s, K_V, e.fn_kind, e.fn_kind != FN_NONE ? n_fn : -1});
            if (e.fn_kind != FN_NONE) ++n_fn;
            int vneg = e.node[0], vpos = e.node[1], is = e.branch;
            for (auto &m : mb) {
                m.B(is, s, false, LIN); // assignment: the only contribution of that row
                m.A(is, vpos, SLOT_ONE, false, LIN);
                m.A(vpos, is, SLOT_ONE, false, LIN);
                m.A(is, vneg, SLOT_ONE, true, LIN);
                m.A(vneg, is, SLOT_ONE, true, LIN);
            }
            break;
        }
        case K_I: {
            int s = n_slots;
            n_slots += SLOTS_SRC;
            slot_of_elem[ei] = s;
            src_of_elem[ei] = (int)src.size() / 5;
            src.insert(src.end(), {ei, s, K_I, e.fn_kind, e.fn_kind != FN_NONE ? n_fn : -1});
            if (e.fn_kind != FN_NONE) ++n_fn;
            int vneg = e.node[0], vpos = e.node[1];
            for (auto &m : mb) {
                m.B(vpos, s, false, LIN);
                m.B(vneg, s, true, LIN);
            }
            break;
        }
        case K_C:
        case K_L: {
            int s = n_slots;
            n_slots += SLOTS_DYN;
            slot_of_elem[ei] = s;
            int vpos, vneg;
            if (e.kind == K_C) {
                vneg = e.node[0];
                vpos = e.node[1];
            } else {
                vpos = e.node[0];
                vneg = e.node[1];
                if (e.branch < n_run || e.branch >= n_start) {
                    err = "ftsb_compile: inductor " + std::to_string(ei) + " needs a start-up row in [n_run, n_start)";
                    return false;
                }
                int is = e.branch; // start-up: a short with a current unknown
                mb[0].A(is, vpos, SLOT_ONE, false, LIN);
                mb[0].A(vpos, is, SLOT_ONE, false, LIN);
                mb[0].A(is, vneg, SLOT_ONE, true, LIN);
                mb[0].A(vneg, is, SLOT_ONE, true, LIN);
            }
            dyn.insert(dyn.end(), {ei, s, e.kind, vpos, vneg, e.kind == K_L ? e.branch : -1});
            ModeBuild &m = mb[2];
            m.A(vpos, vpos, s, false, DYN);
            m.B(vpos, s + 1, true, DYN);
            m.A(vneg, vneg, s, false, DYN);
            m.B(vneg, s + 1, false, DYN);
            m.A(vneg, vpos, s, true, DYN);
            m.A(vpos, vneg, s, true, DYN);
            break;
        }
        case K_D: {
            int s = n_slots;
            n_slots += SLOTS_D;
            slot_of_elem[ei] = s;
            int p = e.node[0], n = e.node[1];
            for (auto &m : mb) {
                m.nl.insert(m.nl.end(), {K_D, p, n, -1, s});
                m.n_gfun += 1;
                m.A(p, p, s + D_G, false, NL);
                m.B(p, s + D_IEQ, true, NL);
                m.A(n, n, s + D_G, false, NL);
                m.B(n, s + D_IEQ, false, NL);
                m.A(p, n, s + D_G, true, NL);
                m.A(n, p, s + D_G, true, NL);
                m.F(p, s + D_I, false);
                m.F(n, s + D_I, true);
            }
            break;
        }
        case K_Q: {
            int s = n_slots;
            n_slots += SLOTS_Q;
            slot_of_elem[ei] = s;
            int c = e.node[0], b = e.node[1], em = e.node[2];
            for (auto &m : mb) {
                m.nl.insert(m.nl.end(), {K_Q, c, b, em, s});
                m.n_gfun += 3;
                m.A(c, c, s + Q_CC, false, NL);
                m.B(c, s + Q_RC, true, NL);
                m.A(b, b, s + Q_BB, false, NL);
                m.B(b, s + Q_RB, false, NL);
                m.A(em, em, s + Q_EE, false, NL);
                m.B(em, s + Q_RE, true, NL);
                m.A(em, c, s + Q_EC, true, NL);
                m.A(c, em, s + Q_CE, true, NL);
                m.A(em, b, s + Q_EB, false, NL);
                m.A(b, em, s + Q_BE, false, NL);
                m.A(c, b, s + Q_CB, false, NL);
                m.A(b, c, s + Q_BC, false, NL);
                m.F(c, s + Q_IC, false);
                m.F(b, s + Q_IB, false);
                m.F(em, s + Q_IE, false);
            }
            break;
        }
        case K_M: {
            int s = n_slots;
            n_slots += SLOTS_M;
            slot_of_elem[ei] = s;
            int d = e.node[0], g = e.node[1], sn = e.node[2];
            for (auto &m : mb) {
                m.nl.insert(m.nl.end(), {K_M, d, g, sn, s});
                m.n_gfun += 3;
                // signs (and the run-time drain/source swap) are folded into the slot values
                m.A(d, d, s + M_DD, false, NL);
                m.B(d, s + M_RD, false, NL);
                m.A(sn, sn, s + M_SS, false, NL);
                m.B(sn, s + M_RS, false, NL);
                if (d >= 0 && sn >= 0) {
                    m.A(d, sn, s + M_DS, false, NL);
                    m.A(sn, d, s + M_SD, false, NL);
                }
                m.A(d, g, s + M_DG, false, NL);
                m.A(sn, g, s + M_SG, false, NL);
                m.F(d, s + M_ID, false);
                m.F(sn, s + M_ID, true);
            }
            break;
        }
        default:
            err = "ftsb_compile: element " + std::to_string(ei) + " has unknown kind " + std::to_string(e.kind);
            return false;
        }
    }

    out.n_slots = n_slots;
    out.n_fn = n_fn;
    out.n_res = (int)res.size() / 2;
    out.n_src = (int)src.size() / 5;
    out.n_dyn = (int)dyn.size() / 6;
    out.off_res = im.push(res);
    out.off_src = im.push(src);
    out.off_dyn = im.push(dyn);
    out.off_src_of_elem = im.push(src_of_elem);
    out.slot_of_elem = slot_of_elem;
    out.src_of_elem = src_of_elem;
    out.kinds.resize(n_elems);
    for (int i = 0; i < n_elems; ++i) out.kinds[i] = elems[i].kind;

    for (int mi = 0; mi < 3; ++mi) {
        ModeBuild &m = mb[mi];
        ModeOff &o = out.mode[mi];
        o.n = m.n;
        o.nv = o.ni = 0;
        std::vector<int> cls(std::max(m.n, 1), 0);
        for (int i = 0; i < m.n; ++i) {
            cls[i] = row_class[i] ? 1 : 0;
            (cls[i] ? o.ni : o.nv)++;
        }
        o.cls = im.push(cls);
        o.n_nl = (int)m.nl.size() / 5;
        o.n_gfun = m.n_gfun;
        o.nl = im.push(m.nl);
        // matrix entries outside the mode's n x n block cannot occur: all indices were range-checked against
        // n_run (nodes, V rows) or used only in OP mode (inductor rows)
        put_gather(im, make_gather(m.a, LIN, 1u << LIN), o.A_lin);
        put_gather(im, make_gather(m.a, DYN, (1u << LIN) | (1u << DYN)), o.A_dyn);
        put_gather(im, make_gather(m.a, NL, 1u << NL), o.J_nl);
        {
            // b_all, row-indexed: b[row] = linear then companion rhs terms (rows without sources are empty)
            GatherBuild g;
            g.ptr.push_back(0);
            for (int row = 0; row < m.n; ++row) {
                g.dst.push_back(row);
                for (int ph = 0; ph < 2; ++ph)
                    for (const Con &c : m.b)
                        if (c.phase == ph && c.dst == row) g.con.push_back((int)c.word);
                g.ptr.push_back((int)g.con.size());
            }
            put_gather(im, g, o.b_all);
        }
        {
            // J_full: every structurally non-zero entry, all phases in order (linear, companion, nonlinear)
            std::vector<Con> all = m.a;
            for (Con &c : all) c.phase = 0;
            GatherBuild g;
            std::map<int, std::vector<const Con *>> by_dst;
            for (const Con &c : m.a) by_dst[c.dst].push_back(&c);
            g.ptr.push_back(0);
            for (auto &kv : by_dst) {
                g.dst.push_back(kv.first);
                for (int ph = 0; ph < 3; ++ph)
                    for (const Con *c : kv.second)
                        if (c->phase == ph) g.con.push_back((int)c->word);
                g.ptr.push_back((int)g.con.size());
            }
            put_gather(im, g, o.J_full);
        }
        {
            // Static structure for the sparse solver: J_full's pattern + predicted fill-in.
            // The prediction eliminates the pattern with a guessed pivot per column: a candidate whose entry is a
            // bare +-1 incidence constant (V-source / start-up inductor rows; conductances are normally < 1), else
            // the row sitting at the diagonal position, else the candidate at the smallest position.  A wrong guess
            // costs speed only: entries outside the static set become run-time fill-in.
            const int n = m.n;
            std::map<int, std::vector<const Con *>> by_dst; // key = row | col << 16: column-major order
            for (const Con &c : m.a) by_dst[c.dst].push_back(&c);
            std::vector<std::map<int, bool>> rows(std::max(n, 1)); // row -> {col -> is a bare unit constant}
            for (auto &kv : by_dst) {
                const int r = kv.first & 0xffff, cc = kv.first >> 16;
                bool unit = kv.second.size() == 1 && (kv.second[0]->word & ~CON_NEG) == (unsigned)SLOT_ONE;
                rows[r][cc] = unit;
            }
            std::vector<std::map<int, bool>> work = rows;
            std::vector<int> pos(std::max(n, 1)), piv(std::max(n, 1));
            std::vector<char> noswap(std::max(n, 1), 0);
            for (int i = 0; i < n; ++i) pos[i] = piv[i] = i;
            bool piv_valid = n > 0;
            for (int k = 0; k < n; ++k) {
                int P = -1, bestpos = 1 << 30;
                bool bestunit = false;
                for (int r = 0; r < n; ++r) {
                    if (pos[r] < k) continue;
                    auto it = work[r].find(k);
                    if (it == work[r].end()) continue;
                    const bool unit = it->second;
                    const int rank = unit ? 0 : (pos[r] == k ? 1 : 2);
                    const int brank = P < 0 ? 3 : (bestunit ? 0 : (bestpos == k ? 1 : 2));
                    if (rank < brank || (rank == brank && pos[r] < bestpos)) {
                        P = r;
                        bestpos = pos[r];
                        bestunit = unit;
                    }
                }
                if (P < 0) { // structurally singular column: nothing to predict
                    piv_valid = false;
                    continue;
                }
                const int rk = piv[k], pp = pos[P];
                noswap[k] = pp == k;
                piv[k] = P;
                piv[pp] = rk;
                pos[P] = k;
                pos[rk] = pp;
                for (int r = 0; r < n; ++r) {
                    if (pos[r] <= k || !work[r].count(k)) continue;
                    for (auto &pc : work[P])
                        if (pc.first > k) work[r][pc.first] = false; // updated (or filled) entries are not constants
                }
            }
            // entries: column-major over the union pattern
            std::vector<std::vector<int>> cols(std::max(n, 1)); // col -> rows ascending
            int n_pred = 0;
            for (int r = 0; r < n; ++r)
                for (auto &pc : work[r]) {
                    cols[pc.first].push_back(r);
                    if (!rows[r].count(pc.first)) ++n_pred;
                }
            std::vector<int> cptr(n + 1, 0), erow, rptr(n + 1, 0), rcol, rent;
            std::map<int, int> ent_of; // row | col << 16 -> e
            for (int cc = 0; cc < n; ++cc) {
                std::sort(cols[cc].begin(), cols[cc].end());
                for (int r : cols[cc]) {
                    ent_of[r | (cc << 16)] = (int)erow.size();
                    erow.push_back(r);
                }
                cptr[cc + 1] = (int)erow.size();
            }
            for (int r = 0; r < n; ++r) {
                for (auto &pc : work[r]) { // std::map: columns ascending
                    rcol.push_back(pc.first);
                    rent.push_back(ent_of[r | (pc.first << 16)]);
                }
                rptr[r + 1] = (int)rcol.size();
            }
            GatherBuild g;
            g.ptr.push_back(0);
            for (int e = 0; e < (int)erow.size(); ++e) {
                g.dst.push_back(e);
                // column of e: find by cptr
                int cc = (int)(std::upper_bound(cptr.begin(), cptr.end(), e) - cptr.begin()) - 1;
                auto it = by_dst.find(erow[e] | (cc << 16));
                if (it != by_dst.end())
                    for (int ph = 0; ph < 3; ++ph)
                        for (const Con *c : it->second)
                            if (c->phase == ph) g.con.push_back((int)c->word);
                g.ptr.push_back((int)g.con.size());
            }
            {
                // how many consecutive columns the windowed elimination (solver_sparse.cuh) could take at once under the
                // predicted pivots: a window ends at a swap or where an earlier pivot row of the window holds the column
                int windows = 0;
                for (int k0 = 0; k0 < n;) {
                    int ext = 0;
                    for (int l = 0; l < 32 && k0 + l < n; ++l) {
                        const int k = k0 + l;
                        bool okc = noswap[k] != 0;
                        for (int kp = k0; kp < k && okc; ++kp)
                            if (work[piv[kp]].count(k)) okc = false;
                        if (!okc) break;
                        ++ext;
                    }
                    k0 += std::max(ext, 1);
                    ++windows;
                }
                o.sp.window_hint = windows ? n / windows : 0;
            }
            o.sp.nnz = (int)erow.size();
            o.sp.n_pred = n_pred;
            o.sp.piv = im.push(piv);
            o.sp.piv_valid = piv_valid ? 1 : 0;
            o.sp.cptr = im.push(cptr);
            o.sp.erow = im.push(erow);
            o.sp.rptr = im.push(rptr);
            o.sp.rcol = im.push(rcol);
            o.sp.rent = im.push(rent);
            {
                // direct column -> entry maps for the longest rows (a supply rail touches every stage)
                std::vector<std::pair<int, int>> bylen;
                for (int r = 0; r < n; ++r)
                    if (rptr[r + 1] - rptr[r] > SP_LONG_ROW) bylen.push_back({-(rptr[r + 1] - rptr[r]), r});
                std::sort(bylen.begin(), bylen.end());
                if ((int)bylen.size() > SP_MAX_LONG) bylen.resize(SP_MAX_LONG);
                std::vector<int> rlong(std::max(n, 1), -1), lmap;
                for (size_t li = 0; li < bylen.size(); ++li) {
                    const int r = bylen[li].second;
                    rlong[r] = (int)li;
                    std::vector<int> mp(n, -1);
                    for (int t = rptr[r]; t < rptr[r + 1]; ++t) mp[rcol[t]] = rent[t];
                    lmap.insert(lmap.end(), mp.begin(), mp.end());
                }
                o.sp.n_long = (int)bylen.size();
                o.sp.rlong = im.push(rlong);
                o.sp.lmap = im.push(lmap);
            }
            put_gather(im, g, o.sp.S);
        }
        {
            // Ax, row-indexed flat list of (col, slot, sign) for the linear + companion part
            GatherBuild g;
            g.ptr.push_back(0);
            for (int row = 0; row < m.n; ++row) {
                g.dst.push_back(row);
                for (int ph = 0; ph < 2; ++ph)
                    for (const Con &c : m.a) {
                        const int r = c.dst & 0xffff, col = c.dst >> 16;
                        if (c.phase == ph && r == row) {
                            if ((c.word & ~CON_NEG) > AX_SLOT_MASK || col >= (1 << (31 - AX_COL_SHIFT))) {
                                err = "ftsb_compile: circuit too large for the packed A*x map";
                                return false;
                            }
                            g.con.push_back((int)(c.word | ((unsigned)col << AX_COL_SHIFT)));
                        }
                    }
                g.ptr.push_back((int)g.con.size());
            }
            put_gather(im, g, o.Ax);
        }
        {
            // r_nl is indexed by row as well (full CSR): r[row] = b[row] + nonlinear rhs terms
            GatherBuild g;
            g.ptr.push_back(0);
            for (int row = 0; row < m.n; ++row) {
                g.dst.push_back(row);
                for (const Con &c : m.b)
                    if (c.phase == NL && c.dst == row) g.con.push_back((int)c.word);
                g.ptr.push_back((int)g.con.size());
            }
            put_gather(im, g, o.r_nl);
        }
        {
            // f_nl is indexed by row: a full CSR over all n rows (rows without device currents are empty)
            GatherBuild g;
            g.ptr.push_back(0);
            for (int row = 0; row < m.n; ++row) {
                g.dst.push_back(row);
                for (const Con &c : m.f)
                    if (c.dst == row) g.con.push_back((int)c.word);
                g.ptr.push_back((int)g.con.size());
            }
            put_gather(im, g, o.f_nl);
        }
    }

    // per-instance layout
    Layout &L = out.lay;
    L.ld = ld;
    int o = 0;
    auto take = [&](int cnt) {
        int r = o;
        o += (cnt + 1) & ~1; // keep 16-byte alignment
        return r;
    };
    L.ob = take(ld);
    L.orr = take(ld);
    L.ox = take(ld);
    L.oxn = take(ld);
    L.oxp = take(ld);
    L.of = take(ld);
    L.oval = take(n_slots);
    L.ostate = take(2 * out.n_dyn);
    L.oring = take(3 * ld);
    L.oprev = take(3 * out.n_src); // value at t=0, at the last evaluation, at the current attempt
    L.doubles = o;
    L.mat_doubles = ((ld * ld) + 1) & ~1;
    L.oA = 0; // filled by the launcher (smem or global scratch)
    L.oJ = 0;
    L.opiv = 0;
    L.ints = 2 * ((ld + 3) & ~3) + 8; // piv[ld] + the CTA solver's non-zero column list and its two counters
    L.mat_in_smem = 0;
    L.bytes = 0;
    out.ibuf = &im.ibuf;
    return true;
}

const std::vector<int> &Compiled::ints() const { return impl->ibuf; }

} // namespace ftsb
