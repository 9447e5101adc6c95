// team_plan.cpp — symbolic elimination along a fixed pivot sequence (team_plan.h)
#include "team_plan.h"

#include <algorithm>
#include <map>

namespace ftsb {

TeamPlan TeamPlanHost::view(const int *b) const
{
    TeamPlan p;
    p.n = n;
    p.nA = nA;
    p.S = S;
    p.nJ = nJ;
    p.arow = b + o_arow;
    p.acol = b + o_acol;
    p.aptr = b + o_aptr;
    p.acon = b + o_acon;
    p.sptr = b + o_sptr;
    p.scon = b + o_scon;
    p.spre = b + o_spre;
    p.nPre = nPre;
    p.pivrow = b + o_pivrow;
    p.pive = b + o_pive;
    p.tptr = b + o_tptr;
    p.tgt = b + o_tgt;
    p.tgtr = b + o_tgtr;
    p.uptr = b + o_uptr;
    p.updm = b + o_updm;
    p.updds = b + o_updds;
    p.bptr = b + o_bptr;
    p.bwde = b + o_bwde;
    p.bwdr = b + o_bwdr;
    p.ellW = ellW;
    p.nSlotA = nSlotA;
    p.ellcol = b + o_ellcol;
    p.ovfptr = b + o_ovfptr;
    p.ovfcol = b + o_ovfcol;
    p.longRow = longRow;
    p.chain = chain;
    p.nFop = nFop;
    p.nBop = nBop;
    p.nFchunk = nFchunk;
    p.nBchunk = nBchunk;
    p.fop = b + o_fop;
    p.fvidx = b + o_fvidx;
    p.fchunk = b + o_fchunk;
    p.bop = b + o_bop;
    p.bvidx = b + o_bvidx;
    p.bchunk = b + o_bchunk;
    p.nEarly = nEarly;
    p.nXop = nXop;
    p.nDFop = nDFop;
    p.nDBop = nDBop;
    p.early = b + o_early;
    p.xop = b + o_xop;
    p.dfop = b + o_dfop;
    p.dbop = b + o_dbop;
    p.levels = 1;
    p.nXlev = nXlev;
    p.nFlev = nFlev;
    p.nBlev = nBlev;
    p.xlptr = b + o_xlptr;
    p.xlop = b + o_xlop;
    p.flptr = b + o_flptr;
    p.flop = b + o_flop;
    p.blptr = b + o_blptr;
    p.blop = b + o_blop;
    p.nXseg = nXseg;
    p.nBseg = nBseg;
    p.xseg = b + o_xseg;
    p.xsop = b + o_xsop;
    p.bseg = b + o_bseg;
    p.bsop = b + o_bsop;
    return p;
}

namespace {
// level schedule of an op list (team_plan.h): locations are small integers; reads / writes per op
struct LevOp {
    int w0, w1;
    std::vector<int> reads, writes;
};
void levelize(const std::vector<LevOp> &ops, int nloc, std::vector<int> &lptr, std::vector<int> &words)
{
    std::vector<int> ready(nloc, 0), lastread(nloc, 0), lev(ops.size(), 0);
    int nlev = 0;
    for (size_t o = 0; o < ops.size(); ++o) {
        int L = 0;
        for (int r : ops[o].reads) L = std::max(L, ready[r]);
        for (int w : ops[o].writes) L = std::max(L, std::max(ready[w], lastread[w]));
        L += 1;
        for (int r : ops[o].reads) lastread[r] = std::max(lastread[r], L);
        for (int w : ops[o].writes) ready[w] = L;
        lev[o] = L;
        nlev = std::max(nlev, L);
    }
    lptr.assign(nlev + 1, 0);
    for (size_t o = 0; o < ops.size(); ++o) lptr[lev[o]]++;
    for (int L = 1; L <= nlev; ++L) lptr[L] += lptr[L - 1];
    std::vector<int> cur(lptr.begin(), lptr.end() - 1); // lptr[L - 1] = first op of level L (1-based levels)
    words.assign(2 * ops.size(), 0);
    for (size_t o = 0; o < ops.size(); ++o) { // stable: the original order inside a level
        const int at = cur[lev[o] - 1]++;
        words[2 * at] = ops[o].w0;
        words[2 * at + 1] = ops[o].w1;
    }
}
// The level schedule in the form the kernel walks (team_plan.h): maximal runs of NARROW levels (at most TEAM_NARROW ops) are
// stored padded to TEAM_NARROW slots per level (padding: the no-op word), every WIDE level as its plain op list.
void segment(const std::vector<int> &lptr, const std::vector<int> &words, std::vector<int> &seg, std::vector<int> &pops)
{
    seg.clear();
    pops.clear();
    const int nlev = (int)lptr.size() - 1;
    int L = 0;
    while (L < nlev) {
        const int cnt = lptr[L + 1] - lptr[L];
        if (cnt > TEAM_NARROW) {
            seg.push_back(-cnt);
            seg.push_back((int)pops.size() / 2);
            pops.insert(pops.end(), words.begin() + 2 * lptr[L], words.begin() + 2 * lptr[L + 1]);
            ++L;
            continue;
        }
        int L1 = L;
        while (L1 < nlev && lptr[L1 + 1] - lptr[L1] <= TEAM_NARROW) ++L1;
        seg.push_back(L1 - L);
        seg.push_back((int)pops.size() / 2);
        for (int j = L; j < L1; ++j)
            for (int k = 0; k < TEAM_NARROW; ++k) {
                const int o = lptr[j] + k;
                if (o < lptr[j + 1]) {
                    pops.push_back(words[2 * o]);
                    pops.push_back(words[2 * o + 1]);
                } else {
                    pops.push_back(0);
                    pops.push_back(TEAM_OP_NOP);
                }
            }
        L = L1;
    }
}
} // namespace

bool build_team_plan(const Compiled &comp, int an, const std::vector<int> &pivseq, TeamPlanHost &out)
{
    const std::vector<int> &ib = comp.ints();
    const ModeOff &mo = comp.mode[an];
    const int n = mo.n;
    if (n < 1 || (int)pivseq.size() != n) return false;
    out = TeamPlanHost();
    out.n = n;

    // ---- A: the linear (+ companion) part, the matrix of the residual A x - b (mna.rs:25-29) ----
    // an entry touched by a companion stamp carries its complete list in A_dyn (linear then companion contributions),
    // every other entry its list in A_lin
    std::map<int, std::vector<int>> a_ent; // row << 16 | col -> contribution words (row-major order of the keys)
    auto take = [&](const GatherOff &g, bool overwrite) {
        for (int d = 0; d < g.n_dst; ++d) {
            const int dst = ib[g.dst + d], r = dst & 0xffff, c = dst >> 16;
            const int key = (r << 16) | c;
            if (!overwrite && a_ent.count(key)) continue;
            std::vector<int> v(ib.begin() + g.con + ib[g.ptr + d], ib.begin() + g.con + ib[g.ptr + d + 1]);
            a_ent[key] = v;
        }
    };
    take(mo.A_lin, true);
    take(mo.A_dyn, true);
    std::vector<int> arow(n + 1, 0), acol;
    std::vector<std::vector<std::pair<int, const std::vector<int> *>>> by_row(n); // row -> (col, contributions), columns ascending
    for (auto &kv : a_ent) {
        const int r = kv.first >> 16, c = kv.first & 0xffff;
        arow[r + 1]++;
        acol.push_back(c);
        by_row[r].push_back({c, &kv.second});
    }
    for (int r = 0; r < n; ++r) arow[r + 1] += arow[r];
    out.nA = (int)acol.size();
    int maxlen = 0;
    for (int r = 0; r < n; ++r) maxlen = std::max(maxlen, (int)by_row[r].size());
    const int W = std::max(3, std::min(4, maxlen)); // the kernels are built for 3 and 4 ELL slots per row
    // rows 0 .. n-1 and one ZERO row n (all slots 0.0, columns n): lanes dealt a row index >= n in the last round of a
    // row loop clamp it to n instead of branching
    std::vector<int> ellcol(2 * (n + 1), 0), ovfptr{0}, ovfcol;
    std::vector<std::vector<int>> slot_con((size_t)W * (n + 1)); // contributions of A slot row * W + w (empty: 0.0)
    for (int r = 0; r <= n; ++r) {
        int cols[4] = {r, r, r, r};
        if (r < n)
            for (int w = 0; w < (int)by_row[r].size(); ++w) {
                if (w < W) {
                    cols[w] = by_row[r][w].first;
                    slot_con[(size_t)r * W + w] = *by_row[r][w].second;
                } else {
                    ovfcol.push_back(by_row[r][w].first);
                    slot_con.push_back(*by_row[r][w].second);
                }
            }
        ellcol[2 * r] = cols[0] | (cols[1] << 16);
        ellcol[2 * r + 1] = cols[2] | (cols[3] << 16);
        if (r < n) ovfptr.push_back((int)ovfcol.size());
    }
    std::vector<int> aptr{0}, acon;
    for (auto &v : slot_con) {
        acon.insert(acon.end(), v.begin(), v.end());
        aptr.push_back((int)acon.size());
    }
    out.ellW = W;
    out.nSlotA = (int)slot_con.size();
    {
        int best = 15;
        for (int r = 0; r < n; ++r)
            if (ovfptr[r + 1] - ovfptr[r] > best) {
                best = ovfptr[r + 1] - ovfptr[r];
                out.longRow = r;
            }
    }

    // ---- J: static entries in J_full order, then the fill-in of this pivot sequence ----
    const GatherOff &jf = mo.J_full;
    const int nJ = jf.n_dst;
    std::vector<std::map<int, int>> rows(n); // row -> {col -> entry}
    for (int d = 0; d < nJ; ++d) {
        const int dst = ib[jf.dst + d];
        rows[dst & 0xffff][dst >> 16] = d;
    }
    std::vector<int> sptr(ib.begin() + jf.ptr, ib.begin() + jf.ptr + nJ + 1), scon(ib.begin() + jf.con, ib.begin() + jf.con + jf.n_con);
    // linear prefix of every static entry: the contributions it shares, in order, with the entry of A (A's list is J's list
    // without the nonlinear stamps)
    std::vector<int> spre(nJ, 0);
    for (int d = 0; d < nJ; ++d) {
        const int dst = ib[jf.dst + d], key = ((dst & 0xffff) << 16) | (dst >> 16);
        auto it = a_ent.find(key);
        if (it == a_ent.end()) continue;
        const std::vector<int> &av = it->second;
        int k = 0;
        while (sptr[d] + k < sptr[d + 1] && k < (int)av.size() && scon[sptr[d] + k] == av[k]) ++k;
        // a prefix of one contribution costs a load either way: kept sums only from two on, numbered compactly (the sums
        // live in the L2-resident scratch row of the instance)
        if (k >= 2 && k < 65536 && out.nPre < 65535) spre[d] = (k << 16) | out.nPre++;
    }
    int S = nJ;
    std::vector<int> pos(n), piv(n);
    for (int i = 0; i < n; ++i) pos[i] = piv[i] = i;
    std::vector<int> pivrow(n), pive(n), tptr{0}, tgt, tgtr, uptr{0}, updm, updds;
    for (int k = 0; k < n; ++k) {
        const int P = pivseq[k];
        if (P < 0 || P >= n || pos[P] < k) return false;
        auto itp = rows[P].find(k);
        if (itp == rows[P].end()) return false;
        const int pP = pos[P];
        pivrow[k] = P;
        pive[k] = itp->second;
        std::vector<int> tr;
        for (int r = 0; r < n; ++r)
            if (r != P && pos[r] >= k && rows[r].count(k)) tr.push_back(r);
        for (int r : tr) {
            const int em = rows[r][k];
            tgt.push_back(em | (pos[r] < pP ? (1 << 30) : 0));
            tgtr.push_back(r);
            for (auto &pc : rows[P]) {
                if (pc.first <= k) continue;
                int dst;
                auto it = rows[r].find(pc.first);
                if (it == rows[r].end()) {
                    dst = S++;
                    rows[r][pc.first] = dst;
                } else
                    dst = it->second;
                updm.push_back(em);
                updds.push_back(dst | (pc.second << 16));
            }
            uptr.push_back((int)updm.size());
        }
        tptr.push_back((int)tgt.size());
        const int rk = piv[k];
        piv[k] = P;
        piv[pP] = rk;
        pos[P] = k;
        pos[rk] = pP;
    }
    if (S + n >= 0xfff0 || out.nSlotA >= 0xfff0 || n >= 0x7ff0) return false; // 16-bit indices in the packed lists
    std::vector<int> bptr{0}, bwde, bwdr;
    for (int i = 0; i < n; ++i) {
        for (int r = 0; r < n; ++r) {
            if (pos[r] >= i) continue;
            auto it = rows[r].find(i);
            if (it != rows[r].end()) {
                bwde.push_back(it->second);
                bwdr.push_back(r);
            }
        }
        bptr.push_back((int)bwde.size());
    }
    // ---- substitutions as op streams with chunked value slots ----
    {
        int maxt = 0;
        for (int k = 0; k < n; ++k) maxt = std::max(maxt, tptr[k + 1] - tptr[k]);
        int maxb = 0;
        for (int i = 0; i < n; ++i) maxb = std::max(maxb, bptr[i + 1] - bptr[i]);
        out.chain = (maxt <= 2 && maxb <= 2) ? 1 : 0; // ladders, chains of two-terminal stages; a supply rail has wide steps
        std::vector<int> fop, fvidx, fchunk{0}, bop, bvidx, bchunk{0};
        // an op's value slots lie inside one chunk; a chunk that is full (or too full for the op) is closed first
        auto begin_op = [&](std::vector<int> &vidx, std::vector<int> &chunk, int n_ops_so_far, int nvals) {
            int used = (int)(vidx.size() % TEAM_CHUNK);
            if (used + nvals > TEAM_CHUNK) {
                while (vidx.size() % TEAM_CHUNK) vidx.push_back(0);
                used = 0;
            }
            if (used == 0 && !vidx.empty()) chunk.push_back(n_ops_so_far);
        };
        auto end_stream = [&](std::vector<int> &vidx, std::vector<int> &chunk, int n_ops) {
            if (vidx.empty()) vidx.push_back(0);
            while (vidx.size() % TEAM_CHUNK) vidx.push_back(0);
            chunk.push_back(n_ops);
        };
        for (int k = 0; k < n; ++k)
            for (int tt = tptr[k]; tt < tptr[k + 1]; ++tt) {
                begin_op(fvidx, fchunk, (int)fop.size(), 1);
                fop.push_back(tgtr[tt] | (pivrow[k] << 16));
                fvidx.push_back(tgt[tt] & 0xffff);
            }
        end_stream(fvidx, fchunk, (int)fop.size());
        for (int i = n - 1; i >= 0; --i) {
            begin_op(bvidx, bchunk, (int)bop.size() / 2, 2);
            bop.push_back(pivrow[i] | (i << 16));
            bop.push_back(-1);
            bvidx.push_back(pive[i]);
            bvidx.push_back(S + i);
            for (int bb = bptr[i]; bb < bptr[i + 1]; ++bb) {
                begin_op(bvidx, bchunk, (int)bop.size() / 2, 1);
                bop.push_back(bwdr[bb]);
                bop.push_back(0);
                bvidx.push_back(bwde[bb]);
            }
        }
        end_stream(bvidx, bchunk, (int)bop.size() / 2);
        out.nFop = (int)fop.size();
        out.nBop = (int)bop.size() / 2;
        out.nFchunk = (int)fchunk.size() - 1;
        out.nBchunk = (int)bchunk.size() - 1;
        if (out.nFchunk * TEAM_CHUNK != (int)fvidx.size() || out.nBchunk * TEAM_CHUNK != (int)bvidx.size()) return false;
        auto push2 = [&](const std::vector<int> &v) {
            const int off = (int)out.ibuf.size();
            out.ibuf.insert(out.ibuf.end(), v.begin(), v.end());
            while (out.ibuf.size() & 3) out.ibuf.push_back(0);
            return off;
        };
        out.o_fop = push2(fop);
        out.o_fvidx = push2(fvidx);
        out.o_fchunk = push2(fchunk);
        out.o_bop = push2(bop);
        out.o_bvidx = push2(bvidx);
        out.o_bchunk = push2(bchunk);
        // ---- direct streams (factors in shared memory) ----
        {
            std::vector<char> written(S, 0); // entries some update of an EARLIER step writes
            std::vector<int> early, xop, dfop, dbop;
            for (int k = 0; k < n; ++k) {
                bool is_early = !written[pive[k]];
                for (int tt = tptr[k]; tt < tptr[k + 1]; ++tt) is_early = is_early && !written[tgt[tt] & 0xffff];
                if (is_early)
                    early.push_back(k);
                else {
                    xop.push_back(k);
                    xop.push_back((int)0x80000000u);
                }
                for (int tt = tptr[k]; tt < tptr[k + 1]; ++tt) {
                    for (int u = uptr[tt]; u < uptr[tt + 1]; ++u) {
                        xop.push_back(updds[u]);
                        xop.push_back(updm[u]);
                        written[updds[u] & 0xffff] = 1;
                    }
                    dfop.push_back(tgtr[tt] | (pivrow[k] << 16));
                    dfop.push_back(tgt[tt] & 0xffff);
                }
            }
            for (int i = n - 1; i >= 0; --i) {
                dbop.push_back(pivrow[i] | (i << 16));
                dbop.push_back((int)(0x80000000u | (unsigned)pive[i]));
                for (int bb = bptr[i]; bb < bptr[i + 1]; ++bb) {
                    dbop.push_back(bwdr[bb]);
                    dbop.push_back(bwde[bb]);
                }
            }
            out.nEarly = (int)early.size();
            out.nXop = (int)xop.size() / 2;
            out.nDFop = (int)dfop.size() / 2;
            out.nDBop = (int)dbop.size() / 2;
            out.o_early = push2(early);
            out.o_xop = push2(xop);
            out.o_dfop = push2(dfop);
            out.o_dbop = push2(dbop);
            // ---- level schedules (locations: factor entry e -> e, reciprocal k -> S + k, y[r] -> r, x[i] -> n + i) ----
            std::vector<LevOp> ops;
            std::vector<int> lptr, words;
            for (int k = 0; k < n; ++k) {
                LevOp hd;
                hd.w0 = k;
                hd.w1 = (int)0x80000000u;
                hd.reads.push_back(pive[k]);
                for (int tt = tptr[k]; tt < tptr[k + 1]; ++tt) {
                    hd.reads.push_back(tgt[tt] & 0xffff);
                    hd.writes.push_back(tgt[tt] & 0xffff);
                }
                hd.writes.push_back(S + k);
                ops.push_back(hd);
                for (int tt = tptr[k]; tt < tptr[k + 1]; ++tt)
                    for (int u = uptr[tt]; u < uptr[tt + 1]; ++u) {
                        LevOp up;
                        up.w0 = updds[u];
                        up.w1 = updm[u];
                        up.reads = {updm[u], (updds[u] >> 16) & 0xffff, updds[u] & 0xffff};
                        up.writes = {updds[u] & 0xffff};
                        ops.push_back(up);
                    }
            }
            // the forward substitution rides along (the right-hand side as one more column): its ops only need their multiplier
            // and the final y of the pivot row, so they fill the same levels as the elimination instead of 256 levels of their own
            for (int k = 0; k < n; ++k)
                for (int tt = tptr[k]; tt < tptr[k + 1]; ++tt) {
                    LevOp f;
                    f.w0 = tgtr[tt] | (pivrow[k] << 16);
                    f.w1 = (tgt[tt] & 0xffff) | 0x40000000;
                    f.reads = {tgt[tt] & 0xffff, S + n + tgtr[tt], S + n + pivrow[k]};
                    f.writes = {S + n + tgtr[tt]};
                    ops.push_back(f);
                }
            levelize(ops, S + 2 * n, lptr, words);
            out.nXlev = (int)lptr.size() - 1;
            out.o_xlptr = push2(lptr);
            out.o_xlop = push2(words);
            {
                std::vector<int> seg, pops;
                segment(lptr, words, seg, pops);
                out.nXseg = (int)seg.size() / 2;
                out.o_xseg = push2(seg);
                out.o_xsop = push2(pops);
            }
            out.nFlev = 0; // (kept in the interface: a separate forward schedule)
            lptr.assign(1, 0);
            words.clear();
            out.o_flptr = push2(lptr);
            out.o_flop = push2(words);
            ops.clear();
            for (int i = n - 1; i >= 0; --i) {
                LevOp dv;
                dv.w0 = pivrow[i] | (i << 16);
                dv.w1 = (int)(0x80000000u | (unsigned)pive[i]);
                dv.reads = {pivrow[i]};
                dv.writes = {n + i};
                ops.push_back(dv);
                for (int bb = bptr[i]; bb < bptr[i + 1]; ++bb) {
                    LevOp up;
                    up.w0 = bwdr[bb] | (i << 16);
                    up.w1 = bwde[bb];
                    up.reads = {n + i, bwdr[bb]};
                    up.writes = {bwdr[bb]};
                    ops.push_back(up);
                }
            }
            levelize(ops, 2 * n, lptr, words);
            out.nBlev = (int)lptr.size() - 1;
            out.o_blptr = push2(lptr);
            out.o_blop = push2(words);
            {
                std::vector<int> seg, pops;
                segment(lptr, words, seg, pops);
                out.nBseg = (int)seg.size() / 2;
                out.o_bseg = push2(seg);
                out.o_bsop = push2(pops);
            }
        }
        out.o_ellcol = push2(ellcol);
        out.o_ovfptr = push2(ovfptr);
        out.o_ovfcol = push2(ovfcol);
    }
    out.S = S;
    out.nJ = nJ;
    out.nT = (int)tgt.size();
    out.nU = (int)updm.size();
    out.nB = (int)bwde.size();
    out.flops_factor = (long long)n + out.nT + 2LL * out.nU;
    out.flops_solve = 2LL * out.nT + n + 2LL * out.nB;
    auto push = [&](const std::vector<int> &v) {
        const int off = (int)out.ibuf.size();
        out.ibuf.insert(out.ibuf.end(), v.begin(), v.end());
        while (out.ibuf.size() & 3) out.ibuf.push_back(0);
        return off;
    };
    out.o_arow = push(arow);
    out.o_acol = push(acol);
    out.o_aptr = push(aptr);
    out.o_acon = push(acon);
    out.o_sptr = push(sptr);
    out.o_scon = push(scon);
    out.o_spre = push(spre);
    out.o_pivrow = push(pivrow);
    out.o_pive = push(pive);
    out.o_tptr = push(tptr);
    out.o_tgt = push(tgt);
    out.o_tgtr = push(tgtr);
    out.o_uptr = push(uptr);
    out.o_updm = push(updm);
    out.o_updds = push(updds);
    out.o_bptr = push(bptr);
    out.o_bwde = push(bwde);
    out.o_bwdr = push(bwdr);
    return true;
}

} // namespace ftsb
