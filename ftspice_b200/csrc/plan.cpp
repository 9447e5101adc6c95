// plan.cpp — host side of the learned elimination plans (plan.h): symbolic elimination along each recorded pivot
// sequence (gauss_lu.rs:3-54 on the structure only).
#include <cuda_runtime.h> // int4

#include <algorithm>
#include <map>

#include "plan.h"

namespace ftsb {

namespace {
struct StepBuild {
    int eP = 0, prow = 0;
    std::vector<int> cand, mult, upd, bu;
};
} // namespace

bool build_plans(int n, const std::vector<int> &jfull_dst, std::vector<std::pair<uint64_t, long long>> perms, PlanHost &out,
                 int max_plans)
{
    out = PlanHost();
    if (n < 2 || n > 16 || perms.empty()) return false;
    std::sort(perms.begin(), perms.end(), [](const auto &a, const auto &b) { return a.second > b.second; });
    long long total = 0;
    for (auto &p : perms) total += p.second;

    // compact entry numbering: the static pattern first (J_full order), fill-in of the plans after it
    std::map<int, int> ent; // row | col << 8 -> compact index
    std::vector<int> jdst;
    for (int d : jfull_dst) {
        const int r = d & 0xffff, c = d >> 16;
        if (r >= n || c >= n) return false;
        const int key = r | (c << 8);
        if (!ent.count(key)) {
            const int id = (int)ent.size();
            ent[key] = id;
        }
        jdst.push_back(ent[key]);
    }
    const int nnz0 = (int)ent.size();
    auto entry = [&](int r, int c) {
        const int key = r | (c << 8);
        auto it = ent.find(key);
        if (it != ent.end()) return it->second;
        const int id = (int)ent.size();
        ent[key] = id;
        return id;
    };

    std::vector<std::vector<int>> seqs;             // accepted pivot sequences
    std::vector<std::vector<StepBuild>> plans;      // [plan][step]
    long long covered = 0;
    for (auto &pc : perms) {
        if ((int)plans.size() >= std::min(PLAN_MAX, std::max(max_plans, 1))) break;
        std::vector<int> piv(n);
        unsigned seen = 0;
        for (int i = 0; i < n; ++i) {
            piv[i] = (int)((pc.first >> (4 * i)) & 15u);
            if (piv[i] >= n) seen = ~0u;
            else seen |= 1u << piv[i];
        }
        if (seen != (1u << n) - 1u) continue; // not a permutation (stale trace slot)
        // symbolic elimination along this sequence
        std::vector<unsigned> rowmask(n, 0u);
        for (int d : jfull_dst) rowmask[d & 0xffff] |= 1u << (d >> 16);
        std::vector<int> pos(n), at(n);
        for (int i = 0; i < n; ++i) pos[i] = at[i] = i;
        std::vector<StepBuild> steps(n);
        bool okp = true;
        const size_t ent_before = ent.size();
        for (int k = 0; k < n && okp; ++k) {
            const int P = piv[k];
            StepBuild &sb = steps[k];
            if (pos[P] < k || !((rowmask[P] >> k) & 1u)) {
                okp = false;
                break;
            }
            for (int r = 0; r < n; ++r)
                if (pos[r] >= k && ((rowmask[r] >> k) & 1u)) sb.cand.push_back(entry(r, k) | (r << 8) | (pos[r] << 16));
            sb.eP = entry(P, k);
            sb.prow = P;
            const unsigned cols = rowmask[P] & ~((2u << k) - 1u);
            // row swap: positions only
            const int rk = at[k], pp = pos[P];
            at[k] = P;
            at[pp] = rk;
            pos[P] = k;
            pos[rk] = pp;
            for (int r = 0; r < n; ++r) {
                if (pos[r] <= k || !((rowmask[r] >> k) & 1u)) continue;
                const int em = entry(r, k);
                sb.mult.push_back(em | (r << 8));
                for (int j = k + 1; j < n; ++j)
                    if ((cols >> j) & 1u) sb.upd.push_back(entry(r, j) | (em << 8) | (entry(P, j) << 16));
                rowmask[r] |= cols;
            }
        }
        if (!okp || ent.size() > 250) {
            // drop the fill-in this sequence added
            for (auto it = ent.begin(); it != ent.end();)
                it = (it->second >= (int)ent_before) ? ent.erase(it) : std::next(it);
            if (!okp) continue;
            break;
        }
        // entries of U per column (rows pivoted before step i that hold column i)
        for (int i = 0; i < n; ++i)
            for (int r = 0; r < n; ++r)
                if (pos[r] < i && ((rowmask[r] >> i) & 1u)) steps[i].bu.push_back(entry(r, i) | (r << 8));
        seqs.push_back(piv);
        plans.push_back(steps);
        covered += pc.second;
    }
    if (plans.empty()) return false;
    const int K = (int)plans.size();
    out.K = K;
    out.S = (int)ent.size();
    out.nnz0 = nnz0;
    out.n = n;
    out.coverage = total ? (double)covered / (double)total : 0.0;

    for (int k = 0; k < n; ++k) { // executed flops along the most frequent plan
        const StepBuild &sb = plans[0][k];
        out.flops_factor += 1 + 3 * (long long)sb.mult.size() + 2 * (long long)sb.upd.size();
        out.flops_solve += 1 + 2 * (long long)sb.bu.size();
    }
    std::vector<int> rec((size_t)K * n * 4), cand, mult, upd, bu;
    for (int p = 0; p < K; ++p)
        for (int k = 0; k < n; ++k) {
            const StepBuild &sb = plans[p][k];
            if (sb.cand.size() > 255 || sb.mult.size() > 255 || cand.size() > 60000 || mult.size() > 60000 || upd.size() > 60000 ||
                bu.size() > 60000)
                return false;
            int *w = &rec[((size_t)p * n + k) * 4];
            w[0] = sb.eP | (sb.prow << 8) | ((int)sb.cand.size() << 16) | ((int)sb.mult.size() << 24);
            w[1] = (int)cand.size() | ((int)mult.size() << 16);
            w[2] = (int)upd.size() | ((int)sb.upd.size() << 16);
            w[3] = (int)bu.size() | ((int)sb.bu.size() << 16);
            cand.insert(cand.end(), sb.cand.begin(), sb.cand.end());
            mult.insert(mult.end(), sb.mult.begin(), sb.mult.end());
            upd.insert(upd.end(), sb.upd.begin(), sb.upd.end());
            bu.insert(bu.end(), sb.bu.begin(), sb.bu.end());
        }
    // next[p][k][w]: most frequent plan q (smallest index) with the same first k pivots and pivot w at step k
    std::vector<int> next_words(((size_t)K * n * 16 + 3) / 4, 0);
    signed char *nx = reinterpret_cast<signed char *>(next_words.data());
    for (int p = 0; p < K; ++p)
        for (int k = 0; k < n; ++k)
            for (int w = 0; w < 16; ++w) {
                int found = -1;
                for (int q = 0; q < K && found < 0; ++q) {
                    if (seqs[q][k] != w) continue;
                    bool same = true;
                    for (int i = 0; i < k && same; ++i) same = seqs[q][i] == seqs[p][i];
                    if (same) found = q;
                }
                nx[((size_t)p * n + k) * 16 + w] = (signed char)found;
            }
    auto push = [&](const std::vector<int> &v) {
        int off = (int)out.ibuf.size();
        out.ibuf.insert(out.ibuf.end(), v.begin(), v.end());
        while (out.ibuf.size() & 3) out.ibuf.push_back(0); // keep 16-byte alignment of what follows
        return off;
    };
    out.off_rec = push(rec);
    out.off_cand = push(cand);
    out.off_mult = push(mult);
    out.off_upd = push(upd);
    out.off_bu = push(bu);
    out.off_next = push(next_words);
    out.off_jdst = push(jdst);
    return true;
}

} // namespace ftsb
