// sched.h -- host-side scheduling of shared-memory gathers (no device code).
//
// The AMP kernel's two random gathers (fold of z into the M bins of a section, gather of FHT_M(beta_l) into the
// n rows of A beta) are bound by the shared-memory data pipe: 32 lanes reading 32 random 4-byte words need ~2.7
// wavefronts instead of 1.  With fixed-point (integer) operands the ORDER in which a lane adds its terms is
// free, so the order is chosen at table-build time such that the 32 lanes of a warp hit distinct banks in every
// step.  One "pool" = the entries one warp reads over K consecutive steps: lane q has <= K entries, each with a
// bank.  This is an edge colouring of the bipartite multigraph lanes x banks with K colours in which every
// vertex sees every colour floor(d/K) or ceil(d/K) times (equitable colouring; de Werra), built by recursive
// Euler splits (K is a power of two).  A lane therefore gets each step at most once, a bank of degree <= K is
// hit at most once per step, and a bank of degree d > K at most ceil(d/K) times.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <functional>
#include <vector>

namespace sb {

struct PoolEdge {
    int lane, bank;  // bank in [0, 32)
    int id;          // caller's entry id
    int step;        // out: step in [0, K)
};

class PoolScheduler {
  public:
    // colours `edges` with K (power of two) steps
    void run(std::vector<PoolEdge> &edges, int K) {
        idx_.resize(edges.size());
        for (size_t i = 0; i < edges.size(); i++) idx_[i] = (int)i;
        rec(edges, idx_.data(), (int)edges.size(), K, 0);
    }

    // cost model: sum over steps of the maximum bank multiplicity (= shared-memory wavefronts of the pool);
    // steps without any entry cost 1 as well (the instruction is still issued)
    static int cost(const std::vector<PoolEdge> &edges, int K) {
        std::vector<int> cnt((size_t)K * 32, 0);
        for (const PoolEdge &e : edges) cnt[(size_t)e.step * 32 + e.bank]++;
        int total = 0;
        for (int t = 0; t < K; t++) {
            int m = 1;
            for (int b = 0; b < 32; b++) m = cnt[(size_t)t * 32 + b] > m ? cnt[(size_t)t * 32 + b] : m;
            total += m;
        }
        return total;
    }

    // Local improvement: a step costs max multiplicity m_t.  Moving a lane's entry from step t to step t' means
    // swapping the lane's two entries (or its entry and an idle slot).  Accept swaps that lower the total cost.
    static void improve(std::vector<PoolEdge> &edges, int K, int rounds = 4) {
        const int NL = 32;
        std::vector<int> at((size_t)NL * K, -1);  // at[lane][step] = edge index
        std::vector<int> cnt((size_t)K * 32, 0);
        for (size_t i = 0; i < edges.size(); i++) {
            at[(size_t)edges[i].lane * K + edges[i].step] = (int)i;
            cnt[(size_t)edges[i].step * 32 + edges[i].bank]++;
        }
        auto stepmax = [&](int t) {
            int m = 1;
            for (int b = 0; b < 32; b++) m = cnt[(size_t)t * 32 + b] > m ? cnt[(size_t)t * 32 + b] : m;
            return m;
        };
        std::vector<int> smax(K);
        for (int t = 0; t < K; t++) smax[t] = stepmax(t);
        for (int r = 0; r < rounds; r++) {
            bool any = false;
            for (int t = 0; t < K; t++) {
                if (smax[t] <= 1) continue;
                // try to empty the conflicts of step t into other steps without raising their maximum
                for (size_t i = 0; i < edges.size() && smax[t] > 1; i++) {
                    PoolEdge &e = edges[i];
                    if (e.step != t || cnt[(size_t)t * 32 + e.bank] < smax[t]) continue;
                    for (int t2 = 0; t2 < K; t2++) {
                        if (t2 == t) continue;
                        const int j = at[(size_t)e.lane * K + t2];  // the lane's entry at t2 (or idle)
                        const int bj = j >= 0 ? edges[j].bank : -1;
                        // e moves to t2: bank count there must stay <= smax[t2] (no cost increase), and strictly
                        // below it unless t2 is already as bad
                        const int c2 = cnt[(size_t)t2 * 32 + e.bank] - (bj == e.bank ? 1 : 0) + 1;
                        if (c2 > smax[t2]) continue;
                        // j moves to t: its bank count at t must stay below the current maximum
                        if (j >= 0) {
                            const int c1 = cnt[(size_t)t * 32 + bj] - (bj == e.bank ? 1 : 0) + 1;
                            if (c1 >= smax[t]) continue;
                        }
                        cnt[(size_t)t * 32 + e.bank]--;
                        cnt[(size_t)t2 * 32 + e.bank]++;
                        at[(size_t)e.lane * K + t2] = (int)i;
                        at[(size_t)e.lane * K + t] = j;
                        if (j >= 0) {
                            cnt[(size_t)t2 * 32 + bj]--;
                            cnt[(size_t)t * 32 + bj]++;
                            edges[j].step = t;
                        }
                        e.step = t2;
                        any = true;
                        break;
                    }
                    const int m = stepmax(t);
                    smax[t] = m;
                }
            }
            if (!any) break;
        }
    }

    // Second heuristic: fill the steps one after the other, each with the smallest bank multiplicity c that still
    // serves every lane that must read in this step (a lane with as many terms left as steps left), found by
    // augmenting paths (lanes x banks with capacity c, banks with the most terms left first); lanes with slack join
    // as long as c does not grow.  Early steps become conflict-free (perfect matchings take one term off EVERY bank),
    // the excess of the heavy banks collects in the few last steps: sum_t c_t approaches its lower bound max_b
    // degree(b), where the equitable colouring pays ceil(d/K) in (nearly) every step.
    void run_greedy(std::vector<PoolEdge> &E, int K) {
        const int m = (int)E.size();
        int reml[32] = {0}, remb[32] = {0}, match[32], order[32], nat[32];
        alive_.assign(m, 1);
        for (const PoolEdge &e : E) { reml[e.lane]++; remb[e.bank]++; }
        for (auto &v : adj_) v.clear();
        for (int i = 0; i < m; i++) adj_[E[i].lane].push_back(i);
        for (int t = 0; t < K; t++) {
            const int left = K - t;
            for (int l = 0; l < 32; l++) {  // this step's candidate edges of every lane, banks with the most terms left first
                cand_[l].clear();
                for (int i : adj_[l]) if (alive_[i]) cand_[l].push_back(i);
                std::sort(cand_[l].begin(), cand_[l].end(), [&](int x, int y) { return remb[E[x].bank] > remb[E[y].bank]; });
                order[l] = l;
            }
            std::sort(order, order + 32, [&](int x, int y) { return reml[x] > reml[y]; });
            for (cap_ = 1;; cap_++) {
                for (int l = 0; l < 32; l++) { match[l] = -1; nat[l] = 0; }
                bool ok = true;
                for (int oi = 0; oi < 32 && ok; oi++) {
                    const int l = order[oi];
                    if (reml[l] < left) break;  // (sorted: the rest has slack)
                    vis_ = 0;
                    ok = aug(E, l, match, nat);
                }
                if (ok) break;
            }
            for (int oi = 0; oi < 32; oi++) {  // lanes with slack: as many as fit without raising the multiplicity
                const int l = order[oi];
                if (reml[l] >= left || reml[l] == 0) continue;
                vis_ = 0;
                aug(E, l, match, nat);
            }
            for (int l = 0; l < 32; l++) {
                const int i = match[l];
                if (i < 0) continue;
                E[i].step = t;
                alive_[i] = 0;
                reml[l]--;
                remb[E[i].bank]--;
            }
        }
    }

    // the cheaper of the two schedules under the wavefront model
    void run_best(std::vector<PoolEdge> &E, int K) {
        run(E, K);
        improve(E, K);
        const int c1 = cost(E, K);
        std::vector<PoolEdge> G = E;
        run_greedy(G, K);
        if (cost(G, K) < c1) E.swap(G);
    }

  private:
    // augmenting path for lane u: banks hold up to cap_ lanes (at_[b][0..nat[b]))
    bool aug(const std::vector<PoolEdge> &E, int u, int *match, int *nat) {
        for (int i : cand_[u]) {
            const int b = E[i].bank;
            if (vis_ & (1u << b)) continue;
            vis_ |= 1u << b;
            if (nat[b] < cap_) { at_[b][nat[b]++] = u; match[u] = i; return true; }
            for (int k = 0; k < nat[b]; k++) {
                const int w = at_[b][k];
                if (aug(E, w, match, nat)) {  // w moved to another bank (its entry in at_[b] is still at index k)
                    at_[b][k] = u;
                    match[u] = i;
                    return true;
                }
            }
        }
        return false;
    }
    std::vector<int> alive_, adj_[32], cand_[32];
    int at_[32][32];
    unsigned vis_ = 0;
    int cap_ = 1;
    std::vector<int> idx_, tmp_, label_, head_, nxt_, other_;
    std::vector<char> used_;

    // vertices: lanes 0..31, banks 32..63
    void rec(std::vector<PoolEdge> &E, int *ids, int m, int K, int base) {
        if (K == 1 || m == 0) {
            for (int i = 0; i < m; i++) E[ids[i]].step = base;
            return;
        }
        // Euler split of the m edges into two halves balanced at every vertex
        const int NV = 64;
        int deg[NV] = {0};
        // adjacency as linked lists over half-edges: half-edge 2*i (at lane), 2*i+1 (at bank)
        head_.assign(NV, -1);
        nxt_.assign((size_t)2 * m, -1);
        used_.assign(m, 0);
        label_.assign(m, 0);
        for (int i = 0; i < m; i++) {
            const PoolEdge &e = E[ids[i]];
            const int u = e.lane, v = 32 + e.bank;
            nxt_[2 * i] = head_[u]; head_[u] = 2 * i;
            nxt_[2 * i + 1] = head_[v]; head_[v] = 2 * i + 1;
            deg[u]++; deg[v]++;
        }
        auto walk = [&](int start) {
            int v = start, lab = 0;
            // alternate the first label between walks so that the surplus of odd vertices is spread
            lab = (walk_parity_ ^= 1);
            for (;;) {
                int h = head_[v];
                while (h >= 0 && used_[h >> 1]) h = nxt_[h];
                head_[v] = h;
                if (h < 0) break;
                const int i = h >> 1;
                used_[i] = 1;
                label_[i] = lab;
                lab ^= 1;
                const PoolEdge &e = E[ids[i]];
                const int u = e.lane, w = 32 + e.bank;
                deg[u]--; deg[w]--;
                v = (v == u) ? w : u;
            }
        };
        for (int v = 0; v < NV; v++)
            while (deg[v] & 1) walk(v);
        for (int v = 0; v < NV; v++)
            while (deg[v] > 0) walk(v);
        // partition ids in place: label 0 first
        tmp_.resize(m);
        int n0 = 0;
        for (int i = 0; i < m; i++) if (label_[i] == 0) tmp_[n0++] = ids[i];
        int p = n0;
        for (int i = 0; i < m; i++) if (label_[i] == 1) tmp_[p++] = ids[i];
        for (int i = 0; i < m; i++) ids[i] = tmp_[i];
        rec(E, ids, n0, K / 2, base);
        rec(E, ids + n0, m - n0, K / 2, base + K / 2);
    }
    int walk_parity_ = 0;
};

}  // namespace sb
