// chain_core.cuh -- seed chaining as lesv's mapper does it, one segment per call (host + device).
//
// Replaces, for one (read strand, contig) seed list:
//   scoring_chain_seeds        reference src/algo/chain_dp.c:104-169   (minimap2-style 1-D chaining DP)
//   chaining_find_candidates   src/algo/chain_dp.c:273-395            (best-first backtrack, <= 40 chains,
//                                                                       100-bp containment filter)
//   join_adjacent_chains       src/algo/chain_dp.c:447-534            (greedy joining, top 20 by score)
// Everything that decides an integer result is kept: strict '>' on scores (first examined predecessor
// wins), the t[]/n_skip early exit, `(int)(dd * .01 * avg_cov)` in IEEE double (two multiplies, no
// FMA), the `right->send - left->sbeg` length in two_chains_are_adjacent (:428), and the exact
// permutations of ks_introsort on partial keys (ksort_exact.cuh).
//
// The reference carries the seeds of every chain along; downstream of the DDF stage only the first
// and last seed of a (joined) chain are read (fill_init_hit_info, app/map/
// chain_and_extend_kmer_matches.c:75-95), so a hit here records its end points plus where its seed
// indices sit in v[] (for the MEM stage, which needs the lists).
#pragma once
#include "ksort_exact.cuh"

namespace lesvchain {

struct Params {
    int max_dist_ref, max_dist_qry, band_width, max_skip;
    int min_cnt, min_score;
    int max_join_long, max_join_short;
};
KSX_HD inline Params make_params(int min_cnt, int min_score)              // ChainWorkDataNew, chain_dp.c:27-58
{
    Params p;
    p.max_dist_qry = 5000; p.max_dist_ref = 5000; p.band_width = 1500; p.max_skip = 25;
    p.min_cnt = min_cnt; p.min_score = min_score;
    p.max_join_long = 20000; p.max_join_short = 2000;
    return p;
}

struct Seed { int q; int s; int len; };       // ChainSeed{qoff, soff (contig/window relative), length}

constexpr int kMaxChains = 40;                // chain_dp.c:334
constexpr int kMaxJoined = 20;                // kMaxExamineInitHit, chain_dp.c:451

struct Hit {
    int score;
    int qbeg, qend;
    long long sbeg, send;
    int v_off, v_cnt;       // chain_find: seed indices v[v_off .. v_off+v_cnt), LAST seed first
    int mem_off, mem_cnt;   // after joining: members[mem_off .. +mem_cnt) = pre-join hit ids, left to right
    int used;
};

struct IntPair { int first, second; };

KSX_HD inline int ilog2_32(uint32_t v)        // chain_dp.c:86-102 (table lookup there)
{
    int r = -1;
    while (v) { v >>= 1; ++r; }
    return r;
}

// chain_dp.c:104-169
KSX_HD inline void score_seeds(const Params& P, const Seed* seeds, int n, bool is_mem, int* f, int* p, int* t, int* v)
{
    int sum_cov = 0;
    for (int i = 0; i < n; ++i) sum_cov += seeds[i].len;
    const int avg_cov = sum_cov / n;
    for (int i = 0; i < n; ++i) { t[i] = 0; f[i] = 0; p[i] = -1; }
    int st = 0;
    for (int i = 0; i < n; ++i) {
        const long long ri = seeds[i].s, qi = seeds[i].q;
        const int cov = seeds[i].len;
        int max_j = -1, max_f = cov, n_skip = 0;
        while (st < i && ri > (long long)seeds[st].s + seeds[st].len + P.max_dist_ref) ++st;
        for (int j = i - 1; j >= st; --j) {
            if (is_mem) {
                if ((long long)seeds[j].q + seeds[j].len > qi || (long long)seeds[j].s + seeds[j].len > ri) continue;
            } else {
                if (seeds[j].q > qi || seeds[j].s > ri) continue;
            }
            const long long dr = ri - seeds[j].s, dq = qi - seeds[j].q;
            if (dr == 0 || dq <= 0) continue;
            if (dq > P.max_dist_qry || dr > P.max_dist_ref) continue;
            const int dd = (int)(dr > dq ? dr - dq : dq - dr);
            if (dd > P.band_width) continue;
            const int min_d = (int)(dq < dr ? dq : dr);
            int sc = min_d > cov ? cov : min_d;
            const int log_dd = dd ? ilog2_32((uint32_t)dd) : 0;
            if (!is_mem) {
                const double pen = dd * .01;             // two IEEE multiplies, as written in :150
                sc -= (int)(pen * avg_cov) + (log_dd >> 1);
            } else sc -= (log_dd >> 1);
            sc += f[j];
            if (sc > max_f) {
                max_f = sc; max_j = j;
                if (n_skip) --n_skip;
            } else if (t[j] == i) {
                if (++n_skip > P.max_skip) break;
            }
            if (p[j] >= 0) t[p[j]] = i;
        }
        f[i] = max_f; p[i] = max_j;
        v[i] = (max_j >= 0 && v[max_j] > max_f) ? v[max_j] : max_f;
    }
}

KSX_HD inline bool hit_contained(const Hit* ha, int hc, const Hit& hn)    // chain_dp.c:257-271
{
    const int E = 100;
    for (int i = 0; i < hc; ++i)
        if (hn.qbeg + E >= ha[i].qbeg && hn.qend <= ha[i].qend + E &&
            hn.sbeg + E >= ha[i].sbeg && hn.send <= ha[i].send + E) return true;
    return false;
}

KSX_HD inline bool chains_adjacent(const Params& P, const Hit& left, const Hit& right)   // chain_dp.c:415-444
{
    if (left.qend > right.qbeg || left.send > right.sbeg) return false;
    const int gap_qry = right.qbeg - left.qend;
    const int gap_ref = (int)(right.sbeg - left.send);
    const int max_gap = gap_qry > gap_ref ? gap_qry : gap_ref;
    const int min_gap = gap_qry < gap_ref ? gap_qry : gap_ref;
    if (max_gap > P.max_join_long || min_gap > P.max_join_short) return false;
    const int left_qlen = left.qend - left.qbeg;
    const int left_slen = (int)(left.send - left.sbeg);
    const int right_qlen = right.qend - right.qbeg;
    const int right_slen = (int)(right.send - left.sbeg);                 // sic (:428)
    if (left_qlen < 1000 || right_qlen < 1000 || left_slen < 1000 || right_slen < 1000) return false;
    if (left.score < 500 || right.score < 500) return false;
    return true;
}

struct HitSoffLt {        // init_hit_soff_lt, chain_dp.c:408
    KSX_HD bool operator()(const Hit& a, const Hit& b) const { return a.sbeg < b.sbeg || (a.sbeg == b.sbeg && a.qbeg < b.qbeg); }
};
struct HitScoreGt {       // init_hit_score_gt, corelib/gapped_candidate.c:5
    KSX_HD bool operator()(const Hit& a, const Hit& b) const { return a.score > b.score; }
};
struct PairLt {           // int_pair_lt, corelib/hbn_aux.c:21
    KSX_HD bool operator()(const IntPair& a, const IntPair& b) const { return a.first < b.first || (a.first == b.first && a.second < b.second); }
};
struct PairGt {           // intpair_gt, chain_dp.c:411
    KSX_HD bool operator()(const IntPair& a, const IntPair& b) const { return a.first > b.first || (a.first == b.first && a.second < b.second); }
};

// join_adjacent_chains, chain_dp.c:447-534.  hits[0..nh) -> out[0..ret), members[] filled.
KSX_HD inline int join_chains(const Params& P, Hit* hits, int nh, Hit* out, int* members)
{
    lesvsort::introsort(hits, (size_t)nh, HitSoffLt());
    IntPair order[kMaxChains];
    for (int i = 0; i < nh; ++i) { order[i].first = hits[i].score; order[i].second = i; hits[i].used = 0; }
    lesvsort::introsort(order, (size_t)nh, PairGt());
    int n_out = 0, n_mem = 0;
    for (int i = 0; i < nh && i < kMaxJoined; ++i) {
        const int hit_idx = order[i].second;
        if (hits[hit_idx].used) continue;
        Hit nw = hits[hit_idx];
        hits[hit_idx].used = 1;
        Hit chain[kMaxChains];                      // copies, in the order the reference pushes them
        int chain_id[kMaxChains];
        int nc = 0;
        chain[nc] = nw; chain_id[nc] = hit_idx; ++nc;
        int k = hit_idx;
        while (k) {
            --k;
            if (hits[k].used) continue;
            if (!chains_adjacent(P, hits[k], nw)) continue;
            chain[nc] = hits[k]; chain_id[nc] = k; ++nc;
            nw.qbeg = hits[k].qbeg; nw.sbeg = hits[k].sbeg; nw.score += hits[k].score;
            hits[k].used = 1;
        }
        k = hit_idx;
        while (k < nh - 1) {
            ++k;
            if (hits[k].used) continue;
            if (!chains_adjacent(P, nw, hits[k])) continue;
            chain[nc] = hits[k]; chain_id[nc] = k; ++nc;
            nw.qend = hits[k].qend; nw.send = hits[k].send; nw.score += hits[k].score;
            hits[k].used = 1;
        }
        // the member order is the introsort permutation of the pushed copies; carry the ids along
        for (int x = 0; x < nc; ++x) chain[x].used = chain_id[x];
        lesvsort::introsort(chain, (size_t)nc, HitSoffLt());
        nw.mem_off = n_mem; nw.mem_cnt = nc;
        for (int x = 0; x < nc; ++x) members[n_mem++] = chain[x].used;
        out[n_out++] = nw;
    }
    lesvsort::introsort(out, (size_t)n_out, HitScoreGt());
    return n_out;
}

KSX_HD inline int find_candidates_scored(const Params& P, const Seed* seeds, int n,
                                         int* f, int* p, int* t, int* v, IntPair* u,
                                         Hit* pre, int* n_pre, Hit* out, int* members);

// chaining_find_candidates, chain_dp.c:273-395.  Scratch f,p,t,v: n ints each; u: n pairs.
// pre[] (kMaxChains) receives the chains before joining (members[] refers to them), out[] (kMaxJoined)
// the final hits in the reference's order.  Returns the number of final hits; *n_pre = chains found.
KSX_HD inline int find_candidates(const Params& P, const Seed* seeds, int n, bool is_mem,
                                  int* f, int* p, int* t, int* v, IntPair* u,
                                  Hit* pre, int* n_pre, Hit* out, int* members)
{
    *n_pre = 0;
    if (n == 0) return 0;
    score_seeds(P, seeds, n, is_mem, f, p, t, v);
    return find_candidates_scored(P, seeds, n, f, p, t, v, u, pre, n_pre, out, members);
}

// chaining_find_candidates after scoring_chain_seeds has filled f, p, v (chain_dp.c:286-395).
KSX_HD inline int find_candidates_scored(const Params& P, const Seed* seeds, int n,
                                         int* f, int* p, int* t, int* v, IntPair* u,
                                         Hit* pre, int* n_pre, Hit* out, int* members)
{
    *n_pre = 0;
    if (n == 0) return 0;
    for (int i = 0; i < n; ++i) t[i] = 0;
    for (int i = 0; i < n; ++i) if (p[i] >= 0) t[p[i]] = 1;
    int n_u = 0;
    for (int i = 0; i < n; ++i)
        if (t[i] == 0 && v[i] >= P.min_score) {
            int j = i;
            while (j >= 0 && f[j] < v[j]) j = p[j];     // the peak of the chain ending at i
            if (j < 0) j = i;
            u[n_u].first = f[j]; u[n_u].second = j;
            ++n_u;
        }
    if (n_u == 0) return 0;
    lesvsort::introsort(u, (size_t)n_u, PairLt());
    for (int i = 0; i < n_u >> 1; ++i) { const IntPair x = u[i]; u[i] = u[n_u - i - 1]; u[n_u - i - 1] = x; }

    for (int i = 0; i < n; ++i) t[i] = 0;
    int n_v = 0, k = 0, added = 0;
    for (int i = 0; i < n_u && added < kMaxChains; ++i) {
        const int n_v0 = n_v, k0 = k;
        int j = u[i].second;
        if (t[j]) continue;
        bool found = false;
        do { v[n_v++] = j; t[j] = 1; j = p[j]; } while (j >= 0 && t[j] == 0);
        if (j < 0) {
            if (n_v - n_v0 >= P.min_cnt) { u[k].first = u[i].first; u[k].second = n_v - n_v0; ++k; found = true; }
        } else if (u[i].first - f[j] >= P.min_score) {
            if (n_v - n_v0 >= P.min_cnt) { u[k].first = u[i].first - f[j]; u[k].second = n_v - n_v0; ++k; found = true; }
        }
        if (found) {
            Hit h;
            h.score = u[k - 1].first;
            h.v_off = n_v0; h.v_cnt = n_v - n_v0;
            const Seed& a = seeds[v[n_v - 1]];
            const Seed& b = seeds[v[n_v0]];
            h.qbeg = a.q; h.sbeg = a.s; h.qend = b.q + b.len; h.send = (long long)b.s + b.len;
            h.mem_off = 0; h.mem_cnt = 0; h.used = 0;
            if (!hit_contained(pre, added, h)) pre[added++] = h;
        }
        if (k0 == k) n_v = n_v0;
    }
    *n_pre = added;
    if (added < 2) {                              // join_adjacent_chains returns at once (:449)
        for (int i = 0; i < added; ++i) { out[i] = pre[i]; out[i].mem_off = i; out[i].mem_cnt = 1; members[i] = i; }
        return added;
    }
    return join_chains(P, pre, added, out, members);
}

}  // namespace lesvchain
