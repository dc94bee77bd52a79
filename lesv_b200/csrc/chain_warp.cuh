// chain_warp.cuh -- scoring_chain_seeds (reference src/algo/chain_dp.c:104-169) computed by one warp, with
// the results of the sequential loop (chain_core.cuh: score_seeds), bit for bit.
//
// For seed i the reference scans its predecessors j = i-1 .. st one by one: a strict '>' keeps the first
// best predecessor, t[] marks predecessors of predecessors, and the scan stops once more than max_skip
// marked, non-improving predecessors have been seen (every improvement pays one skip back).  Here 32
// predecessors are scored at once, one per lane (the loads of seeds[j], f[j], p[j], t[j] are one
// coalesced round trip instead of 32 dependent ones), and the order-dependent part is replayed exactly:
//   * lane l improves the running best  <=>  its score beats max(best so far, scores of lanes < l): a prefix max;
//   * t[j] == i also holds if an earlier lane of the same batch has j as its predecessor: one OR-reduction;
//   * the skip counter is walked over the (few) lanes that change it, giving the lane where the scan stops;
//     lanes behind it are discarded, including their t[] marks.
#pragma once
#include "chain_core.cuh"

namespace lesvchain {

__device__ inline void score_seeds_warp(const Params& P, const Seed* seeds, int n, bool is_mem, int* f, int* p, int* t, int* v)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    int sum_cov = 0;
    for (int i = lane; i < n; i += 32) { sum_cov += seeds[i].len; t[i] = 0; f[i] = 0; p[i] = -1; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum_cov += __shfl_xor_sync(kFull, sum_cov, d);
    const int avg_cov = sum_cov / n;
    __syncwarp();
    int st = 0;
    for (int i = 0; i < n; ++i) {
        const Seed si = seeds[i];
        const long long ri = si.s, qi = si.q;
        const int cov = si.len;
        int max_j = -1, max_f = cov, n_skip = 0;
        while (st < i && ri > (long long)seeds[st].s + seeds[st].len + P.max_dist_ref) ++st;
        bool stop = false;
        for (int jhi = i - 1; jhi >= st && !stop; jhi -= 32) {
            const int j = jhi - lane;
            bool valid = j >= st;
            int sc = 0, pj = -1;
            bool marked = false;
            if (valid) {
                const Seed sj = seeds[j];
                if (is_mem) { if ((long long)sj.q + sj.len > qi || (long long)sj.s + sj.len > ri) valid = false; }
                else { if (sj.q > qi || sj.s > ri) valid = false; }
                const long long dr = ri - sj.s, dq = qi - sj.q;
                if (dr == 0 || dq <= 0) valid = false;
                if (dq > P.max_dist_qry || dr > P.max_dist_ref) valid = false;
                const int dd = (int)(dr > dq ? dr - dq : dq - dr);
                if (dd > P.band_width) valid = false;
                if (valid) {
                    const int min_d = (int)(dq < dr ? dq : dr);
                    sc = min_d > cov ? cov : min_d;
                    const int log_dd = dd ? ilog2_32((uint32_t)dd) : 0;
                    if (!is_mem) {
                        const double pen = dd * .01;             // two IEEE multiplies, as written in :150
                        sc -= (int)(pen * avg_cov) + (log_dd >> 1);
                    } else sc -= (log_dd >> 1);
                    sc += f[j];
                    pj = p[j];
                    marked = t[j] == i;
                }
            }
            // marks set by earlier lanes of this batch (t[p[j']] = i with p[j'] inside the batch)
            const unsigned tgt = (valid && pj >= 0 && pj > jhi - 32) ? (1u << (jhi - pj)) : 0u;
            const unsigned in_batch = __reduce_or_sync(kFull, tgt);
            marked = marked || ((in_batch >> lane) & 1u);
            // who improves the running best: exclusive prefix max
            int x = valid ? sc : INT_MIN;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(kFull, x, d); if (lane >= d && y > x) x = y; }
            int excl = __shfl_up_sync(kFull, x, 1);
            if (lane == 0 || excl < max_f) excl = max_f;
            const bool upd = valid && sc > excl;
            const unsigned um = __ballot_sync(kFull, upd);
            const unsigned im = __ballot_sync(kFull, valid && !upd && marked);
            int brk = 32;
            for (unsigned ev = um | im; ev; ev &= ev - 1) {
                const int l = __ffs(ev) - 1;
                if ((um >> l) & 1u) { if (n_skip) --n_skip; }
                else if (++n_skip > P.max_skip) { brk = l; break; }
            }
            const unsigned before = brk >= 32 ? kFull : ((1u << brk) - 1u);
            const unsigned um2 = um & before;
            if (um2) {
                const int l = 31 - __clz(um2);
                max_f = __shfl_sync(kFull, sc, l);
                max_j = jhi - l;
            }
            if (valid && lane < brk && pj >= 0) t[pj] = i;
            if (brk < 32) stop = true;
            __syncwarp();
        }
        if (lane == 0) {
            f[i] = max_f; p[i] = max_j;
            v[i] = (max_j >= 0 && v[max_j] > max_f) ? v[max_j] : max_f;
        }
        __syncwarp();
    }
}

}  // namespace lesvchain
