// trace_warp.cuh -- edlib_nw (trace_core.cuh: nw_path) computed by one warp.
//
// Same results as lesvtb::nw_path, bit for bit: the same Myers recurrence (calculateBlock, reference
// src/algo/edlib.cpp:393-432), the same stored columns and walk (myers_walk), the same Hirschberg splits
// (obtainAlignmentHirschberg, :1218-1383).  Only the schedule differs: lane b owns the 64-row block b of the
// query (blocks 32 s + b in strip s for queries longer than 2,048 bases) and the columns run through the
// lanes as a wavefront -- lane b works on column t - b at step t and hands its horizontal delta to lane b+1
// with one __shfl_up_sync.  Between strips the deltas of the last lane are parked in a byte per column.
// The block's four Peq words live in registers.
#pragma once
#include "trace_core.cuh"

namespace lesvtw {

using lesvtb::Scratch;
using lesvtb::kOpDel;
using lesvtb::kOpIns;
using lesvtb::kNoAlign;
using lesvtb::kOpsOverflow;
using lesvtb::kScratchOverflow;

constexpr unsigned kFull = 0xFFFFFFFFu;

// Peq words of block b for the four letters (buildPeq, edlib.cpp:330-357: rows past the query match anything)
template <class FQ>
__device__ __forceinline__ void build_eq(FQ Q, int n, int b, uint64_t eq[4])
{
    eq[0] = eq[1] = eq[2] = eq[3] = 0;
    const int i0 = b << 6;
    const int cnt = n - i0 < 64 ? n - i0 : 64;
    for (int r = 0; r < cnt; ++r) {
        const uint32_t c = Q(i0 + r);
        const uint64_t bit = 1ull << r;
        eq[0] |= c == 0 ? bit : 0; eq[1] |= c == 1 ? bit : 0; eq[2] |= c == 2 ? bit : 0; eq[3] |= c == 3 ? bit : 0;
    }
    if (cnt < 64) { const uint64_t pad = ~0ull << cnt; eq[0] |= pad; eq[1] |= pad; eq[2] |= pad; eq[3] |= pad; }
}

// Columns 1..m of Q[0..n) x T[0..m).  STORE: every column goes to sc.P / sc.M / sc.score in myers_fill's
// layout (block-major).  !STORE: only the last column survives, returned as D[0..n] (myers_last_column).
// carry: >= m bytes of per-warp scratch (used when n > 2048).
template <bool STORE, class FQ, class FT>
__device__ inline bool myers_warp(FQ Q, int n, FT T, int m, Scratch& sc, int8_t* carry, int* D)
{
    const int lane = threadIdx.x & 31;
    const int nb = (n + 63) >> 6;
    if (STORE && (size_t)nb * (size_t)m > sc.cap_cells) return false;
    const int n_strips = (nb + 31) >> 5;
    if (STORE) sc.cols = m;
    if (!STORE && lane == 0) D[0] = m;
    for (int s = 0; s < n_strips; ++s) {
        const int b = (s << 5) + lane;
        const bool active = b < nb;
        uint64_t eq[4];
        if (active) build_eq(Q, n, b, eq);
        uint64_t Pv = ~0ull, Mv = 0ull;
        int score = (b + 1) * 64;
        int hout = 0;
        const bool last_lane_of_strip = lane == 31 && s + 1 < n_strips;
        const int steps = m + 31;
        for (int t = 0; t < steps; ++t) {
            const int up = __shfl_up_sync(kFull, hout, 1);
            const int j = t - lane;
            if (active && j >= 0 && j < m) {
                int hin;
                if (lane == 0) hin = s == 0 ? 1 : (int)carry[j];
                else hin = up;
                const uint32_t c = T(j);
                const uint64_t e = c == 0 ? eq[0] : c == 1 ? eq[1] : c == 2 ? eq[2] : eq[3];
                hout = lesvtb::myers_block(Pv, Mv, e, hin, Pv, Mv);
                score += hout;
                if (STORE) {
                    const size_t x = (size_t)b * m + j;
                    sc.P[x] = Pv; sc.M[x] = Mv; sc.score[x] = score;
                }
                if (last_lane_of_strip) carry[j] = (int8_t)hout;
            }
        }
        if (!STORE && active) {                         // rows of this block, from the block's top cell downwards
            int d = score - lesvtb::popc64(Pv) + lesvtb::popc64(Mv);
            const int i0 = b << 6;
            const int cnt = n - i0 < 64 ? n - i0 : 64;
            for (int r = 0; r < cnt; ++r) {
                d += (int)((Pv >> r) & 1) - (int)((Mv >> r) & 1);
                D[i0 + r + 1] = d;
            }
        }
        __syncwarp();
    }
    __syncwarp();
    return true;
}

// nw_path by one warp.  All lanes must call it with the same arguments; every lane returns the same value.
// out: n + m bytes.  sc: this warp's slab (cap_cells >= 52,500 and >= 2 (n + 1)); carry: >= m bytes.
template <class FQ, class FT>
__device__ inline int nw_path_warp(FQ Q, int n, FT T, int m, Scratch& sc, int8_t* carry, uint8_t* out, size_t out_cap)
{
    const int lane = threadIdx.x & 31;
    if (n == 0 || m == 0) return 0;
    if ((size_t)n + (size_t)m > out_cap) return kOpsOverflow;
    if (!lesvtb::needs_hirschberg(n, m)) {
        if (!myers_warp<true>(Q, n, T, m, sc, carry, (int*)nullptr)) return kScratchOverflow;
        int len = 0;
        if (lane == 0) len = lesvtb::myers_walk(sc, (n + 63) >> 6, n, m, out);
        __syncwarp();
        return __shfl_sync(kFull, len, 0);
    }
    if (2 * ((size_t)n + 1) > sc.cap_cells) return kScratchOverflow;
    if (!myers_warp<false>(Q, n, T, m, sc, carry, sc.score)) return kScratchOverflow;
    const int best = sc.score[n];
    __syncwarp();
    // obtainAlignment / obtainAlignmentHirschberg with an explicit stack (same frames on every lane)
    struct Frame { int q0, n, t0, m, best; };
    Frame stack[48];
    int top = 0;
    size_t len = 0;
    stack[top++] = Frame{0, n, 0, m, best};
    while (top > 0) {
        const Frame f = stack[--top];
        if (len + (size_t)f.n + (size_t)f.m > out_cap) return kOpsOverflow;
        if (f.n == 0 || f.m == 0) {
            const uint8_t op = f.n == 0 ? kOpDel : kOpIns;
            for (int x = lane; x < f.n + f.m; x += 32) out[len + x] = op;
            len += (size_t)(f.n + f.m);
            __syncwarp();
            continue;
        }
        auto fq = [&](int i) { return Q(f.q0 + i); };
        auto ft = [&](int j) { return T(f.t0 + j); };
        if (!lesvtb::needs_hirschberg(f.n, f.m)) {
            if (!myers_warp<true>(fq, f.n, ft, f.m, sc, carry, (int*)nullptr)) return kScratchOverflow;
            int l = 0;
            if (lane == 0) l = lesvtb::myers_walk(sc, (f.n + 63) >> 6, f.n, f.m, out + len);
            __syncwarp();
            len += (size_t)__shfl_sync(kFull, l, 0);
            continue;
        }
        if (2 * ((size_t)f.n + 1) > sc.cap_cells || top + 2 > 48) return kScratchOverflow;
        const int lhw = f.m / 2, rhw = f.m - lhw;
        int* L = sc.score;
        int* R = sc.score + f.n + 1;
        myers_warp<false>(fq, f.n, ft, lhw, sc, carry, L);
        auto rq = [&](int i) { return Q(f.q0 + f.n - 1 - i); };
        auto rt = [&](int j) { return T(f.t0 + f.m - 1 - j); };
        myers_warp<false>(rq, f.n, rt, rhw, sc, carry, R);
        // first interior row whose halves add up (edlib.cpp:1300-1310), then the two border cases
        int row = -2;
        for (int r0 = 0; r0 <= f.n - 2 && row == -2; r0 += 32) {
            const int r = r0 + lane;
            const bool hit = r <= f.n - 2 && L[r + 1] + R[f.n - r - 1] == f.best;
            const unsigned ballot = __ballot_sync(kFull, hit);
            if (ballot) row = r0 + __ffs(ballot) - 1;
        }
        int ls = 0, rs = 0;
        if (row >= 0) { ls = L[row + 1]; rs = R[f.n - row - 1]; }
        if (row == -2 && lhw + R[f.n] == f.best) { row = -1; ls = lhw; rs = R[f.n]; }
        if (row == -2 && L[f.n] + rhw == f.best) { row = f.n - 1; ls = L[f.n]; rs = rhw; }
        __syncwarp();
        if (row == -2) return kNoAlign;
        const int ul_h = row + 1;
        stack[top++] = Frame{f.q0 + ul_h, f.n - ul_h, f.t0 + lhw, rhw, rs};
        stack[top++] = Frame{f.q0, ul_h, f.t0, lhw, ls};
    }
    return (int)len;
}

// ---- the same, by all warps of a thread block (giant gaps) ---------------------------------------------------
// The strips of myers_warp do not have to run one after the other: strip s only needs, for column j, the
// delta strip s-1 left for that column.  Warp w takes strips w, w + NW, ...; a strip publishes how many columns
// it has finished (shared memory, every 32 columns) and its successor waits for the 32 columns it is about to
// read.  Same arithmetic, same results; a 16 k x 11 k Hirschberg problem finishes ~NW times sooner.
// carry: NW buffers of carry_stride >= m bytes; progress: kMaxStrips ints of shared memory.
constexpr int kMaxStrips = 512;                       // queries up to 2^20 bases

template <int NW, class FQ, class FT>
__device__ inline void myers_cta_last_column(FQ Q, int n, FT T, int m, int8_t* carry, size_t carry_stride,
                                             volatile int* progress, int* D)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = (n + 63) >> 6;
    const int n_strips = (nb + 31) >> 5;
    for (int s = threadIdx.x; s < n_strips; s += NW * 32) progress[s] = 0;
    if (threadIdx.x == 0) D[0] = m;
    __syncthreads();
    for (int s = warp; s < n_strips; s += NW) {
        const int b = (s << 5) + lane;
        const bool active = b < nb;
        uint64_t eq[4];
        if (active) build_eq(Q, n, b, eq);
        uint64_t Pv = ~0ull, Mv = 0ull;
        int score = (b + 1) * 64;
        int hout = 0;
        const bool publish = s + 1 < n_strips;
        const int8_t* cin = carry + (size_t)((s + NW - 1) % NW) * carry_stride;
        int8_t* cout = carry + (size_t)(s % NW) * carry_stride;
        const int steps = m + 31;
        for (int t = 0; t < steps; ++t) {
            if (s > 0 && (t & 31) == 0 && t < m) {               // the next 32 columns of the strip above must be there
                const int need = t + 32 < m ? t + 32 : m;
                while (progress[s - 1] < need) __nanosleep(64);
                __threadfence_block();
                __syncwarp();
            }
            const int up = __shfl_up_sync(kFull, hout, 1);
            const int j = t - lane;
            if (active && j >= 0 && j < m) {
                int hin;
                if (lane == 0) hin = s == 0 ? 1 : (int)*(const volatile int8_t*)(cin + j);
                else hin = up;
                const uint32_t c = T(j);
                const uint64_t e = c == 0 ? eq[0] : c == 1 ? eq[1] : c == 2 ? eq[2] : eq[3];
                hout = lesvtb::myers_block(Pv, Mv, e, hin, Pv, Mv);
                score += hout;
                if (publish && lane == 31) {
                    cout[j] = (int8_t)hout;
                    if ((j & 31) == 31 || j == m - 1) { __threadfence_block(); progress[s] = j + 1; }
                }
            }
        }
        if (active) {
            int d = score - lesvtb::popc64(Pv) + lesvtb::popc64(Mv);
            const int i0 = b << 6;
            const int cnt = n - i0 < 64 ? n - i0 : 64;
            for (int r = 0; r < cnt; ++r) {
                d += (int)((Pv >> r) & 1) - (int)((Mv >> r) & 1);
                D[i0 + r + 1] = d;
            }
        }
        __syncwarp();
    }
    __syncthreads();
}

// nw_path by a whole thread block (NW warps).  All threads call it with the same arguments and get the same value.
// sc: ONE slab for the block (cap_cells >= 52,500 and >= 2 (n + 1)).
template <int NW, class FQ, class FT>
__device__ inline int nw_path_cta(FQ Q, int n, FT T, int m, Scratch& sc, int8_t* carry, size_t carry_stride,
                                  volatile int* progress, int* sh_len, uint8_t* out, size_t out_cap)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (n == 0 || m == 0) return 0;
    if ((size_t)n + (size_t)m > out_cap) return kOpsOverflow;
    if (2 * ((size_t)n + 1) > sc.cap_cells || ((n + 63) >> 6) > 32 * kMaxStrips) return kScratchOverflow;
    auto leaf = [&](auto fq, int ln, auto ft, int lm, uint8_t* dst) -> int {       // stored columns + walk: warp 0
        if (warp == 0) {
            int l = kScratchOverflow;
            if (myers_warp<true>(fq, ln, ft, lm, sc, carry, (int*)nullptr)) {
                l = 0;
                if (lane == 0) l = lesvtb::myers_walk(sc, (ln + 63) >> 6, ln, lm, dst);
                l = __shfl_sync(kFull, l, 0);
            }
            if (lane == 0) *sh_len = l;
        }
        __syncthreads();
        const int l = *sh_len;
        __syncthreads();
        return l;
    };
    if (!lesvtb::needs_hirschberg(n, m)) return leaf(Q, n, T, m, out);
    myers_cta_last_column<NW>(Q, n, T, m, carry, carry_stride, progress, sc.score);
    const int best = sc.score[n];
    __syncthreads();
    struct Frame { int q0, n, t0, m, best; };
    Frame stack[48];
    int top = 0;
    size_t len = 0;
    stack[top++] = Frame{0, n, 0, m, best};
    while (top > 0) {
        const Frame f = stack[--top];
        if (len + (size_t)f.n + (size_t)f.m > out_cap) return kOpsOverflow;
        if (f.n == 0 || f.m == 0) {
            const uint8_t op = f.n == 0 ? kOpDel : kOpIns;
            for (int x = threadIdx.x; x < f.n + f.m; x += NW * 32) out[len + x] = op;
            len += (size_t)(f.n + f.m);
            __syncthreads();
            continue;
        }
        auto fq = [&](int i) { return Q(f.q0 + i); };
        auto ft = [&](int j) { return T(f.t0 + j); };
        if (!lesvtb::needs_hirschberg(f.n, f.m)) {
            const int l = leaf(fq, f.n, ft, f.m, out + len);
            if (l < 0) return l;
            len += (size_t)l;
            continue;
        }
        if (top + 2 > 48) return kScratchOverflow;
        const int lhw = f.m / 2, rhw = f.m - lhw;
        int* L = sc.score;
        int* R = sc.score + f.n + 1;
        myers_cta_last_column<NW>(fq, f.n, ft, lhw, carry, carry_stride, progress, L);
        auto rq = [&](int i) { return Q(f.q0 + f.n - 1 - i); };
        auto rt = [&](int j) { return T(f.t0 + f.m - 1 - j); };
        myers_cta_last_column<NW>(rq, f.n, rt, rhw, carry, carry_stride, progress, R);
        // first interior row whose halves add up: every warp scans (same data, same answer)
        int row = -2;
        for (int r0 = 0; r0 <= f.n - 2 && row == -2; r0 += 32) {
            const int r = r0 + lane;
            const bool hit = r <= f.n - 2 && L[r + 1] + R[f.n - r - 1] == f.best;
            const unsigned ballot = __ballot_sync(kFull, hit);
            if (ballot) row = r0 + __ffs(ballot) - 1;
        }
        int ls = 0, rs = 0;
        if (row >= 0) { ls = L[row + 1]; rs = R[f.n - row - 1]; }
        if (row == -2 && lhw + R[f.n] == f.best) { row = -1; ls = lhw; rs = R[f.n]; }
        if (row == -2 && L[f.n] + rhw == f.best) { row = f.n - 1; ls = L[f.n]; rs = rhw; }
        __syncthreads();                                  // L / R are overwritten by the next frame
        if (row == -2) return kNoAlign;
        const int ul_h = row + 1;
        stack[top++] = Frame{f.q0 + ul_h, f.n - ul_h, f.t0 + lhw, rhw, rs};
        stack[top++] = Frame{f.q0, ul_h, f.t0, lhw, ls};
    }
    return (int)len;
}

}  // namespace lesvtw
