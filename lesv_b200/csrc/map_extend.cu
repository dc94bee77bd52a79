// map_extend.cu -- the extension stage of the mapper (qx2map) on the GPU: MEM-stage seeding inside each
// candidate window, chaining, and the base-level alignment of the best chains.
//
// Replaces (paths relative to the reference's src/):
//   hbn_extend_query_subseq_hit_list / s_hbn_extend_subject_subseq_hit_list   app/map/hbn_extend_subseq_hit.c:366, 203
//   InitHitFindData_Init / _AddQuery / _FindHits / s_collect_seeds            algo/init_hit_finder.c:360, 334, 387, 134
//   InitHitFindData_SetupMapAlignInfoFromHit, hbn_traceback (+ edlib)         trace_core.cuh
//   remove_contained_hsp, the HSP bookkeeping                                 hbn_extend_subseq_hit.c:102-128, 279-345 (host side, below)
//
// The reference sorts all 2*len 10-mers of a read once (33 % of qx2map's time) and merges every
// window's words against them.  Here the roles are swapped: a CTA sorts the ~len/10 words of one
// window in shared memory (bitonic, <= 16 K words = 192 KB with the counters), then streams the read's
// words of the one strand that window is on through a binary search, twice: once to count how often
// each window word occurs in the read (the caps of s_collect_seeds, :161-199, need both sides'
// multiplicities), once to emit the (window offset, read offset) seeds.  Seeds of all windows are
// sorted by one keys-only radix sort on (window, read offset, window offset) -- the order
// chain_seed_off_lt (:22-27) asks for; the seed SET does not depend on the reference's unstable word
// sort.  Chaining and alignment then run one thread per window / per candidate (chain_core.cuh,
// trace_core.cuh); the reference's sequential "stop after max_hsps, skip contained chains" logic is
// replayed on the host over the speculatively computed candidates.
#include "map_state.cuh"
#include "gpu_prims.cuh"
#include "trace_core.cuh"
#include "trace_warp.cuh"
#include "chain_warp.cuh"

#include <algorithm>
#include <cstring>

using namespace lesvmap;

namespace {

constexpr int kIhfThreads = 512;
constexpr int kMaxCandPerWindow = 3;                  // max_k of s_hbn_extend_subject_subseq_hit_list (:243), task rm
constexpr int kTraceThreadsPerBlock = 64;

struct WinDev {
    uint32_t read;           // read index inside the batch
    uint32_t strand;
    uint64_t s_off;          // residue index of the window's first base inside the subject .pac
    uint32_t wlen;
    uint32_t n_words;        // window words: offsets 0, stride, 2*stride ... <= wlen - word
};

struct CandOut {
    int status;              // lesvtb::Status
    int qoff, qend, soff, send, aln_len, n_match;
    int flags;               // bit 0/1: the reference would have tried the left/right overhang step
    int chain_score, n_mems;
    int mem_qbeg, mem_qend, mem_sbeg, mem_send;       // first / last MEM (chain_seed_list_is_contained)
    unsigned long long ops_off;
};

__device__ __forceinline__ uint32_t load_word(const uint8_t* __restrict__ pac, uint64_t pos, int k)   // k <= 13
{
    const uint8_t* b = pac + (pos >> 2);
    const uint32_t v = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    return (v << ((pos & 3) * 2)) >> (32 - 2 * k);
}
__device__ __forceinline__ uint32_t revcomp_word(uint32_t x, int k)
{
    x = ~x;
    x = __brev(x);
    x = ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);
    return x >> (32 - 2 * k);
}

// word of the read at strand position p
__device__ __forceinline__ uint32_t read_word(const uint8_t* __restrict__ qpac, const ReadDev& rd, uint32_t strand, uint32_t p, int k)
{
    if (!strand) return load_word(qpac, rd.off + p, k);
    return revcomp_word(load_word(qpac, rd.off + (rd.len - p - k), k), k);
}

__device__ __forceinline__ uint32_t lower_bound_hash(const uint64_t* keys, uint32_t n, uint32_t hash)
{
    const uint64_t x = (uint64_t)hash << 32;
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (keys[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}

// One CTA per window (persistent).  keys/cnt live in shared memory when the window has <= smem_words
// words, else in this CTA's slab of global scratch.
__global__ void __launch_bounds__(kIhfThreads)
ihf_kernel(const WinDev* __restrict__ wins, uint32_t n_win, const ReadDev* __restrict__ reads,
           const uint8_t* __restrict__ qpac, const uint8_t* __restrict__ spac, int word, int stride,
           uint32_t smem_words, uint64_t* __restrict__ gscratch, size_t gstride_words,
           uint64_t* __restrict__ seed_keys, uint32_t seed_cap, uint32_t* __restrict__ seed_counter,
           uint32_t* __restrict__ win_count, uint32_t* __restrict__ work_counter)
{
    extern __shared__ uint64_t smem64[];
    __shared__ uint32_t s_w, s_base, s_emit, s_red[kIhfThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        if (tid == 0) s_w = atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t w = s_w;
        if (w >= n_win) break;
        const WinDev W = wins[w];
        const ReadDev rd = reads[W.read];
        const uint32_t n_w = W.n_words;
        uint32_t P = 1;
        while (P < n_w) P <<= 1;
        uint64_t* keys = P <= smem_words ? smem64 : gscratch + (size_t)blockIdx.x * gstride_words;
        uint32_t* cnt = reinterpret_cast<uint32_t*>(keys + P);
        for (uint32_t i = tid; i < P; i += kIhfThreads) {
            keys[i] = i < n_w ? ((uint64_t)load_word(spac, W.s_off + (uint64_t)i * stride, word) << 32) | (i * (uint32_t)stride) : ~0ull;
            cnt[i] = 0;
        }
        __syncthreads();
        for (uint32_t k = 2; k <= P; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t t = tid; t < (P >> 1); t += kIhfThreads) {
                    const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                    const uint64_t a = keys[i], b = keys[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[l] = a; }
                }
                __syncthreads();
            }
        // pass 1: how often does each window word occur in the read (this strand)
        const uint32_t n_r = rd.len >= (uint32_t)word ? rd.len - word + 1 : 0;
        for (uint32_t p = tid; p < n_r; p += kIhfThreads) {
            const uint32_t h = read_word(qpac, rd, W.strand, p, word);
            const uint32_t idx = lower_bound_hash(keys, n_w, h);
            if (idx < n_w && (uint32_t)(keys[idx] >> 32) == h) atomicAdd(&cnt[idx], 1u);
        }
        __syncthreads();
        // caps of s_collect_seeds (init_hit_finder.c:161-199): window side nq <= 8, read side nr <= 8, nq*nr <= 8
        uint32_t mine = 0;
        for (uint32_t i = tid; i < n_w; i += kIhfThreads) {
            const uint32_t nr = cnt[i];
            if (!nr) continue;                        // only group heads are ever counted
            uint32_t nq = 1;
            const uint32_t h = (uint32_t)(keys[i] >> 32);
            while (nq <= 8 && i + nq < n_w && (uint32_t)(keys[i + nq] >> 32) == h) ++nq;
            if (nq * nr <= 8) { cnt[i] = nr | (nq << 16); mine += nq * nr; }
            else cnt[i] = 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, d);
        if (lane == 0) s_red[warp] = mine;
        __syncthreads();
        if (tid == 0) {
            uint32_t tot = 0;
            for (int x = 0; x < kIhfThreads / 32; ++x) tot += s_red[x];
            s_base = atomicAdd(seed_counter, tot);
            s_emit = 0;
            win_count[w] = tot;
        }
        __syncthreads();
        const uint32_t base = s_base;
        for (uint32_t p = tid; p < n_r; p += kIhfThreads) {
            const uint32_t h = read_word(qpac, rd, W.strand, p, word);
            const uint32_t idx = lower_bound_hash(keys, n_w, h);
            if (idx >= n_w || (uint32_t)(keys[idx] >> 32) != h) continue;
            const uint32_t c = cnt[idx];
            if (!c) continue;
            const uint32_t nq = c >> 16;
            const uint32_t pos = atomicAdd(&s_emit, nq);
            for (uint32_t x = 0; x < nq; ++x) {
                const uint32_t o = base + pos + x;
                if (o < seed_cap) seed_keys[o] = ((uint64_t)w << 40) | ((uint64_t)p << 20) | (uint32_t)keys[idx + x];
            }
        }
        __syncthreads();
    }
}

struct ChainBufs {
    lesvchain::Seed* seeds;
    int *f, *p, *t, *v;
    lesvchain::IntPair* u;
};

// One warp per window: chain the sorted seeds (InitHitFindData_FindHits, init_hit_finder.c:387-467), keep the
// seed lists of the <= 3 best chains.  The scoring DP runs on the whole warp (chain_warp.cuh), the
// backtracking on lane 0.
__global__ void __launch_bounds__(128)
ihf_chain_kernel(const uint64_t* __restrict__ seed_keys, const uint32_t* __restrict__ win_off, const uint32_t* __restrict__ win_count,
                 uint32_t n_win, int word, int min_score, ChainBufs cb, int* __restrict__ hs_w, int* __restrict__ hs_r,
                 int* __restrict__ cand_info /* per candidate: score, seed_off, seed_cnt */, int* __restrict__ cand_cnt)
{
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_win) return;
    const uint32_t a = win_off[w], n = win_count[w];
    if (n == 0) { if (lane == 0) cand_cnt[w] = 0; return; }
    lesvchain::Seed* seeds = cb.seeds + a;
    for (uint32_t i = lane; i < n; i += 32) {
        const uint64_t key = seed_keys[a + i];
        seeds[i].q = (int)(key & 0xFFFFF);            // chain "query" = window (roles as in init_hit_finder.c)
        seeds[i].s = (int)((key >> 20) & 0xFFFFF);    // chain "subject" = read
        seeds[i].len = word;
    }
    __syncwarp();
    const lesvchain::Params P = lesvchain::make_params(1, min_score);
    lesvchain::score_seeds_warp(P, seeds, (int)n, false, cb.f + a, cb.p + a, cb.t + a, cb.v + a);
    if (lane != 0) return;
    lesvchain::Hit pre[lesvchain::kMaxChains], out[lesvchain::kMaxJoined];
    int members[lesvchain::kMaxChains], n_pre = 0;
    const int nh = lesvchain::find_candidates_scored(P, seeds, (int)n, cb.f + a, cb.p + a, cb.t + a, cb.v + a, cb.u + a,
                                                     pre, &n_pre, out, members);
    cand_cnt[w] = 0;
    if (nh == 0) return;
    lesvsort::introsort(out, (size_t)nh, lesvchain::HitScoreGt());         // :465, on the already sorted list
    const int* v = cb.v + a;
    int pos = 0;
    const int nc = nh < kMaxCandPerWindow ? nh : kMaxCandPerWindow;
    for (int h = 0; h < nc; ++h) {
        const int start = pos;
        for (int m = 0; m < out[h].mem_cnt; ++m) {
            const lesvchain::Hit& c = pre[members[out[h].mem_off + m]];
            for (int x = c.v_cnt - 1; x >= 0; --x) {
                const lesvchain::Seed& sd = seeds[v[c.v_off + x]];
                hs_w[a + pos] = sd.q; hs_r[a + pos] = sd.s; ++pos;
            }
        }
        int* ci = cand_info + 3 * (kMaxCandPerWindow * (size_t)w + h);
        ci[0] = out[h].score; ci[1] = (int)a + start; ci[2] = pos - start;
    }
    cand_cnt[w] = nc;
}

// ---- base-level alignment, three kernels ---------------------------------------------------------------
// hbn_traceback spends its time in ~1 edlib_nw call per 40 bp (the gaps between trace seeds), and those calls
// do not depend on each other.  So:
//   trace_prep_kernel    one thread per candidate: MEMs, trace seeds, one GapJob per gap, queued by size class
//   gap_small_kernel<C>  one thread per gap, Myers columns in (hardware-interleaved) local memory
//   gap_big_kernel       the few gaps that need more than kSmallCells2 block-columns or edlib's Hirschberg
//                        route: persistent threads with 1 MB slabs in HBM
//   trace_finish_kernel  one thread per candidate: stitches seeds and gap alignments, trims, extends the ends
//                        (edlib_extend), overhang step, identity -- trace_core.cuh's traceback_ts unchanged.
struct GapJob {
    int qa, n, sa, m;            // edlib_nw(Q[qa, qa+n), S[sa, sa+m))
    uint32_t cand;               // index into cand_list
    int len;                     // ops written (or a negative lesvtb::Status)
    unsigned long long slot;     // first op in the slot pool (n + m bytes reserved)
};

struct CandPrep {
    int n_mems, tsc, status;
    uint32_t job_first;
    unsigned long long mem_off;  // mems at mem_pool + mem_off, trace seeds right behind them
};

constexpr int kGapClasses = 5;                         // 3 local-memory classes, big (one warp each), giant (ditto, started first)
constexpr int kSmallCells0 = 128, kSmallCells1 = 256, kSmallCells2 = 512;

struct PrepBufs {
    lesvtb::Mem* mem_pool; unsigned long long mem_cap; unsigned long long* mem_counter;
    GapJob* jobs; uint32_t jobs_cap; uint32_t* job_counter;
    unsigned long long slot_cap; unsigned long long* slot_counter;
    uint32_t* queue[kGapClasses]; uint32_t* queue_count;        // queue_count[kGapClasses]
};

__device__ __forceinline__ int gap_class(int n, int m)
{
    const int nb = (n + 63) >> 6;
    const long long cells = (long long)nb * m;
    if (nb == 1 && m <= kSmallCells0) return 0;
    if (nb <= 2 && cells <= kSmallCells1) return 1;
    if (nb <= 8 && cells <= kSmallCells2) return 2;
    if (cells <= 16384 && !lesvtb::needs_hirschberg(n, m)) return 3;
    return 4;
}

// One warp per candidate.  Co-diagonal seed runs are independent (s_extract_mem works on one run at a time), so
// the lanes take runs round-robin: count, prefix-sum, write.  Trace seeds are a short sequential filter (lane 0);
// the gap jobs follow from consecutive trace seeds in closed form (enumerate_gaps), 32 at a time.
__global__ void __launch_bounds__(128)
trace_prep_kernel(const WinDev* __restrict__ wins, const ReadDev* __restrict__ reads, const uint8_t* __restrict__ qpac,
                  const uint8_t* __restrict__ spac, const int* __restrict__ cand_list, uint32_t n_cand,
                  const int* __restrict__ cand_info, const int* __restrict__ hs_w, const int* __restrict__ hs_r, int word,
                  int* __restrict__ run_start /* scratch, per seed */, int* __restrict__ run_off /* scratch, per seed */,
                  PrepBufs pb, CandPrep* __restrict__ preps)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    const uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (x >= n_cand) return;
    const int c = cand_list[x];                       // = 3 * window + rank
    const WinDev W = wins[c / kMaxCandPerWindow];
    const ReadDev rd = reads[W.read];
    const int* ci = cand_info + 3 * (size_t)c;
    const int seed_off = ci[1], seed_cnt = ci[2];
    const int* sw = hs_w + seed_off;
    const int* sr = hs_r + seed_off;
    int* rs = run_start + seed_off;
    int* ro = run_off + seed_off;
    auto Q = [&](int i) -> uint32_t {
        return W.strand ? 3u - pac_code(qpac, rd.off + (rd.len - 1 - (uint32_t)i)) : pac_code(qpac, rd.off + (uint32_t)i);
    };
    auto S = [&](int j) -> uint32_t { return pac_code(spac, W.s_off + (uint32_t)j); };
    CandPrep P;
    P.n_mems = 0; P.tsc = 0; P.status = lesvtb::kOk; P.job_first = 0; P.mem_off = 0;
    // runs of seeds on one diagonal (init_hit_finder.c:480-492)
    int G = 0;
    for (int base = 0; base < seed_cnt; base += 32) {
        const int j = base + lane;
        const bool head = j < seed_cnt && (j == 0 || sw[j] - sw[j - 1] != sr[j] - sr[j - 1]);
        const unsigned bal = __ballot_sync(kFull, head);
        if (head) rs[G + __popc(bal & lt_mask)] = j;
        G += __popc(bal);
    }
    __syncwarp();
    auto run_of = [&](int g, int& ri, int& wi, int& ml) {
        const int i = rs[g], jend = g + 1 < G ? rs[g + 1] : seed_cnt;
        ri = sr[i]; wi = sw[i]; ml = sw[jend - 1] + word - sw[i];
    };
    for (int g = lane; g < G; g += 32) {
        int ri, wi, ml;
        run_of(g, ri, wi, ml);
        ro[g] = lesvtb::extract_mem_run(Q, S, ri, wi, ml, word, (lesvtb::Mem*)nullptr, 0, 0);
    }
    __syncwarp();
    int n_mems = 0;
    for (int base = 0; base < G; base += 32) {        // exclusive prefix sum of the per-run counts
        const int g = base + lane;
        const int cnt = g < G ? ro[g] : 0;
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(kFull, incl, d); if (lane >= d) incl += v; }
        if (g < G) ro[g] = n_mems + incl - cnt;
        n_mems += __shfl_sync(kFull, incl, 31);
    }
    __syncwarp();
    if (n_mems > 0) {
        const unsigned long long need = 2ull * (unsigned)n_mems + 2;
        unsigned long long off = 0;
        if (lane == 0) off = atomicAdd(pb.mem_counter, need);
        off = __shfl_sync(kFull, off, 0);
        if (off + need > pb.mem_cap) { P.status = lesvtb::kScratchOverflow; n_mems = 0; }
        else {
            lesvtb::Mem* mems = pb.mem_pool + off;
            lesvtb::Mem* ts = mems + n_mems;
            for (int g = lane; g < G; g += 32) {
                int ri, wi, ml;
                run_of(g, ri, wi, ml);
                lesvtb::extract_mem_run(Q, S, ri, wi, ml, word, mems, n_mems, ro[g]);
            }
            __syncwarp();
            P.mem_off = off;
            int tsc = 0;
            if (lane == 0) tsc = lesvtb::trace_points(mems, n_mems, ts);
            tsc = __shfl_sync(kFull, tsc, 0);
            __syncwarp();
            P.tsc = tsc;
            const int n_jobs = tsc - 1;
            const int E = 10, E2 = 20;
            auto gap_of = [&](int i, int& qa, int& n, int& sa, int& m) {          // enumerate_gaps, gap i
                const lesvtb::Mem a = ts[i], b = ts[i + 1];
                if (a.len > E2) { qa = a.q + a.len - E; sa = a.s + a.len - E; }
                else if (i == 0) { qa = a.q; sa = a.s; }
                else { qa = a.q + a.len / 2; sa = a.s + a.len / 2; }
                int qto, sto;
                if (b.len > E2) { qto = b.q + E; sto = b.s + E; }
                else { qto = b.q + b.len / 2; sto = b.s + b.len / 2; }
                n = qto - qa; m = sto - sa;
            };
            unsigned long long bytes = 0;
            for (int i = lane; i < n_jobs; i += 32) { int qa, n, sa, m; gap_of(i, qa, n, sa, m); bytes += (unsigned)(n + m); }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) bytes += __shfl_xor_sync(kFull, bytes, d);
            if (n_jobs > 0) {
                uint32_t j0 = 0;
                unsigned long long s0 = 0;
                if (lane == 0) { j0 = atomicAdd(pb.job_counter, (uint32_t)n_jobs); s0 = atomicAdd(pb.slot_counter, bytes); }
                j0 = __shfl_sync(kFull, j0, 0); s0 = __shfl_sync(kFull, s0, 0);
                if ((unsigned long long)j0 + (unsigned)n_jobs > pb.jobs_cap || s0 + bytes > pb.slot_cap) P.status = lesvtb::kScratchOverflow;
                else {
                    P.job_first = j0;
                    unsigned long long slot = s0;
                    for (int base = 0; base < n_jobs; base += 32) {
                        const int i = base + lane;
                        int qa = 0, n = 0, sa = 0, m = 0;
                        if (i < n_jobs) gap_of(i, qa, n, sa, m);
                        const unsigned long long sz = i < n_jobs ? (unsigned)(n + m) : 0u;
                        unsigned long long incl = sz;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) { const unsigned long long v = __shfl_up_sync(kFull, incl, d); if (lane >= d) incl += v; }
                        int cls = -1;
                        if (i < n_jobs) {
                            GapJob J;
                            J.qa = qa; J.n = n; J.sa = sa; J.m = m; J.cand = x; J.len = 0; J.slot = slot + incl - sz;
                            pb.jobs[j0 + i] = J;
                            if (n > 0 && m > 0) cls = gap_class(n, m);          // edlib_wrapper.c:43: nothing to align otherwise
                        }
#pragma unroll
                        for (int k = 0; k < kGapClasses; ++k) {
                            const unsigned bal = __ballot_sync(kFull, cls == k);
                            if (bal) {
                                uint32_t qb = 0;
                                if (lane == __ffs(bal) - 1) qb = atomicAdd(&pb.queue_count[k], (uint32_t)__popc(bal));
                                qb = __shfl_sync(kFull, qb, __ffs(bal) - 1);
                                if (cls == k) pb.queue[k][qb + __popc(bal & lt_mask)] = j0 + (uint32_t)i;
                            }
                        }
                        slot += __shfl_sync(kFull, incl, 31);
                    }
                }
            }
        }
    }
    P.n_mems = n_mems;
    if (lane == 0) preps[x] = P;
}

template <int CELLS, int CLS>
__global__ void __launch_bounds__(128)
gap_small_kernel(const WinDev* __restrict__ wins, const ReadDev* __restrict__ reads, const uint8_t* __restrict__ qpac,
                 const uint8_t* __restrict__ spac, const int* __restrict__ cand_list, GapJob* __restrict__ jobs,
                 const uint32_t* __restrict__ queue, const uint32_t* __restrict__ queue_count, uint8_t* __restrict__ slots)
{
    uint64_t Pc[CELLS], Mc[CELLS], peq[4 * 8];
    int score[CELLS];
    lesvtb::Scratch sc;
    sc.P = Pc; sc.M = Mc; sc.score = score; sc.peq = peq; sc.cap_cells = CELLS; sc.cap_peq = 32;
    sc.trim_table = nullptr; sc.trim_score = nullptr; sc.path_ave = 0;
    const uint32_t n_jobs = queue_count[CLS];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_jobs; j += gridDim.x * blockDim.x) {
        GapJob& J = jobs[queue[j]];
        const WinDev W = wins[cand_list[J.cand] / kMaxCandPerWindow];
        const ReadDev rd = reads[W.read];
        const int qa = J.qa, sa = J.sa;
        auto fq = [&](int i) -> uint32_t {
            return W.strand ? 3u - pac_code(qpac, rd.off + (rd.len - 1 - (uint32_t)(qa + i))) : pac_code(qpac, rd.off + (uint32_t)(qa + i));
        };
        auto fs = [&](int i) -> uint32_t { return pac_code(spac, W.s_off + (uint32_t)(sa + i)); };
        J.len = lesvtb::nw_path(fq, J.n, fs, J.m, sc, slots + J.slot, (size_t)(J.n + J.m));
    }
}

struct SlabBufs {
    uint64_t* P; uint64_t* M; int* score; uint64_t* peq;     // per thread slab
    size_t cells, peq_words;
};

// Persistent blocks of kBigWarps warps, a 1 MB slab per warp.  Giant gaps (edlib's Hirschberg route or > 16 K
// block-columns) first, one per block, the strips of the query pipelined over the block's warps
// (trace_warp.cuh: nw_path_cta); then the big ones, one per warp (nw_path_warp).
constexpr int kBigWarps = 8;

__global__ void __launch_bounds__(kBigWarps * 32)
gap_big_kernel(const WinDev* __restrict__ wins, const ReadDev* __restrict__ reads, const uint8_t* __restrict__ qpac,
               const uint8_t* __restrict__ spac, const int* __restrict__ cand_list, GapJob* __restrict__ jobs,
               const uint32_t* __restrict__ queue, const uint32_t* __restrict__ giant_queue,
               const uint32_t* __restrict__ queue_count, uint8_t* __restrict__ slots,
               SlabBufs sb, int8_t* __restrict__ carry_pool, size_t carry_stride, uint32_t* __restrict__ work_counter)
{
    __shared__ int progress[lesvtw::kMaxStrips];
    __shared__ int sh_len;
    __shared__ uint32_t sh_job;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    lesvtb::Scratch sc;
    sc.peq = nullptr; sc.cap_cells = sb.cells; sc.cap_peq = 0;
    sc.trim_table = nullptr; sc.trim_score = nullptr; sc.path_ave = 0;
    const uint32_t n_giant = queue_count[4], n_big = queue_count[3];
    // ---- giants: the whole block on one gap, the slab and carry buffers of the block's first warp onwards ----
    {
        const size_t w0 = (size_t)blockIdx.x * kBigWarps;
        sc.P = sb.P + w0 * sb.cells; sc.M = sb.M + w0 * sb.cells; sc.score = sb.score + w0 * sb.cells;
        sc.cap_cells = sb.cells * kBigWarps;             // the slabs of the block's warps are contiguous
        int8_t* carry = carry_pool + w0 * carry_stride;
        for (;;) {
            if (threadIdx.x == 0) sh_job = atomicAdd(work_counter, 1u);
            __syncthreads();
            const uint32_t j = sh_job;
            __syncthreads();
            if (j >= n_giant) break;
            GapJob& J = jobs[giant_queue[j]];
            const WinDev W = wins[cand_list[J.cand] / kMaxCandPerWindow];
            const ReadDev rd = reads[W.read];
            const int qa = J.qa, sa = J.sa, n = J.n, m = J.m;
            auto fq = [&](int i) -> uint32_t {
                return W.strand ? 3u - pac_code(qpac, rd.off + (rd.len - 1 - (uint32_t)(qa + i))) : pac_code(qpac, rd.off + (uint32_t)(qa + i));
            };
            auto fs = [&](int i) -> uint32_t { return pac_code(spac, W.s_off + (uint32_t)(sa + i)); };
            const int r = lesvtw::nw_path_cta<kBigWarps>(fq, n, fs, m, sc, carry, carry_stride, progress, &sh_len, slots + J.slot,
                                                         (size_t)(n + m));
            if (threadIdx.x == 0) J.len = r;
            __syncthreads();
        }
    }
    // ---- big gaps: one per warp ---------------------------------------------------------------------------
    sc.cap_cells = sb.cells;
    sc.P = sb.P + warp * sb.cells; sc.M = sb.M + warp * sb.cells; sc.score = sb.score + warp * sb.cells;
    int8_t* carry = carry_pool + warp * carry_stride;
    for (;;) {
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(work_counter + 1, 1u);
        j = __shfl_sync(0xFFFFFFFFu, j, 0);
        if (j >= n_big) break;
        GapJob& J = jobs[queue[j]];
        const WinDev W = wins[cand_list[J.cand] / kMaxCandPerWindow];
        const ReadDev rd = reads[W.read];
        const int qa = J.qa, sa = J.sa, n = J.n, m = J.m;
        auto fq = [&](int i) -> uint32_t {
            return W.strand ? 3u - pac_code(qpac, rd.off + (rd.len - 1 - (uint32_t)(qa + i))) : pac_code(qpac, rd.off + (uint32_t)(qa + i));
        };
        auto fs = [&](int i) -> uint32_t { return pac_code(spac, W.s_off + (uint32_t)(sa + i)); };
        const int r = lesvtw::nw_path_warp(fq, n, fs, m, sc, carry, slots + J.slot, (size_t)(n + m));
        if (lane == 0) J.len = r;
        __syncwarp();
    }
}

struct TraceBufs {
    SlabBufs sb;
    uint8_t* ops; size_t ops_cap;                             // per thread working ops
    uint8_t* blk;                                             // per thread kBlkOpsCap
    const lesvtb::Mem* mem_pool;
    const GapJob* jobs; const uint8_t* slots;
    uint8_t* pool; unsigned long long pool_cap; unsigned long long* pool_counter;
    const int16_t* trim_table; const int16_t* trim_score; int path_ave;
};

// Persistent warps, one alignment candidate at a time per warp.  The lanes stitch seeds and gap alignments
// (prefix sum over the gap lengths), lane 0 runs the sequential end game (trace_core.cuh: finish_ends), the lanes
// count the matches (calc_ident_perc) and copy the columns into the result pool.
__global__ void __launch_bounds__(128)
trace_finish_kernel(const WinDev* __restrict__ wins, const ReadDev* __restrict__ reads, const uint8_t* __restrict__ qpac,
                    const uint8_t* __restrict__ spac, const int* __restrict__ cand_list, uint32_t n_cand,
                    const int* __restrict__ cand_info, const CandPrep* __restrict__ preps,
                    int min_align_size, double min_ident, int over_hang, TraceBufs tb, CandOut* __restrict__ outs,
                    uint32_t* __restrict__ work_counter)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    lesvtb::Scratch sc;
    sc.P = tb.sb.P + warp * tb.sb.cells; sc.M = tb.sb.M + warp * tb.sb.cells; sc.score = tb.sb.score + warp * tb.sb.cells;
    sc.peq = tb.sb.peq + warp * tb.sb.peq_words; sc.cap_cells = tb.sb.cells; sc.cap_peq = tb.sb.peq_words;
    sc.trim_table = tb.trim_table; sc.trim_score = tb.trim_score; sc.path_ave = tb.path_ave;
    uint8_t* ops = tb.ops + warp * tb.ops_cap;
    uint8_t* blk = tb.blk + warp * lesvtb::kBlkOpsCap;
    const size_t cap = tb.ops_cap;
    for (;;) {
        uint32_t x = 0;
        if (lane == 0) x = atomicAdd(work_counter, 1u);
        x = __shfl_sync(kFull, x, 0);
        if (x >= n_cand) break;
        const int c = cand_list[x];
        const WinDev W = wins[c / kMaxCandPerWindow];
        const ReadDev rd = reads[W.read];
        const CandPrep P = preps[x];
        auto Q = [&](int i) -> uint32_t {
            return W.strand ? 3u - pac_code(qpac, rd.off + (rd.len - 1 - (uint32_t)i)) : pac_code(qpac, rd.off + (uint32_t)i);
        };
        auto S = [&](int j) -> uint32_t { return pac_code(spac, W.s_off + (uint32_t)j); };
        CandOut o;
        memset(&o, 0, sizeof o);
        o.chain_score = cand_info[3 * (size_t)c];
        o.n_mems = P.status == lesvtb::kOk ? P.n_mems : -1;
        if (P.status != lesvtb::kOk) o.status = P.status;
        else if (P.n_mems > 0) {
            const lesvtb::Mem* mems = tb.mem_pool + P.mem_off;
            const lesvtb::Mem* ts = mems + P.n_mems;
            const int n_mems = P.n_mems, tsc = P.tsc, n_gaps = P.tsc - 1;
            o.mem_qbeg = mems[0].q; o.mem_sbeg = mems[0].s;
            o.mem_qend = mems[n_mems - 1].q + mems[n_mems - 1].len; o.mem_send = mems[n_mems - 1].s + mems[n_mems - 1].len;
            const int q0 = mems[0].q, s0 = mems[0].s;
            size_t front = (size_t)q0 + (size_t)s0 + 8;
            const size_t alt = (size_t)(2.5 * q0) + 512;
            if (alt < front) front = alt;
            if (front > cap / 2) front = cap / 2;
            // ---- stitch (hbn_traceback.c:338-390): first seed, then gap i and seed i+1 for every gap ----------
            const int E = 10, E2 = 20;
            const lesvtb::Mem first = ts[0], last = ts[tsc - 1];
            int head = first.len > E2 ? first.len - E : 0;
            if (n_gaps == 0) head += first.len > E2 ? E : first.len;
            int err = 0;
            size_t pos = front;
            if (pos + (size_t)head > cap) err = lesvtb::kOpsOverflow;
            else {
                for (int a = lane; a < head; a += 32) ops[pos + a] = lesvtb::kOpDiag;
                pos += (size_t)head;
            }
            const GapJob* jobs = tb.jobs + P.job_first;
            for (int base = 0; base < n_gaps && err == 0; base += 32) {
                const int i = base + lane;
                int glen = 0, dafter = 0, my_err = 0;
                unsigned long long slot = 0;
                if (i < n_gaps) {
                    const GapJob J = jobs[i];
                    glen = J.len; slot = J.slot;
                    if (glen < 0) { my_err = glen; glen = 0; }
                    const lesvtb::Mem sj = ts[i + 1];
                    dafter = sj.len > E2 ? sj.len - 2 * E : 0;
                    if (i == n_gaps - 1) dafter += sj.len > E2 ? E : sj.len - sj.len / 2;
                }
                const int tot = glen + dafter;
                int incl = tot;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(kFull, incl, d); if (lane >= d) incl += v; }
                const int warp_tot = __shfl_sync(kFull, incl, 31);
                const unsigned errs = __ballot_sync(kFull, my_err != 0);
                if (errs) { err = __shfl_sync(kFull, my_err, __ffs(errs) - 1); break; }
                if (pos + (size_t)warp_tot > cap) { err = lesvtb::kOpsOverflow; break; }
                uint8_t* dst = ops + pos + (size_t)(incl - tot);
                const uint8_t* src = tb.slots + slot;
                for (int a = 0; a < glen; ++a) dst[a] = src[a];
                for (int a = 0; a < dafter; ++a) dst[glen + a] = lesvtb::kOpDiag;
                pos += (size_t)warp_tot;
            }
            __syncwarp();
            if (err) o.status = err;
            else {
                // ---- trim, extend, overhang: sequential, lane 0 ------------------------------------------------
                lesvtb::Result R;
                R.status = lesvtb::kNoAlign; R.qoff = R.qend = R.soff = R.send = 0; R.aln_len = 0; R.n_match = 0; R.ops_beg = 0;
                R.overhang_left = R.overhang_right = 0;
                if (lane == 0)
                    R = lesvtb::finish_ends(Q, (int)rd.len, S, (int)W.wlen, sc, ops, cap, front, pos, first.q, last.q + last.len,
                                            first.s, last.s + last.len, blk, over_hang != 0);
                __syncwarp();
                R.status = __shfl_sync(kFull, R.status, 0);
                R.qoff = __shfl_sync(kFull, R.qoff, 0); R.qend = __shfl_sync(kFull, R.qend, 0);
                R.soff = __shfl_sync(kFull, R.soff, 0); R.send = __shfl_sync(kFull, R.send, 0);
                R.aln_len = __shfl_sync(kFull, R.aln_len, 0);
                R.ops_beg = (size_t)__shfl_sync(kFull, (unsigned)R.ops_beg, 0);
                R.overhang_left = __shfl_sync(kFull, R.overhang_left, 0); R.overhang_right = __shfl_sync(kFull, R.overhang_right, 0);
                if (R.status == lesvtb::kOk) {
                    // ---- calc_ident_perc: 32 columns per step, positions from ballots -------------------------
                    int n_match = 0, qb = R.qoff, sb = R.soff;
                    const uint8_t* al = ops + R.ops_beg;
                    for (int a = 0; a < R.aln_len; a += 32) {
                        const bool valid = a + lane < R.aln_len;
                        const uint8_t op = valid ? al[a + lane] : (uint8_t)255;
                        const unsigned bq = __ballot_sync(kFull, valid && op != lesvtb::kOpDel);
                        const unsigned bs = __ballot_sync(kFull, valid && op != lesvtb::kOpIns);
                        bool same = false;
                        if (valid && op == lesvtb::kOpDiag) same = Q(qb + __popc(bq & lt_mask)) == S(sb + __popc(bs & lt_mask));
                        n_match += __popc(__ballot_sync(kFull, same));
                        qb += __popc(bq); sb += __popc(bs);
                    }
                    lesvtb::finalize_stats(R, n_match, min_align_size, min_ident);
                }
                o.status = R.status; o.qoff = R.qoff; o.qend = R.qend; o.soff = R.soff; o.send = R.send;
                o.aln_len = R.aln_len; o.n_match = R.n_match; o.flags = R.overhang_left | (R.overhang_right << 1);
                if (R.status == lesvtb::kOk) {
                    unsigned long long off = 0;
                    if (lane == 0) off = atomicAdd(tb.pool_counter, (unsigned long long)R.aln_len);
                    off = __shfl_sync(kFull, off, 0);
                    o.ops_off = off;
                    if (off + (unsigned long long)R.aln_len <= tb.pool_cap)
                        for (int a = lane; a < R.aln_len; a += 32) tb.pool[off + a] = ops[R.ops_beg + a];
                    else o.status = lesvtb::kOpsOverflow;
                }
            }
        }
        if (lane == 0) outs[c] = o;
        __syncwarp();
    }
}

__global__ void win_count_clear_kernel(uint32_t* p, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

}  // namespace

struct lesvmap_ext {
    DevBuf<WinDev> d_wins;
    DevBuf<uint64_t> d_keys0, d_keys1, d_gscratch, d_P, d_M, d_peq;
    DevBuf<uint32_t> d_win_count, d_win_off, d_counters, d_tmp;
    DevBuf<int> d_hs_w, d_hs_r, d_cand_info, d_cand_cnt, d_cand_list, d_score;
    DevBuf<lesvchain::Seed> d_seeds;
    DevBuf<int> d_f, d_p, d_t, d_v;
    DevBuf<lesvchain::IntPair> d_u;
    DevBuf<lesvtb::Mem> d_mems;
    DevBuf<uint8_t> d_ops, d_blk, d_pool;
    DevBuf<unsigned long long> d_pool_counter;
    DevBuf<CandOut> d_outs;
    DevBuf<GapJob> d_jobs;
    DevBuf<uint32_t> d_queue;
    DevBuf<uint8_t> d_slots;
    DevBuf<CandPrep> d_preps;
    DevBuf<int8_t> d_carry;
    cudaEvent_t ev[4] = {};
    bool ev_made = false;
    DevBuf<int16_t> d_trim;                  // DALIGNER TABLE | SCORE
    int path_ave = 0;
    std::vector<lesvgpu_hsp> hsps;
    std::vector<int> hsp_first;
    std::vector<uint8_t> ops_host;
    lesvgpu_extend_timing tm = {};
};

void lesvmap_ext_free(lesvmap_ext* e)
{
    if (!e) return;
    e->d_wins.release(); e->d_keys0.release(); e->d_keys1.release(); e->d_gscratch.release(); e->d_P.release(); e->d_M.release();
    e->d_peq.release(); e->d_win_count.release(); e->d_win_off.release(); e->d_counters.release(); e->d_tmp.release();
    e->d_hs_w.release(); e->d_hs_r.release(); e->d_cand_info.release(); e->d_cand_cnt.release(); e->d_cand_list.release();
    e->d_score.release(); e->d_seeds.release(); e->d_f.release(); e->d_p.release(); e->d_t.release(); e->d_v.release(); e->d_u.release();
    e->d_mems.release(); e->d_ops.release(); e->d_blk.release(); e->d_pool.release(); e->d_pool_counter.release();
    e->d_outs.release(); e->d_trim.release();
    e->d_jobs.release(); e->d_queue.release(); e->d_slots.release(); e->d_preps.release(); e->d_carry.release();
    if (e->ev_made) for (auto& x : e->ev) cudaEventDestroy(x);
    delete e;
}

static inline unsigned nblk(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }
static inline int bitlen(uint64_t x) { int b = 0; while (x) { ++b; x >>= 1; } return b; }

namespace {
struct WinScoreGt {       // subseq_hit_score_gt, app/map/hbn_subseq_hit.c:5
    bool operator()(const lesvgpu_subseq_hit& a, const lesvgpu_subseq_hit& b) const { return a.score > b.score; }
};
struct HspScoreGt {       // blasthsp_score_gt, ncbi_blast/setup/blast_hits.c:49
    bool operator()(const lesvgpu_hsp& a, const lesvgpu_hsp& b) const { return 8 * a.aln_len - a.n_match > 8 * b.aln_len - b.n_match; }
};
// subseq_is_contained, hbn_extend_subseq_hit.c:47-62 (note `i < hit_idx - 1` and the int truncation)
bool window_contained(const lesvgpu_subseq_hit* h, int idx)
{
    const int E = 200;
    const int sfrom = (int)h[idx].sfrom, sto = (int)h[idx].sto, qdir = h[idx].qdir;
    for (int i = 0; i < idx - 1; ++i) {
        if (h[i].qdir != qdir) continue;
        const int ss = (int)h[i].sfrom, st = (int)h[i].sto;
        if (sfrom + E >= ss && sto <= st + E) return true;
    }
    return false;
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI: seed + chain + extend for one query batch
// ---------------------------------------------------------------------------------------------
extern "C" int lesvgpu_map_batch(lesvgpu_ctx* ctx, const uint8_t* query_pac, const lesvgpu_seq* qseqs, int qfrom, int qto,
                                 const lesvgpu_map_opts* opts, const lesvgpu_hsp** hsps, const int** first,
                                 const uint8_t** ops_pool)
{
    if (hsps) *hsps = nullptr;
    if (first) *first = nullptr;
    if (ops_pool) *ops_pool = nullptr;
    const lesvgpu_subseq_hit* wins_host = nullptr;
    const int* wfirst = nullptr;
    int rc = lesvgpu_map_candidates(ctx, query_pac, qseqs, qfrom, qto, opts, &wins_host, &wfirst);
    if (rc != LESVGPU_OK) return rc;
    lesvmap_state* M = ctx->map;
    if (!M->ext) M->ext = new lesvmap_ext;
    lesvmap_ext* E = M->ext;
    cudaStream_t s = ctx->stream;
    const int nr = qto - qfrom;
    const int word = opts->memsc_kmer_size > 0 ? opts->memsc_kmer_size : 10;
    const int stride = opts->memsc_kmer_window > 0 ? opts->memsc_kmer_window : 10;
    const int memsc_score = opts->memsc_score > 0 ? opts->memsc_score : 30;
    const int max_hsps = opts->max_hsps > 0 ? opts->max_hsps : 5;
    const int min_align = opts->qcov_hsp_res > 0 ? opts->qcov_hsp_res : 100;
    if (word > 13) { ctx->err = "memsc_kmer_size > 13 is not supported"; return LESVGPU_E_ARG; }
    E->hsps.clear(); E->hsp_first.assign(nr + 1, 0); E->ops_host.clear();
    E->tm = lesvgpu_extend_timing{};

    // ---- which windows the reference could visit (s_hbn_extend_subject_subseq_hit_list :222-233) ------
    struct Group { int read, first, count; };           // windows of one read on one contig, sorted by score
    std::vector<lesvgpu_subseq_hit> sorted;              // all windows, group by group, in visiting order
    std::vector<Group> groups;
    std::vector<int> win_gpu;                            // per sorted window: index in the GPU window list or -1
    std::vector<WinDev> wins;
    uint32_t max_words = 0, max_qsize = 0;
    uint64_t pool_cap = 0;
    for (int r = 0; r < nr; ++r) {
        int i = wfirst[r];
        while (i < wfirst[r + 1]) {
            int j = i + 1;
            while (j < wfirst[r + 1] && wins_host[j].sid == wins_host[i].sid) ++j;
            Group g{r, (int)sorted.size(), j - i};
            sorted.insert(sorted.end(), wins_host + i, wins_host + j);
            lesvsort::introsort(sorted.data() + g.first, (size_t)g.count, WinScoreGt());
            for (int x = 0; x < g.count; ++x) {
                int gi = -1;
                if (x < max_hsps + 2 && !window_contained(sorted.data() + g.first, x)) {
                    const lesvgpu_subseq_hit& h = sorted[g.first + x];
                    WinDev W;
                    W.read = (uint32_t)r; W.strand = (uint32_t)h.qdir;
                    W.s_off = M->sseqs[h.sid].seq_offset + h.sfrom;
                    W.wlen = (uint32_t)(h.sto - h.sfrom);
                    W.n_words = W.wlen >= (uint32_t)word ? (W.wlen - word) / stride + 1 : 0;
                    if (W.wlen >= (1u << 20) || M->h_reads[r].len >= (1u << 20)) { ctx->err = "read or window longer than 2^20"; return LESVGPU_E_ARG; }
                    gi = (int)wins.size();
                    wins.push_back(W);
                    max_words = std::max(max_words, W.n_words);
                    max_qsize = std::max(max_qsize, M->h_reads[r].len);
                    pool_cap += 3ull * M->h_reads[r].len + 4096;
                }
                win_gpu.push_back(gi);
            }
            groups.push_back(g);
            i = j;
        }
    }
    const uint32_t n_win = (uint32_t)wins.size();
    if (first) *first = E->hsp_first.data();
    if (n_win == 0) { if (hsps) *hsps = E->hsps.data(); return LESVGPU_OK; }
    if (n_win >= (1u << 24)) { ctx->err = "more than 2^24 windows in one batch"; return LESVGPU_E_ARG; }

    cudaEvent_t* ev = ctx->ev;
    CK(E->d_wins.reserve(n_win));
    CK(E->d_win_count.reserve(n_win + 1)); CK(E->d_win_off.reserve(n_win + 1));
    CK(E->d_counters.reserve(8));
    CK(E->d_cand_cnt.reserve(n_win)); CK(E->d_cand_info.reserve(3 * (size_t)kMaxCandPerWindow * n_win));
    CK(E->d_outs.reserve((size_t)kMaxCandPerWindow * n_win));
    CK(cudaEventRecord(ev[0], s));
    CK(cudaMemcpyAsync(E->d_wins.p, wins.data(), sizeof(WinDev) * n_win, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(E->d_counters.p, 0, sizeof(uint32_t) * 8, s));
    CK(cudaMemsetAsync(E->d_win_count.p, 0, sizeof(uint32_t) * (n_win + 1), s));

    // ---- MEM-stage seeds ----------------------------------------------------------------------------
    uint32_t Pmax = 1;
    while (Pmax < max_words) Pmax <<= 1;
    const uint32_t smem_words = std::min<uint32_t>(Pmax, 16384);
    const size_t smem_bytes = (size_t)smem_words * 12;
    int dev_sms = ctx->n_sm;
    const int ctas = (int)std::min<uint32_t>(n_win, (uint32_t)dev_sms * (smem_bytes > 100 * 1024 ? 1 : smem_bytes > 50 * 1024 ? 2 : 4));
    size_t gstride = 0;
    if (Pmax > smem_words) { gstride = (size_t)Pmax * 3 / 2 + 8; CK(E->d_gscratch.reserve(gstride * ctas)); }
    uint64_t seed_cap = 0;
    for (const WinDev& W : wins) seed_cap += 2ull * W.n_words + 64;
    uint32_t n_seeds = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (seed_cap >= 0xFFFFFFF0ull) { ctx->err = "too many MEM-stage seeds in one batch"; return LESVGPU_E_NOMEM; }
        CK(E->d_keys0.reserve(seed_cap)); CK(E->d_keys1.reserve(seed_cap));
        CK(cudaMemsetAsync(E->d_counters.p, 0, sizeof(uint32_t) * 8, s));
        CK(lesv_need_smem(ihf_kernel, smem_bytes));
        ihf_kernel<<<ctas, kIhfThreads, smem_bytes, s>>>(E->d_wins.p, n_win, M->d_reads.p, M->d_qpac.p, ctx->d_subject.p, word, stride,
                                                         smem_words, E->d_gscratch.p, gstride, E->d_keys0.p, (uint32_t)seed_cap,
                                                         E->d_counters.p, E->d_win_count.p, E->d_counters.p + 1); ++g_lesv_launches;
        CK(cudaMemcpyAsync(&n_seeds, E->d_counters.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        if (n_seeds <= seed_cap) break;
        seed_cap = (uint64_t)n_seeds + 1024;           // rare: repeats; run again with room for everything
        if (attempt == 1) { ctx->err = "MEM-stage seed buffer overflow"; return LESVGPU_E_NOMEM; }
    }
    CK(cudaEventRecord(ev[1], s));
    CK(E->d_tmp.reserve(std::max(radix_tmp_words(n_seeds), scan_tmp_words((size_t)n_win + 1))));
    const int where = radix_sort_pairs(E->d_keys0.p, nullptr, E->d_keys1.p, nullptr, n_seeds, 0, 40 + std::max(1, bitlen(n_win - 1)),
                                       E->d_tmp.p, s);
    const uint64_t* seed_keys = where ? E->d_keys1.p : E->d_keys0.p;
    exclusive_scan_u32(E->d_win_count.p, E->d_win_off.p, n_win, nullptr, E->d_tmp.p, s);
    CK(cudaEventRecord(ev[2], s));

    // ---- chaining inside each window ------------------------------------------------------------------
    CK(E->d_seeds.reserve(n_seeds)); CK(E->d_f.reserve(n_seeds)); CK(E->d_p.reserve(n_seeds)); CK(E->d_t.reserve(n_seeds));
    CK(E->d_v.reserve(n_seeds)); CK(E->d_u.reserve(n_seeds)); CK(E->d_hs_w.reserve(n_seeds)); CK(E->d_hs_r.reserve(n_seeds));
    ChainBufs cb{E->d_seeds.p, E->d_f.p, E->d_p.p, E->d_t.p, E->d_v.p, E->d_u.p};
    ihf_chain_kernel<<<nblk((uint64_t)n_win * 32, 128), 128, 0, s>>>(seed_keys, E->d_win_off.p, E->d_win_count.p, n_win, word, memsc_score, cb,
                                                    E->d_hs_w.p, E->d_hs_r.p, E->d_cand_info.p, E->d_cand_cnt.p); ++g_lesv_launches;
    std::vector<int> cand_cnt(n_win);
    CK(cudaMemcpyAsync(cand_cnt.data(), E->d_cand_cnt.p, sizeof(int) * n_win, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ev[3], s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());

    // ---- base-level alignment of every candidate --------------------------------------------------------
    std::vector<int> cand_list;
    for (uint32_t w = 0; w < n_win; ++w)
        for (int k = 0; k < cand_cnt[w]; ++k) cand_list.push_back(kMaxCandPerWindow * (int)w + k);
    const uint32_t n_cand = (uint32_t)cand_list.size();
    std::vector<CandOut> outs((size_t)kMaxCandPerWindow * n_win);
    if (n_cand) {
        const size_t cells = 52500;                       // traceback mode of edlib holds < 1 MB = 52,428 block-columns
        const size_t peq_words = 4 * ((size_t)max_qsize / 64 + 2);
        const size_t ops_cap = (size_t)(3.5 * max_qsize) + 2048;
        const uint32_t n_big_warps = (uint32_t)dev_sms * 8;
        const uint32_t n_threads = std::min<uint32_t>((n_cand + 3) / 4 * 4, n_big_warps);    // warps of trace_finish_kernel
        const uint32_t n_slabs = n_big_warps;
        uint32_t max_wlen = 0;
        for (const WinDev& W : wins) max_wlen = std::max(max_wlen, W.wlen);
        CK(E->d_P.reserve(cells * n_slabs)); CK(E->d_M.reserve(cells * n_slabs)); CK(E->d_score.reserve(cells * n_slabs));
        CK(E->d_peq.reserve(peq_words * n_slabs));
        CK(E->d_ops.reserve(ops_cap * n_threads)); CK(E->d_blk.reserve((size_t)lesvtb::kBlkOpsCap * n_threads));
        const unsigned long long mem_cap = 4ull * n_seeds + 64ull * n_cand + 1024;
        const uint32_t jobs_cap = (uint32_t)std::min<unsigned long long>(2ull * n_seeds + 8ull * n_cand + 1024, 0xFFFFFFF0ull);
        unsigned long long slot_cap = 40ull * jobs_cap;
        for (int c : cand_list) slot_cap += (unsigned long long)M->h_reads[wins[c / kMaxCandPerWindow].read].len + wins[c / kMaxCandPerWindow].wlen;
        CK(E->d_mems.reserve(mem_cap));
        CK(E->d_jobs.reserve(jobs_cap)); CK(E->d_queue.reserve((size_t)kGapClasses * jobs_cap)); CK(E->d_slots.reserve(slot_cap));
        CK(E->d_preps.reserve(n_cand));
        CK(E->d_pool.reserve(pool_cap)); CK(E->d_pool_counter.reserve(8));
        CK(E->d_cand_list.reserve(n_cand));
        CK(cudaMemcpyAsync(E->d_cand_list.p, cand_list.data(), sizeof(int) * n_cand, cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(E->d_pool_counter.p, 0, 8 * sizeof(unsigned long long), s));     // counters, see below
        CK(cudaMemsetAsync(E->d_counters.p + 2, 0, 6 * sizeof(uint32_t), s));
        if (!E->d_trim.p) {
            std::vector<int16_t> tt(65536);
            lesvtb::dalign_tables(tt.data(), tt.data() + 32768, &E->path_ave);
            CK(E->d_trim.reserve(65536));
            CK(cudaMemcpyAsync(E->d_trim.p, tt.data(), sizeof(int16_t) * 65536, cudaMemcpyHostToDevice, s));
            CK(cudaStreamSynchronize(s));
        }
        if (!E->ev_made) {
            for (auto& e : E->ev) CK(cudaEventCreate(&e));
            E->ev_made = true;
        }
        // d_pool_counter: [0] alignment pool, [1] MEM pool, [2] slot pool; as uint32 behind them: [6] job counter, [8..11] queue counts
        uint32_t* c32 = reinterpret_cast<uint32_t*>(E->d_pool_counter.p + 3);
        PrepBufs pb;
        pb.mem_pool = E->d_mems.p; pb.mem_cap = mem_cap; pb.mem_counter = E->d_pool_counter.p + 1;
        pb.jobs = E->d_jobs.p; pb.jobs_cap = jobs_cap; pb.job_counter = c32;
        pb.slot_cap = slot_cap; pb.slot_counter = E->d_pool_counter.p + 2;
        for (int k = 0; k < kGapClasses; ++k) pb.queue[k] = E->d_queue.p + (size_t)k * jobs_cap;
        pb.queue_count = c32 + 2;
        trace_prep_kernel<<<nblk((uint64_t)n_cand * 32, 128), 128, 0, s>>>(E->d_wins.p, M->d_reads.p, M->d_qpac.p, ctx->d_subject.p,
                                                                           E->d_cand_list.p, n_cand, E->d_cand_info.p, E->d_hs_w.p, E->d_hs_r.p,
                                                                           word, E->d_f.p, E->d_p.p, pb, E->d_preps.p); ++g_lesv_launches;
        CK(cudaEventRecord(E->ev[0], s));
        CK(cudaEventRecord(E->ev[2], s));               // the side stream of the giant gaps starts here
        SlabBufs sb;
        sb.P = E->d_P.p; sb.M = E->d_M.p; sb.score = E->d_score.p; sb.peq = E->d_peq.p; sb.cells = cells; sb.peq_words = peq_words;
        const size_t carry_stride = ((size_t)max_wlen + 64 + 15) & ~(size_t)15;
        CK(E->d_carry.reserve(carry_stride * n_big_warps));
        // the few giant gaps (one warp each, latency-bound: 25 ms for a handful of them) run beside the three small-gap
        // launches on a side stream: disjoint job queues, disjoint result slots
        cudaStream_t bs = ctx->side[0];
        CK(cudaStreamWaitEvent(bs, E->ev[2], 0));
        gap_big_kernel<<<n_big_warps / kBigWarps, kBigWarps * 32, 0, bs>>>(E->d_wins.p, M->d_reads.p, M->d_qpac.p, ctx->d_subject.p, E->d_cand_list.p, E->d_jobs.p,
                                                       pb.queue[3], pb.queue[4], pb.queue_count, E->d_slots.p, sb, E->d_carry.p, carry_stride,
                                                       E->d_counters.p + 3); ++g_lesv_launches;
        CK(cudaEventRecord(E->ev[3], bs));
        const unsigned small_grid = (unsigned)dev_sms * 8;
        gap_small_kernel<kSmallCells0, 0><<<small_grid, 128, 0, s>>>(E->d_wins.p, M->d_reads.p, M->d_qpac.p, ctx->d_subject.p, E->d_cand_list.p,
                                                                     E->d_jobs.p, pb.queue[0], pb.queue_count, E->d_slots.p); ++g_lesv_launches;
        gap_small_kernel<kSmallCells1, 1><<<small_grid, 128, 0, s>>>(E->d_wins.p, M->d_reads.p, M->d_qpac.p, ctx->d_subject.p, E->d_cand_list.p,
                                                                     E->d_jobs.p, pb.queue[1], pb.queue_count, E->d_slots.p); ++g_lesv_launches;
        gap_small_kernel<kSmallCells2, 2><<<small_grid, 128, 0, s>>>(E->d_wins.p, M->d_reads.p, M->d_qpac.p, ctx->d_subject.p, E->d_cand_list.p,
                                                                     E->d_jobs.p, pb.queue[2], pb.queue_count, E->d_slots.p); ++g_lesv_launches;
        CK(cudaStreamWaitEvent(s, E->ev[3], 0));
        CK(cudaEventRecord(E->ev[1], s));
        TraceBufs tb;
        tb.sb = sb;
        tb.trim_table = E->d_trim.p; tb.trim_score = E->d_trim.p + 32768; tb.path_ave = E->path_ave;
        tb.ops = E->d_ops.p; tb.ops_cap = ops_cap; tb.blk = E->d_blk.p; tb.mem_pool = E->d_mems.p;
        tb.jobs = E->d_jobs.p; tb.slots = E->d_slots.p;
        tb.pool = E->d_pool.p; tb.pool_cap = pool_cap; tb.pool_counter = E->d_pool_counter.p;
        trace_finish_kernel<<<n_threads / 4, 128, 0, s>>>(
            E->d_wins.p, M->d_reads.p, M->d_qpac.p, ctx->d_subject.p, E->d_cand_list.p, n_cand, E->d_cand_info.p, E->d_preps.p,
            min_align, opts->perc_identity, opts->skip_overhang ? 0 : 1, tb, E->d_outs.p, E->d_counters.p + 2); ++g_lesv_launches;
        CK(cudaEventRecord(ev[4], s));
        CK(cudaMemcpyAsync(outs.data(), E->d_outs.p, sizeof(CandOut) * outs.size(), cudaMemcpyDeviceToHost, s));
        unsigned long long pool_used = 0;
        CK(cudaMemcpyAsync(&pool_used, E->d_pool_counter.p, sizeof pool_used, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        if (pool_used > pool_cap) { ctx->err = "alignment pool overflow"; return LESVGPU_E_NOMEM; }
        if (ops_pool && opts->want_ops) {
            E->ops_host.resize(pool_used);
            if (pool_used) CK(cudaMemcpyAsync(E->ops_host.data(), E->d_pool.p, pool_used, cudaMemcpyDeviceToHost, s));
        }
        CK(cudaEventRecord(ev[5], s));
        CK(cudaStreamSynchronize(s));
        cudaEventElapsedTime(&E->tm.trace_ms, ev[3], ev[4]);
        cudaEventElapsedTime(&E->tm.prep_ms, ev[3], E->ev[0]);
        cudaEventElapsedTime(&E->tm.gap_ms, E->ev[0], E->ev[1]);
        cudaEventElapsedTime(&E->tm.finish_ms, E->ev[1], ev[4]);
        {
            uint32_t qc[kGapClasses + 2] = {};
            cudaMemcpy(qc, reinterpret_cast<uint32_t*>(E->d_pool_counter.p + 3), sizeof qc, cudaMemcpyDeviceToHost);
            E->tm.n_gap_jobs = qc[0];
            if (getenv("LESV_GAP_DEBUG")) {
                std::vector<GapJob> hj(qc[0]);
                cudaMemcpy(hj.data(), E->d_jobs.p, sizeof(GapJob) * qc[0], cudaMemcpyDeviceToHost);
                std::sort(hj.begin(), hj.end(), [](const GapJob& a, const GapJob& b) { return (long long)a.n * a.m > (long long)b.n * b.m; });
                fprintf(stderr, "[gap debug] giant=%u big=%u; largest:", qc[6], qc[5]);
                for (size_t k = 0; k < 12 && k < hj.size(); ++k) fprintf(stderr, " %dx%d", hj[k].n, hj[k].m);
                fprintf(stderr, "\n");
            }
            for (int k = 0; k < 4; ++k) E->tm.n_gaps_class[k] = qc[2 + k];
            E->tm.n_gaps_class[3] += qc[2 + 4];
        }
        cudaEventElapsedTime(&E->tm.d2h_ms, ev[4], ev[5]);
        E->tm.d2h_bytes = sizeof(CandOut) * outs.size() + E->ops_host.size();
    }
    cudaEventElapsedTime(&E->tm.ihf_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&E->tm.sort_ms, ev[1], ev[2]);
    cudaEventElapsedTime(&E->tm.chain_ms, ev[2], ev[3]);
    E->tm.n_windows = n_win; E->tm.n_seeds = n_seeds; E->tm.n_candidates = n_cand;

    // ---- host: the reference's sequential selection over the computed candidates --------------------------
    size_t gi = 0;
    for (int r = 0; r < nr; ++r) {
        const int qsize = (int)M->h_reads[r].len;
        for (; gi < groups.size() && groups[gi].read == r; ++gi) {
            const Group& g = groups[gi];
            std::vector<lesvgpu_hsp> list;
            for (int x = 0; x < g.count && x < max_hsps + 2 && (int)list.size() < max_hsps; ++x) {
                const int gw = win_gpu[g.first + x];
                if (gw < 0) continue;                   // contained window
                const lesvgpu_subseq_hit& h = sorted[g.first + x];
                for (int k = 0; k < cand_cnt[gw]; ++k) {
                    const CandOut& o = outs[(size_t)kMaxCandPerWindow * gw + k];
                    if (o.n_mems == 0) continue;
                    if (o.status == lesvtb::kNeedsHirschberg) {
                        // trace_core.cuh marks this "cannot happen" (SHW blocks are <= 1459 x 1459).  Dropping the
                        // candidate would silently diverge from the reference, so it is an error like the overflows.
                        ++E->tm.n_unsupported;
                        ctx->err = "an end-extension block needs the Hirschberg route, which the GPU path does not have";
                        return LESVGPU_E_CUDA;
                    }
                    if (o.status < 0) {
                        ctx->err = "alignment scratch overflow (status " + std::to_string(o.status) + ")";
                        return LESVGPU_E_NOMEM;
                    }
                    // chain_seed_list_is_contained, hbn_extend_subseq_hit.c:64-100
                    bool contained = false;
                    {
                        const int E200 = 200;
                        const int qbeg = o.mem_qbeg, qend = o.mem_qend;
                        const int sbeg = o.mem_sbeg + (int)h.sfrom, send = o.mem_send + (int)h.sfrom;
                        for (const lesvgpu_hsp& p : list) {
                            if (p.qdir != h.qdir) continue;
                            if (qbeg + E200 >= p.qoff && qend <= p.qend + E200 && sbeg + E200 >= p.soff && send <= p.send + E200) { contained = true; break; }
                        }
                    }
                    if (contained) continue;
                    if (o.status != lesvtb::kOk) continue;
                    lesvgpu_hsp p;
                    p.qid = qfrom + r; p.qdir = h.qdir; p.qoff = o.qoff; p.qend = o.qend; p.qsize = qsize;
                    p.sid = h.sid; p.soff = o.soff + (int)h.sfrom; p.send = o.send + (int)h.sfrom;
                    p.ssize = (int64_t)M->sseqs[h.sid].seq_size;
                    p.aln_len = o.aln_len; p.n_match = o.n_match;
                    p.chain_score = o.chain_score; p.ddf_score = h.score;
                    p.ops_off = (size_t)o.ops_off;
                    p.flags = o.flags;
                    list.push_back(p);
                    if ((int)list.size() == max_hsps) break;
                }
            }
            if (list.empty()) continue;
            // remove_contained_hsp, hbn_extend_subseq_hit.c:102-128
            lesvsort::introsort(list.data(), list.size(), HspScoreGt());
            std::vector<char> dead(list.size(), 0);
            for (size_t i = 0; i < list.size(); ++i) {
                if (dead[i]) continue;
                for (size_t j = i + 1; j < list.size(); ++j) {
                    if (dead[j] || list[j].qdir != list[i].qdir) continue;
                    const int E200 = 200;
                    if (list[j].qoff + E200 >= list[i].qoff && list[j].qend <= list[i].qend + E200 &&
                        list[j].soff + E200 >= list[i].soff && list[j].send <= list[i].send + E200) dead[j] = 1;
                }
            }
            for (size_t i = 0; i < list.size(); ++i) if (!dead[i]) E->hsps.push_back(list[i]);
        }
        E->hsp_first[r + 1] = (int)E->hsps.size();
    }
    if (hsps) *hsps = E->hsps.data();
    if (ops_pool && opts->want_ops) *ops_pool = E->ops_host.data();
    return LESVGPU_OK;
}

extern "C" void lesvgpu_map_last_extend_timing(const lesvgpu_ctx* ctx, lesvgpu_extend_timing* t)
{
    if (ctx && ctx->map && ctx->map->ext && t) *t = ctx->map->ext->tm;
}

// ---------------------------------------------------------------------------------------------
// test hook: edlib_nw of one pair by one warp (trace_warp.cuh), residues coded 0..3
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(32)
debug_nw_kernel(const uint8_t* __restrict__ q, int n, const uint8_t* __restrict__ t, int m, SlabBufs sb, int8_t* carry,
                uint8_t* __restrict__ out, int* __restrict__ n_out)
{
    lesvtb::Scratch sc;
    sc.P = sb.P; sc.M = sb.M; sc.score = sb.score; sc.peq = nullptr; sc.cap_cells = sb.cells; sc.cap_peq = 0;
    sc.trim_table = nullptr; sc.trim_score = nullptr; sc.path_ave = 0;
    auto fq = [&](int i) -> uint32_t { return q[i]; };
    auto ft = [&](int j) -> uint32_t { return t[j]; };
    const int r = lesvtw::nw_path_warp(fq, n, ft, m, sc, carry, out, (size_t)(n + m));
    if (threadIdx.x == 0) *n_out = r;
}

__global__ void __launch_bounds__(kBigWarps * 32)
debug_nw_block_kernel(const uint8_t* __restrict__ q, int n, const uint8_t* __restrict__ t, int m, SlabBufs sb, int8_t* carry,
                      size_t carry_stride, uint8_t* __restrict__ out, int* __restrict__ n_out)
{
    __shared__ int progress[lesvtw::kMaxStrips];
    __shared__ int sh_len;
    lesvtb::Scratch sc;
    sc.P = sb.P; sc.M = sb.M; sc.score = sb.score; sc.peq = nullptr; sc.cap_cells = sb.cells; sc.cap_peq = 0;
    sc.trim_table = nullptr; sc.trim_score = nullptr; sc.path_ave = 0;
    auto fq = [&](int i) -> uint32_t { return q[i]; };
    auto ft = [&](int j) -> uint32_t { return t[j]; };
    const int r = lesvtw::nw_path_cta<kBigWarps>(fq, n, ft, m, sc, carry, carry_stride, progress, &sh_len, out, (size_t)(n + m));
    if (threadIdx.x == 0) *n_out = r;
}
}  // namespace

static int debug_nw(lesvgpu_ctx* ctx, const uint8_t* q, int n, const uint8_t* t, int m, uint8_t* ops_out, int* n_ops, bool block)
{
    if (!ctx || !q || !t || !ops_out || !n_ops || n < 0 || m < 0) return LESVGPU_E_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return LESVGPU_E_CUDA;
    const size_t cells = std::max<size_t>(52500, 2 * ((size_t)n + 1));
    const size_t carry_stride = ((size_t)m + 64 + 15) & ~(size_t)15;
    DevBuf<uint64_t> P, Mv; DevBuf<int> score; DevBuf<uint8_t> dq, dt, dout; DevBuf<int8_t> carry; DevBuf<int> dn;
    CK(P.reserve(cells)); CK(Mv.reserve(cells)); CK(score.reserve(cells)); CK(dq.reserve(n + 1)); CK(dt.reserve(m + 1));
    CK(dout.reserve((size_t)n + m + 1)); CK(carry.reserve(carry_stride * kBigWarps)); CK(dn.reserve(1));
    CK(cudaMemcpy(dq.p, q, n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dt.p, t, m, cudaMemcpyHostToDevice));
    SlabBufs sb{P.p, Mv.p, score.p, nullptr, cells, 0};
    if (block) debug_nw_block_kernel<<<1, kBigWarps * 32, 0, ctx->stream>>>(dq.p, n, dt.p, m, sb, carry.p, carry_stride, dout.p, dn.p);
    else debug_nw_kernel<<<1, 32, 0, ctx->stream>>>(dq.p, n, dt.p, m, sb, carry.p, dout.p, dn.p);
    ++g_lesv_launches;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    CK(cudaMemcpy(n_ops, dn.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (*n_ops > 0) CK(cudaMemcpy(ops_out, dout.p, *n_ops, cudaMemcpyDeviceToHost));
    P.release(); Mv.release(); score.release(); dq.release(); dt.release(); dout.release(); carry.release(); dn.release();
    return LESVGPU_OK;
}

extern "C" int lesvgpu_debug_edlib_nw(lesvgpu_ctx* ctx, const uint8_t* q, int n, const uint8_t* t, int m, uint8_t* ops_out, int* n_ops)
{
    return debug_nw(ctx, q, n, t, m, ops_out, n_ops, false);
}
extern "C" int lesvgpu_debug_edlib_nw_block(lesvgpu_ctx* ctx, const uint8_t* q, int n, const uint8_t* t, int m, uint8_t* ops_out, int* n_ops)
{
    return debug_nw(ctx, q, n, t, m, ops_out, n_ops, true);
}
