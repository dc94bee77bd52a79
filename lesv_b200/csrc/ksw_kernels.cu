// ksw_kernels.cu -- K-ksw: banded dual-affine global DP with traceback on sm_100a.
//
// Replaces ksw_extd2_sse + ksw_backtrack (reference src/algo/ksw2_extd2_sse.c:26-393,
// src/algo/ksw2.h:121-153) as called by nw_ksw2_extd2 (src/algo/ksw2_wrapper.c:97-211).
//
// Mapping.  One CTA per alignment, NT threads, one 16-column block of the band per thread held
// entirely in registers (ksw_core.cuh).  Thread tid owns the block B == tid (mod NT) of the live
// window; as the band slides right along the anti-diagonals a thread drops the block that left
// on the left and picks up the one NT blocks further right.  Per anti-diagonal each thread does
// one block update (~310 integer instructions for 16 cells), leaves its column-15 values in shared
// memory for its right neighbour and stores 16 traceback bytes with one coalesced 128-bit store.
// One named barrier per anti-diagonal among the warps that hold blocks.  The packed target and the
// packed reversed query (2 bits per base) are staged in shared memory once per alignment.
// CTAs are persistent and pull alignments from a queue sorted by decreasing cost; each CTA owns
// one traceback arena in HBM that is reused from alignment to alignment.  The traceback of an
// alignment is walked by one warp (64 rows prefetched per round trip to HBM) -- in the kernel of the
// wide single-CTA tiers by an extra warp, while the other warps already fill the next alignment.
#include "ksw_kernels.cuh"
#include "ksw_core.cuh"
#include "gpu_prims.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace lesvksw {

// ---------------------------------------------------------------------------------------------
// packing: byte codes -> 2-bit words
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pac_base(const uint8_t* __restrict__ pac, uint64_t l)
{
    return (pac[l >> 2] >> ((~l & 3) << 1)) & 3u;             // _get_pac, hbn_aux.h:126
}

// One CTA per alignment.  Writes the packed target (column c of block B at bits 2c of word B) and
// the packed REVERSED query behind one zero pad word.  For .pac sources this is also where the read
// range is reverse-complemented (RawReadReader_ExtractSubRead, raw_reads_reader.c:83-124): the
// reversed query of a REV read is simply the complement of the forward range.
__global__ void pack_kernel(const KswJob* __restrict__ jobs, const uint8_t* __restrict__ raw,
                            const uint8_t* __restrict__ subject_pac, uint32_t* __restrict__ pool)
{
    const KswJob job = jobs[blockIdx.x];
    const int ntw = (int)tp_words(job.tlen), nqw = (int)qrp_words(job.qlen);
    const bool pac = job.src == 1;
    for (int wi = threadIdx.x; wi < ntw; wi += blockDim.x) {
        uint32_t w = 0;
        const int base = wi * 16;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int idx = base + k;
            if (idx < job.tlen) {
                const uint32_t c = pac ? pac_base(subject_pac, job.seq_off_t + (uint64_t)idx)
                                       : (uint32_t)(raw[job.seq_off_t + idx] & 3);
                w |= c << (2 * k);
            }
        }
        pool[job.tp_off + wi] = w;
    }
    for (int wi = threadIdx.x; wi < nqw; wi += blockDim.x) {
        uint32_t w = 0;
        const int base = (wi - 1) * 16;                 // word 0 is the zero pad
        if (wi > 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int idx = base + k;               // index into the reversed query
                if (idx < job.qlen) {
                    uint32_t c;
                    if (!pac) c = (uint32_t)(raw[job.seq_off_q + (job.qlen - 1 - idx)] & 3);
                    else if (job.qdir) c = 3u - pac_base(raw, job.seq_off_q + (uint64_t)idx);
                    else c = pac_base(raw, job.seq_off_q + (uint64_t)(job.qlen - 1 - idx));
                    w |= c << (2 * k);
                }
            }
        }
        pool[job.qrp_off + wi] = w;
    }
}

void launch_pack(const KswJob* d_jobs, int njobs, const uint8_t* d_raw, const uint8_t* d_subject_pac,
                 uint32_t* d_pool, cudaStream_t s)
{
    if (njobs > 0) pack_kernel<<<njobs, 256, 0, s>>>(d_jobs, d_raw, d_subject_pac, d_pool);
}

// ---------------------------------------------------------------------------------------------
// backtrack (warp 0 of the CTA)
// ---------------------------------------------------------------------------------------------
constexpr int kBtRows = 64;

struct BtShared {
    uint4 cache[kBtRows][2];
    int blk[kBtRows][2];
};

__device__ __forceinline__ void bt_emit(uint32_t op, uint32_t len, uint32_t& cur_op, uint32_t& cur_len,
                                        uint32_t* cigar, int& n, bool write)
{
    if (op == cur_op) { cur_len += len; return; }
    if (cur_len && write) cigar[n] = cur_len << 4 | cur_op;
    if (cur_len) ++n;
    cur_op = op; cur_len = len;
}

// Alignment scan folded into the traceback walk.  The walk visits the alignment columns from the
// LAST to the first, so the three forward scans of the reference are restated right-to-left:
//   truncate_align_bad_ends (hbn_traceback.c:547-606): the kept range ends with the right-most
//     window of 8 identical columns and starts with the left-most one; walking leftwards, the first
//     time the run of identical columns reaches 8 fixes the right end, and every later column at
//     which the run is >= 8 is a new candidate for the left end (the last candidate wins);
//   calc_effective_ident_perc (hbn_traceback_aux.c:22-95): counters restart at the right end and are
//     snapshotted at every left-end candidate; gap runs >= 20 columns do not count;
//   find_sv_signature (find_sv_signature.c:126-220): gap runs >= min_indel are recorded when the run
//     closes; records right of the right end are dropped, the snapshot bounds them on the left.
struct ScanState {
    int m;                 // identical columns in a row, ending at the current column
    int gap_op, gap_len;   // open gap run (0 = none; 1 = I, 2 = D) and its length
    int gap_i, gap_j;      // cell of its right-most column
    int have_right, right_fail;
    int qe, se;
    int eff_mat, eff_len, mat, len, n_runs;
    // snapshot at the current left-end candidate
    int s_valid, s_at_end, qb, sb, s_eff_mat, s_eff_len, s_mat, s_len, s_runs;
};

__device__ __forceinline__ void scan_close_gap(ScanState& z, int min_indel, KswGapRun* runs, bool write)
{
    if (!z.gap_op) return;
    const int n = z.gap_len;
    if (z.have_right) {
        z.len += n;
        if (n < 20) z.eff_len += n;
        if (min_indel > 0 && n >= min_indel) {
            if (write && z.n_runs < kSigCap) {
                KswGapRun g;
                if (z.gap_op == 2) {            // deletion: target columns gap_i-n+1 .. gap_i, read stays at gap_j+1
                    g.type = 0; g.qfrom = z.gap_j + 1; g.qto = z.gap_j + 2;
                    g.sfrom = z.gap_i - n + 1; g.sto = z.gap_i + 1;
                } else {                        // insertion: read bases gap_j-n+1 .. gap_j, target stays at gap_i+1
                    g.type = 6; g.qfrom = z.gap_j - n + 1; g.qto = z.gap_j + 1;
                    g.sfrom = z.gap_i + 1; g.sto = z.gap_i + 2;
                }
                runs[z.n_runs] = g;
            }
            ++z.n_runs;
        }
    }
    z.gap_op = 0; z.gap_len = 0;
}

__device__ __forceinline__ void scan_column(ScanState& z, uint32_t op, bool ident, int i, int j, int qlen, int tlen,
                                            int min_indel, KswGapRun* runs, bool write)
{
    if (op == 0) {
        scan_close_gap(z, min_indel, runs, write);
        z.m = ident ? z.m + 1 : 0;
        z.eff_len += 1; z.eff_mat += ident; z.len += 1; z.mat += ident;
        if (z.m >= 8) {
            if (!z.have_right) {
                z.have_right = 1; z.qe = j + 8; z.se = i + 8;
                z.right_fail = (i == 0 && j == 0);
                z.eff_len = z.eff_mat = z.len = z.mat = 8; z.n_runs = 0;
            }
            z.s_valid = 1; z.qb = j; z.sb = i;
            z.s_at_end = (i + 8 == tlen && j + 8 == qlen);
            z.s_eff_mat = z.eff_mat; z.s_eff_len = z.eff_len; z.s_mat = z.mat; z.s_len = z.len; z.s_runs = z.n_runs;
        }
    } else {
        z.m = 0;
        if ((int)op != z.gap_op) {
            scan_close_gap(z, min_indel, runs, write);
            z.gap_op = (int)op; z.gap_len = 0; z.gap_i = i; z.gap_j = j;
        }
        ++z.gap_len;
    }
}

// All 32 lanes run the same walk out of the shared-memory row cache; lane 0 writes results.
// kOverlap (ksw_dp_overlap_kernel): the traceback may be stored top-down from the end of the arena (reversed), the next
// alignment's DP is writing into the same arena while this walk runs, so rows are read with L2-only loads and the
// row the walk still needs is published in *progress before every refill (-1 once nothing is needed any more).
template <bool kOverlap = false>
__device__ void backtrack_warp(const KswJob& job, const uint8_t* __restrict__ arena, size_t prow,
                               const uint32_t* tp_s, const uint32_t* qrp_s,
                               uint32_t* __restrict__ cigar, KswGapRun* __restrict__ runs, KswOut* out,
                               BtShared& sh, int lane, size_t arena_size = 0, bool reversed = false,
                               volatile int* progress = nullptr, int fold_q = 0)
{
    const int qlen = job.qlen, tlen = job.tlen, w = job.w, min_indel = job.min_indel;
    const bool wr_cigar = job.cigar_cap != 0 && lane == 0, wr_runs = lane == 0;
    int i = tlen - 1, j = qlen - 1, state = 0, n = 0, n_match = 0;
    uint32_t cur_op = 0xFFFFFFFFu, cur_len = 0, aln = 0;
    ScanState z;
    memset(&z, 0, sizeof z);
    while (i >= 0 && j >= 0) {
        const int r0 = i + j;
        if (kOverlap && lane == 0) *progress = r0;
        // ---- refill: rows r0 .. r0-63, two 16-byte blocks around the predicted column -------
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = lane + 32 * h, r = r0 - k;
            int b0 = -1, b1 = -1;
            if (r >= 0) {
                const RowGeom g = row_geom(r, qlen, tlen, w);
                const int bst = g.st0 >> 4, ben = g.en0 >> 4;
                int ic = i - ((k + 1) >> 1);
                ic = ic < 16 * bst ? 16 * bst : ic;
                ic = ic > 16 * ben + 15 ? 16 * ben + 15 : ic;
                b0 = ic >> 4;
                b1 = b0 + ((ic & 15) >= 8 ? 1 : -1);
                if (b1 < bst || b1 > ben) b1 = -1;
                const uint8_t* row = arena + ((kOverlap && reversed) ? arena_size - (size_t)(r + 1) * prow : (size_t)r * prow);
                size_t o0 = (size_t)tb_block_slot(b0, job.tb_blocks) * 16, o1 = b1 >= 0 ? (size_t)tb_block_slot(b1, job.tb_blocks) * 16 : 0;
                if (fold_q) {                           // column strips: rows folded modulo qlen (ksw_kernels.cuh)
                    const int pass = r / fold_q;
                    row = arena + (size_t)(r - pass * fold_q) * prow + (size_t)pass * 16;
                    o0 = (size_t)b0 * 16; o1 = b1 >= 0 ? (size_t)b1 * 16 : 0;
                }
                if (kOverlap) {
                    sh.cache[k][0] = __ldcg(reinterpret_cast<const uint4*>(row + o0));
                    if (b1 >= 0) sh.cache[k][1] = __ldcg(reinterpret_cast<const uint4*>(row + o1));
                } else {
                    sh.cache[k][0] = *reinterpret_cast<const uint4*>(row + o0);
                    if (b1 >= 0) sh.cache[k][1] = *reinterpret_cast<const uint4*>(row + o1);
                }
            }
            sh.blk[k][0] = b0; sh.blk[k][1] = b1;
        }
        __syncwarp();
        // ---- walk ----------------------------------------------------------------------------
        while (i >= 0 && j >= 0) {
            const int r = i + j, k = r0 - r;
            if (k >= kBtRows) break;
            const RowGeom g = row_geom(r, qlen, tlen, w);
            const int off = (g.st0 >> 4) << 4, off_end = ((g.en0 >> 4) << 4) + 15;
            int force = -1;
            if (i < off) force = 2;
            if (i > off_end) force = 1;
            uint32_t byte = 0;
            if (force < 0) {
                const int bi = i >> 4;
                const int slot = sh.blk[k][0] == bi ? 0 : sh.blk[k][1] == bi ? 1 : -1;
                if (slot < 0) break;                                   // not prefetched: refill here
                byte = reinterpret_cast<const uint8_t*>(&sh.cache[k][slot])[tb_byte_pos(i & 15)];
            }
            state = tb_next_state(state, byte, force);
            const uint32_t op = state == 0 ? 0u : (state == 1 || state == 3) ? 2u : 1u;
            bt_emit(op, 1, cur_op, cur_len, cigar, n, wr_cigar);
            ++aln;
            bool ident = false;
            if (op == 0) {                       // identical column? (calc_ident_perc, hbn_traceback_aux.c:4-19)
                const int qi = qlen - 1 - j;     // index into the reversed query
                const uint32_t tb = (tp_s[i >> 4] >> (2 * (i & 15))) & 3u;
                const uint32_t qb = (qrp_s[1 + (qi >> 4)] >> (2 * (qi & 15))) & 3u;
                ident = tb == qb;
                n_match += ident;
            }
            scan_column(z, op, ident, i, j, qlen, tlen, min_indel, runs, wr_runs);
            if (op == 0) { --i; --j; } else if (op == 2) --i; else --j;
        }
        __syncwarp();
    }
    if (kOverlap && lane == 0) *progress = -1;
    // leading gap (ksw2.h:146-147); it lies left of any 8-column window, so the scan ignores it
    if (i >= 0) { bt_emit(2, (uint32_t)(i + 1), cur_op, cur_len, cigar, n, wr_cigar); aln += i + 1; }
    if (j >= 0) { bt_emit(1, (uint32_t)(j + 1), cur_op, cur_len, cigar, n, wr_cigar); aln += j + 1; }
    bt_emit(0xFFFFFFFEu, 0, cur_op, cur_len, cigar, n, wr_cigar);      // flush
    __syncwarp();
    // the walk produced last-op-first; the reference returns first-op-first (ksw2.h:148-150)
    if (job.cigar_cap)
        for (int a = lane; a < n / 2; a += 32) {
            const uint32_t x = cigar[a], y = cigar[n - 1 - a];
            cigar[a] = y; cigar[n - 1 - a] = x;
        }
    if (lane == 0) {
        KswOut o;
        o.n_cigar = n; o.aln_size = (int)aln; o.n_match = n_match; o.status = 1;
        o.trunc_ok = z.have_right && !z.right_fail && z.s_valid && !z.s_at_end;
        o.qb = z.qb; o.qe = z.qe; o.sb = z.sb; o.se = z.se;
        o.eff_mat = z.s_eff_mat; o.eff_len = z.s_eff_len; o.trim_mat = z.s_mat; o.trim_len = z.s_len;
        o.n_runs = z.s_runs;
        *out = o;
    }
}

// ---------------------------------------------------------------------------------------------
// DP kernel
// ---------------------------------------------------------------------------------------------
// The only cross-thread dependency of a row is "column 15 of the block to my left, one row earlier".  Inside a CTA every
// thread leaves those values in shared memory and a barrier ends the row (a polled ring between the warps of one SM was
// measured slower, twice: the warps of the busiest scheduler set the pace either way and polling costs them issue slots).
// Across the CTAs of a cluster no barrier is cheap: the last lane of a CTA publishes into a small ring of tagged slots in
// the next CTA's shared memory (DSMEM), the first lane of that CTA picks the value up just before it needs it (pair 0 is
// processed last), and a consumed-rows counter flowing the other way keeps a producer from lapping its consumer.
constexpr int kMaxNT = 640;
constexpr int kMaxWarps = kMaxNT / 32;
constexpr int kRingD = 4;

struct __align__(16) CarrySlot { uint32_t xv, x2, tag, pad; };
struct PipeShared {
    CarrySlot inbox[kMaxWarps][kRingD];   // inbox[w]: rows published by the left neighbour of warp w
    uint32_t consumed[kMaxWarps];         // rows warp w has taken out of inbox[w]
    uint32_t right_consumed;              // cluster: rows taken by warp 0 of the CTA to the right
    int job;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Slot protocol without fences: a slot is ONE 16-byte shared-memory word {xv, x2, tag, -}, written
// with a single 128-bit store and polled with a single 128-bit load, so data and tag can never be
// seen torn; the consumed-rows counter is a single word.  volatile = re-read on every poll.
__device__ __forceinline__ void bar_active(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }
template <bool kCluster> __device__ __forceinline__ void st_slot(uint32_t addr, uint32_t a, uint32_t b, uint32_t tag)
{
    if (kCluster) asm volatile("st.volatile.shared::cluster.v4.u32 [%0], {%1, %2, %3, %3};" ::"r"(addr), "r"(a), "r"(b), "r"(tag) : "memory");
    else asm volatile("st.volatile.shared::cta.v4.u32 [%0], {%1, %2, %3, %3};" ::"r"(addr), "r"(a), "r"(b), "r"(tag) : "memory");
}
template <bool kCluster> __device__ __forceinline__ uint32_t ld_slot(uint32_t addr, uint32_t& a, uint32_t& b)
{
    uint32_t tag, pad;
    if (kCluster) asm volatile("ld.volatile.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(tag), "=r"(pad) : "r"(addr) : "memory");
    else asm volatile("ld.volatile.shared::cta.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(tag), "=r"(pad) : "r"(addr) : "memory");
    return tag;
}
template <bool kCluster> __device__ __forceinline__ void st_word(uint32_t addr, uint32_t v)
{
    if (kCluster) asm volatile("st.volatile.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
    else asm volatile("st.volatile.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
template <bool kCluster> __device__ __forceinline__ uint32_t ld_word(uint32_t addr)
{
    uint32_t v;
    if (kCluster) asm volatile("ld.volatile.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    else asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// address of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_rank(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// KswLaunchSlot behind the work counter: when the first CTA of the launch started, when the last one finished
__device__ __forceinline__ void stamp_start(int* counter) { atomicMin(reinterpret_cast<unsigned long long*>(counter + 2), global_ns()); }
__device__ __forceinline__ void stamp_end(int* counter) { atomicMax(reinterpret_cast<unsigned long long*>(counter + 4), global_ns()); }

// kCluster = false: one CTA per alignment (tiers of 32..160 threads, and wider ones whose sequences do not fit the double
//                   buffer of ksw_dp_overlap_kernel); warp 0 walks the traceback after the last row.
// kCluster = true:  one alignment on a thread-block cluster; the carry into the first and out of the last thread of a
//                   CTA crosses CTAs through the tagged slots described above.
template <bool kCluster>
__global__ void __launch_bounds__(kMaxNT)
ksw_dp_kernel(const KswJob* __restrict__ jobs, const int* __restrict__ order, int njobs,
              const uint32_t* __restrict__ pool, uint8_t* __restrict__ arena_base,
              const size_t* __restrict__ arena_off,
              uint32_t* __restrict__ cigar_pool, KswGapRun* __restrict__ run_pool,
              KswOut* __restrict__ outs, int* counter)
{
    // dynamic shared memory: 16 bytes per thread of the launch (column-15 values by row parity, one address bit apart),
    // then the packed sequences.  One-warp CTAs of the narrow tiers share an SM by the dozen: nothing here is sized for
    // the widest launch.
    extern __shared__ __align__(16) uint32_t dyn_s[];
    uint2 (*carry_all)[2] = reinterpret_cast<uint2 (*)[2]>(dyn_s);
    uint32_t* const seq_s = dyn_s + 4 * blockDim.x;
    __shared__ PipeShared ps;
    __shared__ BtShared bt;
    constexpr int kChunk = kCluster ? 256 : 64;             // rows per band-edge table
    __shared__ int4 edge_tab[kChunk];                       // RowEdges of the current chunk of rows
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto edges_from_table = [&](int rr) -> RowEdges {
        const int4 v = edge_tab[rr & (kChunk - 1)];
        RowEdges e; e.st0 = v.x; e.wr = v.y; e.ben = v.z; e.bst_prev = v.w;
        return e;
    };
    // Programmatic dependent launch: the launch classes of a batch are chained in one stream, each allowed to start
    // as soon as every CTA of the class before it is resident (lesvgpu.cu: run_jobs).  Nothing is waited for.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) stamp_start(counter);
    const uint32_t ncta = kCluster ? cluster_nctarank() : 1u, rank = kCluster ? cluster_ctarank() : 0u;
    // The queue is sorted by decreasing traceback size and cluster i starts with job i, so the arena
    // of cluster i only has to hold job i: everything it pulls later is smaller.
    const int cluster_id = (int)(blockIdx.x / ncta);
    uint8_t* const arena = arena_base + arena_off[cluster_id];
    bool first = true;

    for (;;) {
        if (kCluster) {
            if (rank == 0 && tid == 0) {
                const int j = first ? cluster_id : atomicAdd(counter, 1);
                for (uint32_t rr = 0; rr < ncta; ++rr) st_word<true>(map_rank(smem_addr(&ps.job), rr), (uint32_t)j);
            }
            __syncthreads();
            cluster_sync_all();
        } else {
            if (tid == 0) ps.job = first ? cluster_id : atomicAdd(counter, 1);
            __syncthreads();
        }
        first = false;
        const int ji = ps.job;
        if (ji >= njobs) break;
        const int jid = order[ji];
        const KswJob job = jobs[jid];
        const int qlen = job.qlen, tlen = job.tlen, w = job.w, tb_blocks = job.tb_blocks;
        const int ntw = (int)tp_words(tlen), nqw = (int)qrp_words(qlen);
        uint32_t* tp_s = seq_s;
        uint32_t* qrp_s = seq_s + ntw;
        for (int x = tid; x < ntw; x += blockDim.x) tp_s[x] = pool[job.tp_off + x];
        for (int x = tid; x < nqw; x += blockDim.x) qrp_s[x] = pool[job.qrp_off + x];
        if (kCluster) {
            for (int x = tid; x < kMaxWarps * kRingD; x += blockDim.x) ps.inbox[x / kRingD][x % kRingD].tag = 0xFFFFFFFFu;
            if (tid < kMaxWarps) ps.consumed[tid] = 0;
            if (tid == 0) ps.right_consumed = 0;
        }
        __syncthreads();
        if (kCluster) cluster_sync_all();          // every CTA's ring is reset before anyone publishes

        // this alignment uses the first NT threads of every CTA (the launch is sized for its tier)
        const int NT = job.nt, NW = NT >> 5, NTT = NT * (int)ncta;
        ThreadCtx c;
        thread_ctx_reset(c, (int)rank * NT + tid);
        Carry prev_out = carry_init();
        const size_t prow = (size_t)tb_blocks * 16;
        const int nrow_all = qlen + tlen - 1;
        if (!kCluster) {
            // Lean barrier loop: every thread leaves its column-15 values in carry_all[thread][row parity] and loads its left
            // neighbour's of the row before at the top of the next row (no shuffle, no lane-0 / lane-31 special cases).
            uint32_t carry_rd = smem_addr(&carry_all[tid ? tid - 1 : NT - 1][1]), carry_wr = smem_addr(&carry_all[tid][0]);
            uint8_t* rowp = arena;
            if (warp < NW)                          // the warps that hold blocks, on a named barrier of their own
                for (int r0 = 0; r0 < nrow_all; r0 += kChunk) {
                    const int r1 = r0 + kChunk < nrow_all ? r0 + kChunk : nrow_all;
                    for (int x = tid; x < r1 - r0; x += NT) {          // band edges of the chunk's rows, once for the CTA
                        const RowEdges e = row_edges(r0 + x, qlen, tlen, w);
                        edge_tab[x] = make_int4(e.st0, e.wr, e.ben, e.bst_prev);
                    }
                    bar_active(NT);
                    for (int r = r0; r < r1; ++r, rowp += prow) {
                        Carry cin;
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cin.xv), "=r"(cin.x2) : "r"(carry_rd) : "memory");
                        Carry cout;
                        uint32_t tb[4];
                        if (thread_row_step<true>(c, NT, r, qlen, tlen, w, tb_blocks, tp_s, qrp_s,
                                                  [&]() -> Carry { return cin; }, cout, tb, edges_from_table)) {
                            *reinterpret_cast<uint4*>(rowp + c.tb_off) = make_uint4(tb[0], tb[1], tb[2], tb[3]);
                            prev_out = cout;
                        }
                        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(carry_wr), "r"(prev_out.xv), "r"(prev_out.x2) : "memory");
                        carry_rd ^= 8u; carry_wr ^= 8u;
                        bar_active(NT);
                    }
                }
        } else {
            // Cluster: inside a CTA the same lean barrier loop as above (the polled ring between warps of one SM costs
            // issue slots of the schedulers that are the bottleneck: 3.5 instead of 5.5 GCUPS per SM); only the carry INTO
            // warp 0 and OUT OF the last warp crosses CTAs, through the tagged DSMEM slots.  Lane 0 of warp 0 fetches it
            // as late as possible (just before pair 0, the last pair of the row), so the CTA to the left may lag by most of
            // a row before anybody waits; a consumed-rows counter flowing back keeps it from lapping the ring.
            const uint32_t in_addr = smem_addr(&ps.inbox[0][0]);                                   // written by the CTA to the left
            const uint32_t cons_addr = map_rank(smem_addr(&ps.right_consumed), (rank + ncta - 1) % ncta);
            const uint32_t out_addr = map_rank(smem_addr(&ps.inbox[0][0]), (rank + 1) % ncta);
            const uint32_t watch_addr = smem_addr(&ps.right_consumed);
            const bool first_lane = warp == 0 && lane == 0, last_lane = warp == NW - 1 && lane == 31;
            uint32_t carry_rd = smem_addr(&carry_all[tid ? tid - 1 : NT - 1][1]), carry_wr = smem_addr(&carry_all[tid][0]);
            uint8_t* rowp = arena;
            if (warp < NW)
                for (int r0 = 0; r0 < nrow_all; r0 += kChunk) {
                    const int r1 = r0 + kChunk < nrow_all ? r0 + kChunk : nrow_all;
                    for (int x = tid; x < r1 - r0; x += NT) {
                        const RowEdges e = row_edges(r0 + x, qlen, tlen, w);
                        edge_tab[x] = make_int4(e.st0, e.wr, e.ben, e.bst_prev);
                    }
                    bar_active(NT);
                    for (int r = r0; r < r1; ++r, rowp += prow) {
                        Carry cin;                             // thread 0 reads a slot nobody writes and takes consume() instead
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cin.xv), "=r"(cin.x2) : "r"(carry_rd) : "memory");
                        bool fetched = false;
                        auto consume = [&]() -> Carry {        // first lane of the CTA: take row r-1 out of the inbox
                            Carry v = carry_init();
                            fetched = true;
                            if (r == 0) return v;
                            const uint32_t slot = in_addr + (uint32_t)((r - 1) & (kRingD - 1)) * (uint32_t)sizeof(CarrySlot);
                            while (ld_slot<false>(slot, v.xv, v.x2) != (uint32_t)(r - 1)) { }
                            st_word<true>(cons_addr, (uint32_t)r);
                            return v;
                        };
                        Carry cout;
                        uint32_t tb[4];
                        const bool did = thread_row_step(c, NTT, r, qlen, tlen, w, tb_blocks, tp_s, qrp_s,
                                                         [&]() -> Carry { return first_lane ? consume() : cin; }, cout, tb,
                                                         edges_from_table);
                        if (did) {
                            *reinterpret_cast<uint4*>(rowp + c.tb_off) = make_uint4(tb[0], tb[1], tb[2], tb[3]);
                            prev_out = cout;
                        }
                        if (first_lane && !fetched) consume();                     // flow control only
                        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(carry_wr), "r"(prev_out.xv), "r"(prev_out.x2) : "memory");
                        carry_rd ^= 8u; carry_wr ^= 8u;
                        if (last_lane) {                                           // publish row r to the CTA on the right
                            if (r >= kRingD)
                                while ((int)ld_word<false>(watch_addr) < r - kRingD + 1) { }
                            const uint32_t slot = out_addr + (uint32_t)(r & (kRingD - 1)) * (uint32_t)sizeof(CarrySlot);
                            st_slot<true>(slot, prev_out.xv, prev_out.x2, (uint32_t)r);
                        }
                        __syncwarp();              // the polling lanes are back before the aligned barrier
                        bar_active(NT);
                    }
                }
        }
        __syncthreads();
        if (kCluster) { __threadfence(); cluster_sync_all(); }
        if (rank == 0 && warp == 0)
            backtrack_warp(job, arena, prow, tp_s, qrp_s, cigar_pool + job.cigar_off,
                           run_pool + job.sig_off, outs + jid, bt, lane);
        __syncthreads();
        if (kCluster) cluster_sync_all();
    }
    if (tid == 0) stamp_end(counter);
}

// ---------------------------------------------------------------------------------------------
// Narrow bands: several alignments per warp
// ---------------------------------------------------------------------------------------------
// A band of 64 columns lives in 7 or 8 sixteen-column blocks, one of 128..192 in at most 16: with one block per thread a
// warp that holds ONE such alignment computes with a quarter or half of its lanes (BASELINE configs[4]: 63-107 GCUPS at band
// 64).  Here a warp is cut into 32 / G groups of G lanes and every group runs its own alignment: its own queue entry, row
// counter, geometry and traceback arena.  The groups of a warp step through their anti-diagonals together; the carry of
// a block comes from the lane to its left in the group (a rotation inside the group, so the ring wrap needs no shared
// memory), nothing is synchronised beyond the warp, and the packed sequences are read straight from the pool (they are
// tiny next to the traceback and stay in L1).  A group that has finished its last row is walked by the whole warp
// (backtrack_warp, which keeps all 32 lanes busy with the row prefetch), the other groups wait for that walk.
template <int G>
__global__ void __launch_bounds__(32, 20)
ksw_dp_group_kernel(const KswJob* __restrict__ jobs, const int* __restrict__ order, int njobs,
                    const uint32_t* __restrict__ pool, uint8_t* __restrict__ arena_base,
                    const size_t* __restrict__ arena_off, uint32_t* __restrict__ cigar_pool,
                    KswGapRun* __restrict__ run_pool, KswOut* __restrict__ outs, int* counter)
{
    __shared__ BtShared bt;
    constexpr int NG = 32 / G;
    constexpr unsigned kFull = 0xFFFFFFFFu;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = (int)threadIdx.x, gl = lane & (G - 1), grp = lane / G;
    const unsigned gmask = (G == 32 ? kFull : ((1u << G) - 1u)) << (grp * G);
    const int group_id = (int)blockIdx.x * NG + grp;
    if (lane == 0) stamp_start(counter);
    uint8_t* const arena = arena_base + arena_off[group_id];
    const int left = grp * G + ((gl + G - 1) & (G - 1));           // the lane that holds the block to my left (ring inside the group)
    bool have = false, first = true, exhausted = false;
    int jid = 0, r = 0, nrow = 0, qlen = 0, tlen = 0, w = 0, tb_blocks = 1;
    size_t prow = 16;
    const uint32_t *tp = pool, *qrp = pool;
    ThreadCtx c;
    thread_ctx_reset(c, gl);
    Carry prev_out = carry_init();
    for (;;) {
        if (!have && !exhausted) {                                  // uniform inside a group
            int j = 0;
            if (gl == 0) j = first ? group_id : atomicAdd(counter, 1);
            j = __shfl_sync(gmask, j, grp * G);
            first = false;
            if (j < njobs) {
                jid = order[j];
                const KswJob job = jobs[jid];
                qlen = job.qlen; tlen = job.tlen; w = job.w; tb_blocks = job.tb_blocks;
                prow = (size_t)tb_blocks * 16;
                nrow = qlen + tlen - 1;
                tp = pool + job.tp_off; qrp = pool + job.qrp_off;
                thread_ctx_reset(c, gl);
                prev_out = carry_init();
                r = 0;
                have = true;
            } else exhausted = true;
        }
        if (!__any_sync(kFull, have)) break;
        if (have && r < nrow) {
            Carry cin;
            cin.xv = __shfl_sync(gmask, prev_out.xv, left);
            cin.x2 = __shfl_sync(gmask, prev_out.x2, left);
            Carry cout;
            uint32_t tb[4];
            if (thread_row_step<true, false>(c, G, r, qlen, tlen, w, tb_blocks, tp, qrp, [&]() -> Carry { return cin; }, cout, tb)) {
                *reinterpret_cast<uint4*>(arena + (size_t)r * prow + c.tb_off) = make_uint4(tb[0], tb[1], tb[2], tb[3]);
                prev_out = cout;
            }
            ++r;
        }
        __syncwarp();
        unsigned dm = __ballot_sync(kFull, have && r >= nrow && gl == 0);
        if (dm) __threadfence_block();                              // the rows were written by other lanes of this warp
        while (dm) {
            const int leader = __ffs((int)dm) - 1;
            dm &= dm - 1;
            const int wj = __shfl_sync(kFull, jid, leader);
            const KswJob job = jobs[wj];
            uint8_t* const wa = arena_base + arena_off[(int)blockIdx.x * NG + leader / G];
            backtrack_warp(job, wa, (size_t)job.tb_blocks * 16, pool + job.tp_off, pool + job.qrp_off, cigar_pool + job.cigar_off,
                           run_pool + job.sig_off, outs + wj, bt, lane);
            __syncwarp();
            if (grp == leader / G) have = false;
        }
    }
    if (lane == 0) stamp_end(counter);
}

// ---------------------------------------------------------------------------------------------
// Single-CTA alignments, DP and traceback walk overlapped
// ---------------------------------------------------------------------------------------------
// The walk of an alignment (one warp, ~7 % of its DP time, the other warps idle) does not have to finish before the
// CTA starts the next alignment: an extra warp walks alignment m-1 while the DP warps fill the traceback of
// alignment m.  No second arena is needed: consecutive alignments store their rows in opposite directions
// (even ones bottom-up from offset 0, odd ones top-down from the end of the arena), so the walk -- which runs from the
// last row to the first, an order of magnitude faster than the DP -- always moves ahead of the write front that
// follows it.  The DP warps check every 256 rows that this is really so (the walker publishes the row it still
// needs) and the packed sequences are double-buffered in shared memory.  The DP warps synchronise on named
// barrier 1, the walker takes part in nothing.
#ifndef KSW_WALK_SLEEP_NS
#define KSW_WALK_SLEEP_NS 4000      // the walker's poll for the next alignment shares a scheduler with DP warps: poll rarely
#endif
__device__ __forceinline__ void bar_dp(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }
#define KSW_PRAGMA_(x) _Pragma(#x)
#define KSW_PRAGMA(x) KSW_PRAGMA_(x)
#ifdef KSW_ROW_UNROLL
#define KSW_ROW_LOOP_PRAGMA KSW_PRAGMA(unroll KSW_ROW_UNROLL)
#else
#define KSW_ROW_LOOP_PRAGMA KSW_PRAGMA(unroll 1)
#endif
__device__ __forceinline__ void bar_rows(int n) { asm volatile("bar.sync 2, %0;" ::"r"(n) : "memory"); }

template <int kTier>                                    // CTA = kTier DP threads + the walker warp: fixes the register budget
__global__ void __launch_bounds__(kTier + 32, kTier <= kTierLow ? 2 : 1)
ksw_dp_overlap_kernel(const KswJob* __restrict__ jobs, const int* __restrict__ order, int njobs,
                      const uint32_t* __restrict__ pool, uint8_t* __restrict__ arena_base,
                      const size_t* __restrict__ arena_off /* n_ctas + 1 entries */, uint32_t half_words,
                      uint32_t* __restrict__ cigar_pool, KswGapRun* __restrict__ run_pool,
                      KswOut* __restrict__ outs, int* counter)
{
    extern __shared__ uint32_t seq_s[];                 // two halves of half_words words
    __shared__ BtShared bt;
    __shared__ __align__(16) uint2 carry_all[kTier][2];    // per thread and row parity: the two parities differ in one address bit
    __shared__ int4 edge_tab[256];                          // RowEdges of the current chunk of 256 rows
    __shared__ int sh_job;
    __shared__ int wk_jid[2];
    __shared__ volatile int wk_posted, wk_done, wk_finished, wk_row;
    __shared__ volatile unsigned long long sh_prow, sh_prev_prow, sh_asize;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ndp = (int)blockDim.x - 32;               // DP threads; the last warp walks
    const int cta = (int)blockIdx.x;
    uint8_t* const arena = arena_base + arena_off[cta];
    const size_t asize = arena_off[cta + 1] - arena_off[cta];
    if (tid == 0) { wk_posted = 0; wk_done = 0; wk_finished = 0; wk_row = -1; sh_prev_prow = 0; stamp_start(counter); }
    __syncthreads();
    if (tid >= ndp) {
        // ---- walker ---------------------------------------------------------------------------------------
        for (int n = 0;; ++n) {
            int fin = 0;
            if (lane == 0) {
                while (wk_posted <= n && !wk_finished) __nanosleep(KSW_WALK_SLEEP_NS);
                fin = wk_posted <= n;
            }
            fin = __shfl_sync(0xFFFFFFFFu, fin, 0);
            if (fin) break;
            __threadfence_block();
            const int jid = wk_jid[n & 1];
            const KswJob job = jobs[jid];
            const uint32_t* tp_s = seq_s + (size_t)(n & 1) * half_words;
            const uint32_t* qrp_s = tp_s + tp_words(job.tlen);
            backtrack_warp<true>(job, arena, (size_t)job.tb_blocks * 16, tp_s, qrp_s, cigar_pool + job.cigar_off,
                                 run_pool + job.sig_off, outs + jid, bt, lane, asize, (n & 1) != 0, &wk_row);
            __syncwarp();
            if (lane == 0) { __threadfence(); wk_done = n + 1; }
        }
        if (lane == 0) stamp_end(counter);
        return;
    }
    // ---- DP warps ---------------------------------------------------------------------------------------------
    // The walker is the block's last warp and lands on the warp scheduler of warp 0; a CTA rarely needs all of its DP
    // warps (NW <= ndp / 32), so the logical warps are rotated by one: physical warp 0 becomes the LAST logical warp,
    // the one most likely to sit an alignment out, and the walker shares its scheduler with one idle warp.
    const int nwdp = ndp >> 5;
    const int lwarp = (warp + nwdp - 1) % nwdp, ltid = lwarp * 32 + lane;
    bool first = true;
    for (int m = 0;; ++m) {
        if (tid == 0) sh_job = first ? cta : atomicAdd(counter, 1);
        bar_dp(ndp);
        first = false;
        const int ji = sh_job;
        if (ji >= njobs) break;
        if (warp == 0)                                  // walk m-2 is over: its sequence buffer and rows are free
            for (;;) {                                  // (whole-warp poll: no divergence in front of the aligned barrier)
                int ok = lane == 0 ? (wk_done >= m - 1) : 0;
                ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
                if (ok) break;
                __nanosleep(500);
            }
        bar_dp(ndp);
        const int jid = order[ji];
        const KswJob job = jobs[jid];
        const int qlen = job.qlen, tlen = job.tlen, w = job.w, tb_blocks = job.tb_blocks;
        const int ntw = (int)tp_words(tlen), nqw = (int)qrp_words(qlen);
        uint32_t* tp_s = seq_s + (size_t)(m & 1) * half_words;
        uint32_t* qrp_s = tp_s + ntw;
        for (int x = tid; x < ntw; x += ndp) tp_s[x] = pool[job.tp_off + x];
        for (int x = tid; x < nqw; x += ndp) qrp_s[x] = pool[job.qrp_off + x];
        bar_dp(ndp);
        const int NT = job.nt, NW = NT >> 5;
        const bool reversed = (m & 1) != 0;
        const size_t prow = (size_t)tb_blocks * 16;
        ThreadCtx c;
        thread_ctx_reset(c, ltid);
        Carry prev_out = carry_init();
        const int nrow_all = qlen + tlen - 1;
        if (tid == 0) { sh_prow = (unsigned long long)prow; sh_asize = (unsigned long long)asize; }
        // row r lives at rowp; consecutive rows are prow apart, upwards or downwards
        uint8_t* rowp = arena + (reversed ? asize - prow : (size_t)0);
        const long long rstep = reversed ? -(long long)prow : (long long)prow;
        // Only the warps that hold blocks of this alignment run the row loop, on a named barrier of their own (2) sized for
        // them; the others wait at the all-warps barrier (1) behind it.  Rows come in chunks of 256: before a chunk the
        // first logical warp checks that its rows do not reach into what walk m-1 still needs (rows <= wk_row of it).
        if (lwarp < NW) {
            // column-15 values of every thread by row parity: a thread stores its own and loads its left neighbour's of the
            // row before (no shuffle, no lane-0 / lane-31 special cases); both addresses flip one bit per row
            uint32_t carry_rd = smem_addr(&carry_all[ltid ? ltid - 1 : NT - 1][1]), carry_wr = smem_addr(&carry_all[ltid][0]);
            for (int r0 = 0; r0 < nrow_all; r0 += 256) {
                const int r1 = r0 + 256 < nrow_all ? r0 + 256 : nrow_all;
                if (m > 0 && lwarp == 0) {
                    const unsigned long long mine = (unsigned long long)r1 * sh_prow;
                    for (;;) {                          // (whole-warp poll: no divergence in front of the aligned barrier)
                        int ok = 0;
                        if (lane == 0) {
                            const int need = wk_done >= m ? -1 : wk_row;
                            ok = (need < 0 && wk_done >= m) ||
                                 (need >= 0 && mine + (unsigned long long)(need + 1) * sh_prev_prow <= sh_asize);
                        }
                        ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
                        if (ok) break;
                        __nanosleep(500);
                    }
                }
                // the band edges of the chunk's rows, once for the CTA: the warps that hold edge blocks are the last to
                // arrive at every row barrier, so what they do per row is kept to a table look-up
                for (int x = ltid; x < r1 - r0; x += NT) {
                    const RowEdges e = row_edges(r0 + x, qlen, tlen, w);
                    edge_tab[x] = make_int4(e.st0, e.wr, e.ben, e.bst_prev);
                }
                bar_rows(NT);
                KSW_ROW_LOOP_PRAGMA
                for (int r = r0; r < r1; ++r, rowp += rstep) {
                    Carry cin;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cin.xv), "=r"(cin.x2) : "r"(carry_rd) : "memory");
                    Carry cout;
                    uint32_t tb[4];
                    if (thread_row_step<true>(c, NT, r, qlen, tlen, w, tb_blocks, tp_s, qrp_s,
                                              [&]() -> Carry { return cin; }, cout, tb,
                                              [&](int rr) -> RowEdges {
                                                  const int4 v = edge_tab[rr & 255];
                                                  RowEdges e; e.st0 = v.x; e.wr = v.y; e.ben = v.z; e.bst_prev = v.w;
                                                  return e;
                                              })) {
                        *reinterpret_cast<uint4*>(rowp + c.tb_off) = make_uint4(tb[0], tb[1], tb[2], tb[3]);
                        prev_out = cout;
                    }
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(carry_wr), "r"(prev_out.xv), "r"(prev_out.x2) : "memory");
                    carry_rd ^= 8u; carry_wr ^= 8u;
                    bar_rows(NT);
                }
            }
        }
        __threadfence();                                 // the traceback is read by the walker through L2
        bar_dp(ndp);
        if (tid == 0) {
            wk_jid[m & 1] = jid;
            wk_row = nrow_all - 1;                       // until the walker publishes its own position
            __threadfence_block();
            wk_posted = m + 1;
        }
        if (tid == 0) sh_prev_prow = (unsigned long long)prow;
    }
    if (tid == 0) { __threadfence_block(); wk_finished = 1; }
}

// ---------------------------------------------------------------------------------------------
// Full-matrix alignments as column strips
// ---------------------------------------------------------------------------------------------
// An anti-diagonal of the WHOLE matrix (qx2svr's dual-HSP chaining runs without a band) grows from one cell to
// min(qlen, tlen) cells and shrinks again, so a cluster sized for the widest row idles half of the time.  Here the
// columns are cut into strips of kStripNT blocks; a strip is an ordinary single-CTA job whose threads keep their block
// for the whole pass.  Strip s takes the column-15 values of strip s-1 from a per-row edge buffer in global memory
// (written once per row by the last thread of strip s-1, prefetched one row ahead by the first thread of strip s) and
// may work on row r once strip s-1 has finished row r-1: progress counters, published and checked every 64 rows
// (polled by a whole warp: a single lane spinning in front of __syncthreads() lets the rest of its warp run through
// the aligned barrier -- undefined behaviour that showed up as a slightly different alignment).
// Strips are handed out in (strip, alignment) order, so the predecessor of a strip is always running or done -- no
// CTA ever waits for work that has not been started.  A CTA that finishes a strip goes on to the next one of any
// alignment; whoever finishes the last strip of an alignment walks its traceback.
constexpr int kStripNT = kStripBlocks;
struct StripJob {
    int jid;                 // alignment (index into jobs)
    int s, n_strips;
    int prog_off;            // progress[prog_off + s], done counter at progress[prog_off + n_strips]
    unsigned long long arena_off;   // this alignment's traceback
    unsigned long long edge_off;    // edges[edge_off + s * nrow + r]: column-15 values of strip s in row r
};

__global__ void __launch_bounds__(kStripNT, kStripCtasPerSm)
ksw_dp_strip_kernel(const KswJob* __restrict__ jobs, const StripJob* __restrict__ sjobs, int n_sjobs,
                    const uint32_t* __restrict__ pool, uint8_t* __restrict__ arena_base, uint2* __restrict__ edges,
                    int* __restrict__ progress, uint32_t* __restrict__ cigar_pool, KswGapRun* __restrict__ run_pool,
                    KswOut* __restrict__ outs, int* counter)
{
    extern __shared__ uint32_t seq_s[];
    __shared__ BtShared bt;
    __shared__ __align__(16) uint2 carry_all[kStripNT][2];   // column-15 values per thread and row parity
    __shared__ int sh_job, sh_last;
    __shared__ int bt_dummy;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool first = true;
    if (tid == 0) stamp_start(counter);
    for (;;) {
        if (tid == 0) sh_job = first ? (int)blockIdx.x : atomicAdd(counter, 1);
        __syncthreads();
        first = false;
        const int si = sh_job;
        if (si >= n_sjobs) { if (tid == 0) stamp_end(counter); break; }
        const StripJob sj = sjobs[si];
        const KswJob job = jobs[sj.jid];
        const int qlen = job.qlen, tlen = job.tlen, w = job.w, tb_blocks = job.tb_blocks;
        const int ntw = (int)tp_words(tlen), nqw = (int)qrp_words(qlen);
        uint32_t* tp_s = seq_s;
        uint32_t* qrp_s = seq_s + ntw;
        for (int x = tid; x < ntw; x += kStripNT) tp_s[x] = pool[job.tp_off + x];
        for (int x = tid; x < nqw; x += kStripNT) qrp_s[x] = pool[job.qrp_off + x];
        __syncthreads();
        const int nrow_all = qlen + tlen - 1;
        const int B0 = sj.s * kStripNT;
        // rows in which some block of the strip can be live (a superset is fine: dead rows fall through)
        const int c0 = 16 * B0 - 16, c1e = 16 * (B0 + kStripNT - 1) + 15;
        int r_lo = c0 > 2 * c0 - w ? c0 : 2 * c0 - w;
        r_lo = r_lo < 0 ? 0 : r_lo;
        long long hi = (long long)c1e + qlen - 1, hi2 = 2ll * c1e + w;
        hi = hi < hi2 ? hi : hi2;
        const int r_hi = hi < nrow_all - 1 ? (int)hi : nrow_all - 1;
        uint8_t* const arena = arena_base + sj.arena_off;
        const size_t prow = fold_row_bytes(qlen, tlen);  // rows folded modulo qlen: row r at physical row r % qlen, block B at slot B + r / qlen
        const uint2* edge_in = edges + sj.edge_off + (size_t)(sj.s - 1) * nrow_all;      // strip s-1 (s > 0)
        uint2* edge_out = edges + sj.edge_off + (size_t)sj.s * nrow_all;
        int* prog = progress + sj.prog_off;
        ThreadCtx c;
        thread_ctx_reset(c, B0 + tid);
        Carry prev_out = carry_init();
        uint2 edge_next = make_uint2(0u, 0u);          // value for the NEXT row, loaded one row ahead (thread 0)
        uint32_t carry_rd = smem_addr(&carry_all[tid ? tid - 1 : 0][1]), carry_wr = smem_addr(&carry_all[tid][0]);
        int frow = r_lo % qlen;                         // physical row of r
        uint8_t* rowp = arena + (size_t)frow * prow + (size_t)(B0 + tid + r_lo / qlen) * 16;     // this thread's 16 bytes of row r
        for (int r = r_lo; r <= r_hi; ++r) {
            if (((r - r_lo) & 63) == 0) {
                if (sj.s > 0 && warp == 0) {            // rows r .. r+63 read rows r-1 .. r+62 of the strip to the left
                    int need = r + 63;                  // (the whole warp polls: no divergence in front of the barrier)
                    if (need > nrow_all) need = nrow_all;
                    for (;;) {
                        int v = lane == 0 ? atomicAdd(prog + sj.s - 1, 0) : 0;
                        v = __shfl_sync(0xFFFFFFFFu, v, 0);
                        if (v >= need) break;
                        __nanosleep(200);
                    }
                    __threadfence();
                    if (r > 0) {
                        const uint2 e = __ldcv(edge_in + (r - 1));
                        if (tid == 0) edge_next = e;
                    }
                }
                if (r > r_lo) {                         // publish: rows < r of this strip are complete
                    __threadfence();
                    __syncthreads();
                    if (tid == 0) atomicExch(prog + sj.s, r);
                } else __syncthreads();
            }
            const uint2 edge_now = edge_next;
            if (sj.s > 0 && tid == 0 && r + 1 <= r_hi) edge_next = __ldcv(edge_in + r);   // for row r+1
            Carry cin;                                  // left neighbour's column-15 values of the row before (thread 0: the edge buffer)
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cin.xv), "=r"(cin.x2) : "r"(carry_rd) : "memory");
            if (tid == 0) { cin.xv = edge_now.x; cin.x2 = edge_now.y; }
            Carry cout;
            uint32_t tb[4];
            if (thread_row_step<true>(c, 1 << 24, r, qlen, tlen, w, tb_blocks, tp_s, qrp_s,
                                      [&]() -> Carry { return cin; }, cout, tb)) {
                *reinterpret_cast<uint4*>(rowp) = make_uint4(tb[0], tb[1], tb[2], tb[3]);
                prev_out = cout;
                if (tid == kStripNT - 1) edge_out[r] = make_uint2(cout.xv, cout.x2);
            }
            rowp += prow;
            if (++frow == qlen) { frow = 0; rowp += 16 - (ptrdiff_t)((size_t)qlen * prow); }     // next pass: back to physical row 0, one slot to the right
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(carry_wr), "r"(prev_out.xv), "r"(prev_out.x2) : "memory");
            carry_rd ^= 8u; carry_wr ^= 8u;
            __syncthreads();
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            atomicExch(prog + sj.s, 0x7fffffff);
            __threadfence();
            sh_last = atomicAdd(prog + sj.n_strips, 1) == sj.n_strips - 1;
        }
        __syncthreads();
        if (sh_last) {                                  // every strip of this alignment is done: walk it
            __threadfence();
            if (warp == 0)
                backtrack_warp<true>(job, arena, prow, tp_s, qrp_s, cigar_pool + job.cigar_off, run_pool + job.sig_off,
                                     outs + sj.jid, bt, lane, 0, false, (volatile int*)&bt_dummy, qlen);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// wide bands (window > 512 blocks: full-band retries, chained dual HSPs)
// ---------------------------------------------------------------------------------------------
// Same per-block code, but a thread steps through several band slots per anti-diagonal and the
// slot contexts live in global memory (L2-resident, structure of arrays so that a warp's loads
// coalesce).  ~10 % extra traffic on a path that is rare by construction.
constexpr int kWideNT = 512;
constexpr int kCtxWords = sizeof(ThreadCtx) / 4;
static_assert(sizeof(ThreadCtx) % 4 == 0, "ThreadCtx is an array of 32-bit words");

__device__ __forceinline__ void ctx_load(ThreadCtx& c, const uint32_t* base, int nslot, int s)
{
    uint32_t* w = reinterpret_cast<uint32_t*>(&c);
#pragma unroll
    for (int k = 0; k < kCtxWords; ++k) w[k] = base[(size_t)k * nslot + s];
}
__device__ __forceinline__ void ctx_store(const ThreadCtx& c, uint32_t* base, int nslot, int s)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&c);
#pragma unroll
    for (int k = 0; k < kCtxWords; ++k) base[(size_t)k * nslot + s] = w[k];
}

__global__ void __launch_bounds__(kWideNT, 1)
ksw_dp_wide_kernel(const KswJob* __restrict__ jobs, const int* __restrict__ order, int njobs,
                   const uint32_t* __restrict__ pool, uint8_t* __restrict__ arena_base, size_t arena_stride,
                   uint32_t* __restrict__ scratch_base, size_t scratch_stride_words, int max_slots,
                   uint32_t* __restrict__ cigar_pool, KswGapRun* __restrict__ run_pool,
                   KswOut* __restrict__ outs, int* counter)
{
    extern __shared__ uint32_t seq_s[];
    __shared__ int s_job;
    __shared__ BtShared bt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* const arena = arena_base + (size_t)blockIdx.x * arena_stride;
    uint32_t* const scratch = scratch_base + (size_t)blockIdx.x * scratch_stride_words;

    if (tid == 0) stamp_start(counter);
    for (;;) {
        if (tid == 0) s_job = atomicAdd(counter, 1);
        __syncthreads();
        const int ji = s_job;
        if (ji >= njobs) { if (tid == 0) stamp_end(counter); break; }
        const int jid = order[ji];
        const KswJob job = jobs[jid];
        const int qlen = job.qlen, tlen = job.tlen, w = job.w;
        const int ntw = (int)tp_words(tlen), nqw = (int)qrp_words(qlen);
        uint32_t* tp_s = seq_s;
        uint32_t* qrp_s = seq_s + ntw;
        for (int x = tid; x < ntw; x += kWideNT) tp_s[x] = pool[job.tp_off + x];
        for (int x = tid; x < nqw; x += kWideNT) qrp_s[x] = pool[job.qrp_off + x];
        int nslot = window_blocks(tlen, w);
        nslot = (nslot + kWideNT - 1) / kWideNT * kWideNT;        // <= max_slots by construction
        uint32_t* ctxs = scratch;                                  // [kCtxWords][nslot]
        uint2* pub = reinterpret_cast<uint2*>(scratch + (size_t)kCtxWords * max_slots);   // [2][nslot]
        {
            for (int sidx = tid; sidx < nslot; sidx += kWideNT) {
                ThreadCtx c;
                thread_ctx_reset(c, sidx);
                ctx_store(c, ctxs, nslot, sidx);
            }
        }
        __syncthreads();
        const int nrow = qlen + tlen - 1;
        const size_t prow = (size_t)job.tb_blocks * 16;
        for (int r = 0; r < nrow; ++r) {
            const RowGeom g = row_geom(r, qlen, tlen, w);
            const int bst = g.st0 >> 4, bwr = g.wr >> 4;
            uint2* pub_r = pub + (size_t)(r & 1) * nslot;
            const uint2* pub_p = pub + (size_t)((r + 1) & 1) * nslot;
            // visit only slots whose block can be live this row: blocks bst..bwr
            for (int B = bst + tid; B <= bwr; B += kWideNT) {
                const int sidx = B % nslot;
                ThreadCtx c;
                ctx_load(c, ctxs, nslot, sidx);
                const uint2 pv = pub_p[(sidx + nslot - 1) % nslot];
                Carry cin; cin.xv = pv.x; cin.x2 = pv.y;
                Carry cout;
                uint32_t tb[4];
                if (thread_row_step<false, false>(c, nslot, r, qlen, tlen, w, job.tb_blocks, tp_s, qrp_s,
                                                  [&]() -> Carry { return cin; }, cout, tb)) {
                    *reinterpret_cast<uint4*>(arena + (size_t)r * prow + (size_t)c.tb_off) =
                        make_uint4(tb[0], tb[1], tb[2], tb[3]);
                    pub_r[sidx] = make_uint2(cout.xv, cout.x2);
                }
                ctx_store(c, ctxs, nslot, sidx);
            }
            __syncthreads();
        }
        if (warp == 0)
            backtrack_warp(job, arena, prow, tp_s, qrp_s, cigar_pool + job.cigar_off,
                           run_pool + job.sig_off, outs + jid, bt, lane);
        __syncthreads();
    }
}

size_t ksw_wide_scratch_words(int max_slots) { return (size_t)kCtxWords * max_slots + 4 * (size_t)max_slots; }
int ksw_wide_slots(int tlen, int w)
{
    const int n = window_blocks(tlen, w);
    return (n + kWideNT - 1) / kWideNT * kWideNT;
}

void launch_ksw_wide(int n_ctas, size_t smem, const KswJob* d_jobs, const int* d_order, int njobs,
                     const uint32_t* d_pool, uint8_t* d_arena, size_t arena_stride, uint32_t* d_scratch,
                     size_t scratch_stride_words, int max_slots, uint32_t* d_cigar, KswGapRun* d_runs,
                     KswOut* d_out, int* d_counter, cudaStream_t s)
{
    lesv_need_smem(ksw_dp_wide_kernel, smem);
    ksw_dp_wide_kernel<<<n_ctas, kWideNT, smem, s>>>(d_jobs, d_order, njobs, d_pool, d_arena, arena_stride,
                                                     d_scratch, scratch_stride_words, max_slots, d_cigar,
                                                     d_runs, d_out, d_counter);
}

size_t ksw_smem_bytes(int max_qlen, int max_tlen)
{
    return (size_t)(tp_words(max_tlen) + qrp_words(max_qlen)) * sizeof(uint32_t);
}

// <= 96 registers per thread -> 680 threads per SM
int ksw_ctas_per_sm(int nt)
{
    if (nt <= kGroupTierMax) return 20 * (32 / nt);    // groups per SM: 20 one-warp CTAs (ksw_dp_group_kernel's launch bounds)
    const int n = 680 / nt;
    return n < 1 ? 1 : n > 21 ? 21 : n;
}

using KswKernelFn = void (*)(const KswJob*, const int*, int, const uint32_t*, uint8_t*, const size_t*, uint32_t*,
                             KswGapRun*, KswOut*, int*);

cudaError_t launch_ksw(int nt, int ncta, int n_clusters, size_t smem, const KswJob* d_jobs, const int* d_order,
                       int njobs, const uint32_t* d_pool, uint8_t* d_arena, const size_t* arena_stride, uint32_t* d_cigar,
                       KswGapRun* d_runs, KswOut* d_out, int* d_counter, cudaStream_t s, bool after_resident)
{
    // after_resident: programmatic stream serialization -- this launch may start once all CTAs of the previous
    // kernel in the stream have executed griddepcontrol.launch_dependents (i.e. are resident), not when it ends.
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = after_resident ? 1 : 0;
    cudaLaunchConfig_t cfg = {};
    cfg.stream = s; cfg.attrs = at; cfg.numAttrs = 1;
    if (ncta <= 1 && nt <= kGroupTierMax) {             // several alignments per warp: n_clusters counts GROUPS (a multiple of 32 / nt)
        cfg.gridDim = dim3((unsigned)(n_clusters / (32 / nt)));
        cfg.blockDim = dim3(32u); cfg.dynamicSmemBytes = 0;
        if (nt == 8)
            return cudaLaunchKernelEx(&cfg, ksw_dp_group_kernel<8>, d_jobs, d_order, njobs, d_pool, d_arena, arena_stride, d_cigar, d_runs,
                                      d_out, d_counter);
        return cudaLaunchKernelEx(&cfg, ksw_dp_group_kernel<16>, d_jobs, d_order, njobs, d_pool, d_arena, arena_stride, d_cigar, d_runs,
                                  d_out, d_counter);
    }
    if (ncta <= 1) {
        cfg.gridDim = dim3((unsigned)n_clusters);
        if (nt >= kTierLow && 2 * smem <= 200 * 1024) {
            // DP of alignment m overlapped with the traceback walk of alignment m-1 (one extra warp)
            auto kern = nt <= kTierLow ? ksw_dp_overlap_kernel<kTierLow> : nt <= kTierMid ? ksw_dp_overlap_kernel<kTierMid>
                      : nt <= kTierTop ? ksw_dp_overlap_kernel<kTierTop> : ksw_dp_overlap_kernel<kTierXL>;
            lesv_need_smem(kern, 2 * smem);
            cfg.blockDim = dim3((unsigned)nt + 32); cfg.dynamicSmemBytes = 2 * smem;
            return cudaLaunchKernelEx(&cfg, kern, d_jobs, d_order, njobs, d_pool, d_arena, arena_stride,
                                      (uint32_t)(smem / 4), d_cigar, d_runs, d_out, d_counter);
        }
        lesv_need_smem(ksw_dp_kernel<false>, smem + (size_t)nt * 16);
        cfg.blockDim = dim3((unsigned)nt); cfg.dynamicSmemBytes = smem + (size_t)nt * 16;
        return cudaLaunchKernelEx(&cfg, ksw_dp_kernel<false>, d_jobs, d_order, njobs, d_pool, d_arena, arena_stride,
                                  d_cigar, d_runs, d_out, d_counter);
    }
    smem += (size_t)nt * 16;
    lesv_need_smem(ksw_dp_kernel<true>, smem);
    if (ncta > 8) cudaFuncSetAttribute(ksw_dp_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cfg.gridDim = dim3((unsigned)(n_clusters * ncta));
    cfg.blockDim = dim3((unsigned)nt);
    cfg.dynamicSmemBytes = smem;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = (unsigned)ncta; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, ksw_dp_kernel<true>, d_jobs, d_order, njobs, d_pool, d_arena, arena_stride,
                              d_cigar, d_runs, d_out, d_counter);
}

cudaError_t launch_ksw_strips(int n_ctas, size_t smem, const KswJob* d_jobs, const void* d_sjobs, int n_sjobs,
                              const uint32_t* d_pool, uint8_t* d_arena, uint2* d_edges, int* d_progress,
                              uint32_t* d_cigar, KswGapRun* d_runs, KswOut* d_out, int* d_counter, cudaStream_t s)
{
    lesv_need_smem(ksw_dp_strip_kernel, smem);
    ksw_dp_strip_kernel<<<n_ctas, kStripNT, smem, s>>>(d_jobs, static_cast<const StripJob*>(d_sjobs), n_sjobs, d_pool, d_arena,
                                                       d_edges, d_progress, d_cigar, d_runs, d_out, d_counter);
    return cudaGetLastError();
}

// how many clusters of this shape can be co-resident on the device (0 = shape not launchable)
int ksw_max_clusters(int nt, int ncta, size_t smem)
{
    if (ncta <= 1) return 1 << 20;
    KswKernelFn kern = ksw_dp_kernel<true>;
    smem += (size_t)nt * 16;
    lesv_need_smem(kern, smem);
    if (ncta > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ncta); cfg.blockDim = dim3((unsigned)nt); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

}  // namespace lesvksw
