// ksw_kernels.cu -- K-ksw: banded dual-affine global DP with traceback on sm_100a.
//
// Replaces ksw_extd2_sse + ksw_backtrack (reference src/algo/ksw2_extd2_sse.c:26-393,
// src/algo/ksw2.h:121-153) as called by nw_ksw2_extd2 (src/algo/ksw2_wrapper.c:97-211).
//
// Mapping.  One CTA per alignment, NT threads, one 16-column block of the band per thread held
// entirely in registers (ksw_core.cuh).  Thread tid owns the block B == tid (mod NT) of the live
// window; as the band slides right along the anti-diagonals a thread drops the block that left
// on the left and picks up the one NT blocks further right.  Per anti-diagonal each thread does
// one block update (~300 integer instructions for 16 cells), hands its column-15 values to its
// right neighbour (warp shuffle; shared memory across warps) and stores 16 traceback bytes with
// one coalesced 128-bit store.  One __syncthreads per anti-diagonal.  The packed target and the
// packed reversed query (2 bits per base) are staged in shared memory once per alignment.
// CTAs are persistent and pull alignments from a queue sorted by decreasing cost; each CTA owns
// one traceback arena in HBM that is reused from alignment to alignment.  After the last
// anti-diagonal warp 0 walks the traceback (64 rows prefetched per round trip to HBM).
#include "ksw_kernels.cuh"
#include "ksw_core.cuh"
#include <stdio.h>

namespace lesvksw {

// ---------------------------------------------------------------------------------------------
// packing: byte codes -> 2-bit words
// ---------------------------------------------------------------------------------------------
__global__ void pack_kernel(const KswJob* __restrict__ jobs, const uint8_t* __restrict__ raw,
                            uint32_t* __restrict__ pool)
{
    const KswJob job = jobs[blockIdx.x];
    const uint8_t* t = raw + job.seq_off_t;
    const uint8_t* q = raw + job.seq_off_q;
    const int ntw = (int)tp_words(job.tlen), nqw = (int)qrp_words(job.qlen);
    for (int wi = threadIdx.x; wi < ntw; wi += blockDim.x) {
        uint32_t w = 0;
        const int base = wi * 16;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int idx = base + k;
            if (idx < job.tlen) w |= (uint32_t)(t[idx] & 3) << (2 * k);
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
                if (idx < job.qlen) w |= (uint32_t)(q[job.qlen - 1 - idx] & 3) << (2 * k);
            }
        }
        pool[job.qrp_off + wi] = w;
    }
}

void launch_pack(const KswJob* d_jobs, int njobs, const uint8_t* d_raw, uint32_t* d_pool, cudaStream_t s)
{
    if (njobs > 0) pack_kernel<<<njobs, 256, 0, s>>>(d_jobs, d_raw, d_pool);
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
                                        uint32_t* cigar, int& n, int lane)
{
    if (op == cur_op) { cur_len += len; return; }
    if (cur_len && lane == 0) cigar[n] = cur_len << 4 | cur_op;
    if (cur_len) ++n;
    cur_op = op; cur_len = len;
}

// All 32 lanes run the same walk out of the shared-memory row cache; lane 0 writes the CIGAR.
__device__ void backtrack_warp(const KswJob& job, const uint8_t* __restrict__ arena, size_t prow,
                               const uint32_t* tp_s, const uint32_t* qrp_s,
                               uint32_t* __restrict__ cigar, KswOut* out, BtShared& sh, int lane)
{
    const int qlen = job.qlen, tlen = job.tlen, w = job.w;
    int i = tlen - 1, j = qlen - 1, state = 0, n = 0, n_match = 0;
    uint32_t cur_op = 0xFFFFFFFFu, cur_len = 0, aln = 0;
    while (i >= 0 && j >= 0) {
        const int r0 = i + j;
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
                const uint8_t* row = arena + (size_t)r * prow;
                sh.cache[k][0] = *reinterpret_cast<const uint4*>(row + (size_t)(b0 - bst) * 16);
                if (b1 >= 0) sh.cache[k][1] = *reinterpret_cast<const uint4*>(row + (size_t)(b1 - bst) * 16);
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
            bt_emit(op, 1, cur_op, cur_len, cigar, n, lane);
            ++aln;
            if (op == 0) {                       // calc_ident_perc numerator (hbn_traceback_aux.c:4-19)
                const int qi = qlen - 1 - j;     // index into the reversed query
                const uint32_t tb = (tp_s[i >> 4] >> (2 * (i & 15))) & 3u;
                const uint32_t qb = (qrp_s[1 + (qi >> 4)] >> (2 * (qi & 15))) & 3u;
                n_match += tb == qb;
                --i; --j;
            } else if (op == 2) --i; else --j;
        }
        __syncwarp();
    }
    if (i >= 0) { bt_emit(2, (uint32_t)(i + 1), cur_op, cur_len, cigar, n, lane); aln += i + 1; }
    if (j >= 0) { bt_emit(1, (uint32_t)(j + 1), cur_op, cur_len, cigar, n, lane); aln += j + 1; }
    bt_emit(0xFFFFFFFEu, 0, cur_op, cur_len, cigar, n, lane);          // flush
    __syncwarp();
    // the walk produced last-op-first; the reference returns first-op-first (ksw2.h:148-150)
    for (int a = lane; a < n / 2; a += 32) {
        const uint32_t x = cigar[a], y = cigar[n - 1 - a];
        cigar[a] = y; cigar[n - 1 - a] = x;
    }
    if (lane == 0) { out->n_cigar = n; out->aln_size = (int)aln; out->n_match = n_match; out->status = 0; }
}

// ---------------------------------------------------------------------------------------------
// DP kernel
// ---------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT, (NT >= 512 ? 1 : 512 / NT))
ksw_dp_kernel(const KswJob* __restrict__ jobs, const int* __restrict__ order, int njobs,
              const uint32_t* __restrict__ pool, uint8_t* __restrict__ arena_base, size_t arena_stride,
              uint32_t* __restrict__ cigar_pool, KswOut* __restrict__ outs, int* counter)
{
    constexpr int NW = NT / 32;
    extern __shared__ uint32_t seq_s[];
    __shared__ uint2 carry_slot[2][NW];
    __shared__ int s_job;
    __shared__ BtShared bt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* const arena = arena_base + (size_t)blockIdx.x * arena_stride;

    for (;;) {
        if (tid == 0) s_job = atomicAdd(counter, 1);
        __syncthreads();
        const int ji = s_job;
        if (ji >= njobs) break;
        const int jid = order[ji];
        const KswJob job = jobs[jid];
        const int qlen = job.qlen, tlen = job.tlen, w = job.w;
        const int ntw = (int)tp_words(tlen), nqw = (int)qrp_words(qlen);
        uint32_t* tp_s = seq_s;
        uint32_t* qrp_s = seq_s + ntw;
        for (int x = tid; x < ntw; x += NT) tp_s[x] = pool[job.tp_off + x];
        for (int x = tid; x < nqw; x += NT) qrp_s[x] = pool[job.qrp_off + x];
        __syncthreads();

        ThreadCtx c;
        c.myB = -1; c.qwi = -0x7fffffff; c.tword = 0; c.qlo = c.qhi = 0;
        c.b.init();
        Carry prev_out = carry_init();
        const int nrow = qlen + tlen - 1;
        const size_t prow = (size_t)job.tb_blocks * 16;
        for (int r = 0; r < nrow; ++r) {
            Carry cin;
            cin.xv = __shfl_up_sync(0xFFFFFFFFu, prev_out.xv, 1);
            cin.x2 = __shfl_up_sync(0xFFFFFFFFu, prev_out.x2, 1);
            if (lane == 0) {
                const uint2 v = carry_slot[(r + 1) & 1][(warp + NW - 1) % NW];   // written in row r-1
                cin.xv = v.x; cin.x2 = v.y;
            }
            Carry cout;
            uint32_t tb[4];
            int tbb = 0;
            if (thread_row_step(c, tid, NT, r, qlen, tlen, w, tp_s, qrp_s, cin, cout, tb, tbb)) {
                *reinterpret_cast<uint4*>(arena + (size_t)r * prow + (size_t)tbb * 16) =
                    make_uint4(tb[0], tb[1], tb[2], tb[3]);
                prev_out = cout;
            }
            if (lane == 31) carry_slot[r & 1][warp] = make_uint2(prev_out.xv, prev_out.x2);
            __syncthreads();
        }
        if (warp == 0)
            backtrack_warp(job, arena, prow, tp_s, qrp_s, cigar_pool + job.cigar_off, outs + jid, bt, lane);
        __syncthreads();
    }
}

size_t ksw_smem_bytes(int max_qlen, int max_tlen)
{
    return (size_t)(tp_words(max_tlen) + qrp_words(max_qlen)) * sizeof(uint32_t);
}

int ksw_ctas_per_sm(int nt) { return nt >= 512 ? 1 : 512 / nt; }

template <int NT>
static void launch_t(int n_ctas, size_t smem, const KswJob* d_jobs, const int* d_order, int njobs,
                     const uint32_t* d_pool, uint8_t* d_arena, size_t arena_stride, uint32_t* d_cigar,
                     KswOut* d_out, int* d_counter, cudaStream_t s)
{
    cudaFuncSetAttribute(ksw_dp_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ksw_dp_kernel<NT><<<n_ctas, NT, smem, s>>>(d_jobs, d_order, njobs, d_pool, d_arena, arena_stride,
                                               d_cigar, d_out, d_counter);
}

void launch_ksw(int nt, int n_ctas, size_t smem, const KswJob* d_jobs, const int* d_order, int njobs,
                const uint32_t* d_pool, uint8_t* d_arena, size_t arena_stride, uint32_t* d_cigar,
                KswOut* d_out, int* d_counter, cudaStream_t s)
{
    switch (nt) {
        case 64:  launch_t<64>(n_ctas, smem, d_jobs, d_order, njobs, d_pool, d_arena, arena_stride, d_cigar, d_out, d_counter, s); break;
        case 128: launch_t<128>(n_ctas, smem, d_jobs, d_order, njobs, d_pool, d_arena, arena_stride, d_cigar, d_out, d_counter, s); break;
        case 256: launch_t<256>(n_ctas, smem, d_jobs, d_order, njobs, d_pool, d_arena, arena_stride, d_cigar, d_out, d_counter, s); break;
        default:  launch_t<512>(n_ctas, smem, d_jobs, d_order, njobs, d_pool, d_arena, arena_stride, d_cigar, d_out, d_counter, s); break;
    }
}

}  // namespace lesvksw
