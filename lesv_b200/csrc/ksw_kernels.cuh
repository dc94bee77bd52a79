// ksw_kernels.cuh -- device-side job descriptors and kernel entry points of K-ksw.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace lesvksw {

// One pairwise alignment (a call of nw_ksw2_extd2, reference src/algo/ksw2_wrapper.c:97-211).
struct KswJob {
    int qlen, tlen, w;        // w = band as ksw_extd2_sse receives it (already >= 0)
    int tb_blocks;            // 16-byte blocks per traceback row (n_col_, ksw2_extd2_sse.c:85-86)
    uint32_t tp_off;          // word offset of the packed target (16 bases per word) in the pool
    uint32_t qrp_off;         // word offset of the packed REVERSED query (first word is a zero pad)
    uint64_t seq_off_q;       // byte offset of the query codes in the raw upload buffer
    uint64_t seq_off_t;       // byte offset of the target codes
    uint64_t cigar_off;       // u32 offset of this job's CIGAR slot in the cigar pool
    uint32_t cigar_cap;       // capacity of that slot (qlen + tlen + 2)
    uint32_t pad_;
};

// per job results written by the kernels
struct KswOut {
    int n_cigar;              // CIGAR entries (reference order: first op first), 0 = failure
    int aln_size;             // alignment columns
    int n_match;              // identical columns (calc_ident_perc numerator)
    int status;               // 0 ok, else LESVGPU_E_*
};

// words of packed sequence pool one job needs
__host__ __device__ static inline uint32_t tp_words(int tlen) { return (uint32_t)(tlen >> 4) + 4; }
__host__ __device__ static inline uint32_t qrp_words(int qlen) { return (uint32_t)(qlen >> 4) + 7; }

// largest band window (in 16-column blocks) the register-resident kernel with NT threads holds
static inline int window_blocks(int tlen, int w)
{
    int cols = w + 1 < tlen ? w + 1 : tlen;
    return (cols + 15) / 16 + 3;
}

void launch_pack(const KswJob* d_jobs, int njobs, const uint8_t* d_raw, uint32_t* d_pool, cudaStream_t s);

// Persistent DP + backtrack kernel; `nt` in {64,128,256,512}. job_order[0..njobs) lists the jobs
// of this launch (indices into d_jobs); d_counter must be zero on entry.
// smem = ksw_smem_bytes(longest query, longest target among these jobs).
void launch_ksw(int nt, int n_ctas, size_t smem, const KswJob* d_jobs, const int* d_job_order, int njobs,
                const uint32_t* d_pool, uint8_t* d_arena, size_t arena_stride, uint32_t* d_cigar,
                KswOut* d_out, int* d_counter, cudaStream_t s);

int ksw_ctas_per_sm(int nt);
size_t ksw_smem_bytes(int max_qlen, int max_tlen);

}  // namespace lesvksw
