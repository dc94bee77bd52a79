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
    // where the residues come from (pack kernel):
    //   src = 0: one code (0..3) per byte; seq_off_q / seq_off_t are byte offsets into the raw upload
    //   src = 1: lesv .pac (2 bits per base, base l at byte l>>2, bits (~l&3)<<1, hbn_aux.h:125-126);
    //            seq_off_q = residue index of forward-read base qoff inside the raw upload,
    //            seq_off_t = residue index of the first target base inside the resident subject .pac
    uint64_t seq_off_q;
    uint64_t seq_off_t;
    int32_t src;
    int32_t qdir;             // src = 1: 1 = align the reverse complement of the read range (REV)
    uint64_t cigar_off;       // u32 offset of this job's CIGAR slot in the cigar pool
    uint32_t cigar_cap;       // capacity of that slot (qlen + tlen + 2); 0 = do not write a CIGAR
    int32_t min_indel;        // > 0: scan the alignment for gap runs >= min_indel (qx2svsig)
    uint32_t sig_off;         // first KswGapRun slot of this job (kSigCap slots)
    int32_t nt;               // threads per CTA this alignment uses (multiple of 32, <= blockDim.x)
};

constexpr int kSigCap = 64;   // gap-run records kept per alignment

// A gap run found by the alignment scan (find_sv_signature, find_sv_signature.c:126-220), in
// coordinates of the aligned sub-sequences: DEL = gap in the read, INS = gap in the reference.
struct KswGapRun {
    int type;                 // 0 = DEL, 6 = INS (eGapAlignDel / eGapAlignIns, gapinfo.h:47-53)
    int qfrom, qto, sfrom, sto;
};

// per job results written by the kernels
struct KswOut {
    int n_cigar;              // CIGAR entries (reference order: first op first), 0 = failure
    int aln_size;             // alignment columns
    int n_match;              // identical columns (calc_ident_perc numerator)
    int status;               // 1 = alignment done
    // alignment scan (truncate_align_bad_ends, calc_effective_ident_perc, find_sv_signature)
    int trunc_ok;             // truncate_align_bad_ends' return value
    int qb, qe, sb, se;       // trimmed sub-sequence coordinates
    int eff_mat, eff_len;     // calc_effective_ident_perc = 100.0 * eff_mat / eff_len
    int trim_mat, trim_len;   // calc_ident_perc over the trimmed alignment
    int n_runs;               // gap runs >= min_indel inside the trimmed alignment (may exceed kSigCap)
};

// words of packed sequence pool one job needs
__host__ __device__ static inline uint32_t tp_words(int tlen) { return (uint32_t)(tlen >> 4) + 4; }
__host__ __device__ static inline uint32_t qrp_words(int qlen) { return (uint32_t)(qlen >> 4) + 7; }

// CTA sizes of the single-CTA launch classes.  The three largest run ksw_dp_overlap_kernel, whose CTA is the tier plus
// one walker warp: 352 / 512 / 640 threads = 11 / 16 / 20 warps, i.e. whole multiples of the four schedulers for the two
// big ones (4 and 5 warps each; 96 registers per thread fit 5 warps per scheduler, 80 would be the limit with 6).
#ifndef KSW_TIER_MID
#define KSW_TIER_MID 480
#define KSW_TIER_TOP 608
#endif
#ifndef KSW_TIER_LOW
#define KSW_TIER_LOW 320
#endif
constexpr int kTierLow = KSW_TIER_LOW, kTierMid = KSW_TIER_MID, kTierTop = KSW_TIER_TOP;
// One more single-CTA tier above them: 736 DP threads + the walker = 24 warps, six per scheduler at 80 registers (a few
// spills).  Bands of 9.7-11.8 k columns run here at about twice the rate per SM of the 2-CTA cluster that held them before.
#ifndef KSW_TIER_XL
#define KSW_TIER_XL 736
#endif
constexpr int kTierXL = KSW_TIER_XL;
// Band windows of <= 8 / <= 16 blocks (bands up to 64 / 192 columns): 4 / 2 alignments per warp (ksw_dp_group_kernel).
// For these classes the "clusters" of a launch are lane groups and `nt` is the group size.
constexpr int kGroupTierMax = 16;

// largest band window (in 16-column blocks) the register-resident kernel with NT threads holds
__host__ __device__ static inline int window_blocks(int tlen, int w)
{
    int cols = w + 1 < tlen ? w + 1 : tlen;
    return (cols + 15) / 16 + 3;
}

// Work counter of one launch plus its first-CTA-started / last-CTA-finished times (globaltimer, ns): the launch classes of
// a batch overlap inside one stream (programmatic dependent launch), where an event between two launches would serialise
// them, so the kernels stamp their own times.  Kernels receive &slot.counter.
struct KswLaunchSlot {
    int counter;
    int pad;
    unsigned long long t_start, t_end;
};
constexpr int kLaunchSlotInts = (int)(sizeof(KswLaunchSlot) / sizeof(int));

void launch_pack(const KswJob* d_jobs, int njobs, const uint8_t* d_raw, const uint8_t* d_subject_pac,
                 uint32_t* d_pool, cudaStream_t s);

// Persistent DP + backtrack kernel.  `nt` = threads per CTA (multiple of 32, <= 640), `ncta` = CTAs
// per alignment (1, or a thread-block cluster of 2/4/8/16 CTAs for bands wider than 640 blocks);
// job_order[0..njobs) lists the jobs of this launch by decreasing traceback size; cluster i starts
// with job i and writes its traceback at d_arena + d_arena_off[i]; *d_counter must be n_clusters.
// smem = ksw_smem_bytes(longest query, longest target among these jobs).
cudaError_t launch_ksw(int nt, int ncta, int n_clusters, size_t smem, const KswJob* d_jobs, const int* d_job_order,
                       int njobs, const uint32_t* d_pool, uint8_t* d_arena, const size_t* d_arena_off, uint32_t* d_cigar,
                       KswGapRun* d_runs, KswOut* d_out, int* d_counter, cudaStream_t s, bool after_resident = false);
int ksw_max_clusters(int nt, int ncta, size_t smem);

// Full-matrix alignments cut into column strips of kStripBlocks blocks (ksw_kernels.cu: ksw_dp_strip_kernel).
#ifndef KSW_STRIP_BLOCKS
#define KSW_STRIP_BLOCKS 256
#endif
constexpr int kStripBlocks = KSW_STRIP_BLOCKS;                  // threads per strip CTA
constexpr int kStripCtasPerSm = kStripBlocks <= 320 ? 2 : 1;    // 2 x 8 warps: four per scheduler, like the banded kernel
struct KswStripJob {
    int jid, s, n_strips, prog_off;
    unsigned long long arena_off, edge_off;
};
cudaError_t launch_ksw_strips(int n_ctas, size_t smem, const KswJob* d_jobs, const void* d_sjobs, int n_sjobs,
                              const uint32_t* d_pool, uint8_t* d_arena, uint2* d_edges, int* d_progress,
                              uint32_t* d_cigar, KswGapRun* d_runs, KswOut* d_out, int* d_counter, cudaStream_t s);

// Fallback for bands no cluster shape can hold: slot contexts in global scratch (ksw_wide_scratch_words per CTA).
void launch_ksw_wide(int n_ctas, size_t smem, const KswJob* d_jobs, const int* d_order, int njobs,
                     const uint32_t* d_pool, uint8_t* d_arena, size_t arena_stride, uint32_t* d_scratch,
                     size_t scratch_stride_words, int max_slots, uint32_t* d_cigar, KswGapRun* d_runs,
                     KswOut* d_out, int* d_counter, cudaStream_t s);
size_t ksw_wide_scratch_words(int max_slots);
int ksw_wide_slots(int tlen, int w);

int ksw_ctas_per_sm(int nt);
size_t ksw_smem_bytes(int max_qlen, int max_tlen);

}  // namespace lesvksw
