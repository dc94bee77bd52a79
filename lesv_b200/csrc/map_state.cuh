// map_state.cuh -- device-side records and the per-context state of the mapper stages (map_stage.cu:
// index / seeding / chaining / windows; map_extend.cu: MEM stage and base-level alignment).
#pragma once
#include "ctx_internal.cuh"
#include "chain_core.cuh"

#include <vector>

namespace lesvmap {

constexpr uint32_t kNoSlot = 0xFFFFFFFFu;

struct IndexDev {
    const uint64_t* off;         // offset_list
    const uint64_t* dir_hash;
    const uint32_t* dir_start;   // n_dir + 1
    const uint32_t* bucket;      // 2^bb + 1
    uint32_t* qcount;            // per directory entry: query k-mers of the current batch that hit it
    uint32_t n_dir;
    int k, bucket_shift;         // bucket = hash >> bucket_shift
};

struct ReadDev {
    uint64_t off;                // residue index of the read's first base inside the uploaded bytes
    uint32_t len;
    uint32_t nk;                 // sampled k-mers per strand
    uint32_t km_base;            // index of the read's first forward k-mer; reverse ones follow at +nk
};

struct DdfHit {                  // what fill_init_hit_info leaves of a chain (chain_and_extend_kmer_matches.c:75-95)
    int score, qbeg, qend;
    int seg;
    long long sbeg, send;
};


__device__ __forceinline__ uint32_t pac_code(const uint8_t* __restrict__ pac, uint64_t l)
{
    return (pac[l >> 2] >> ((~l & 3) << 1)) & 3u;             // _get_pac, corelib/hbn_aux.h:125-126
}

}  // namespace lesvmap

struct lesvmap_ext;                                   // buffers of the extension stage (map_extend.cu)
void lesvmap_ext_free(lesvmap_ext*);

struct lesvmap_state {
    lesvmap_ext* ext = nullptr;
    int qfrom = 0;                     // first read of the batch the device buffers hold
    std::vector<lesvmap::ReadDev> h_reads;
    int k = 0, w = 0, max_occ = 0;
    std::vector<lesvgpu_seq> sseqs;
    uint64_t total_res = 0;            // end of the last contig
    int soff_bits = 1;
    uint64_t n_off = 0; uint32_t n_dir = 0; int bucket_bits = 0;
    DevBuf<uint64_t> d_seq_off, d_off, d_dir_hash;
    DevBuf<uint32_t> d_dir_start, d_bucket, d_qcount;
    // batch buffers
    DevBuf<uint8_t> d_qpac;
    DevBuf<lesvmap::ReadDev> d_reads;
    DevBuf<uint32_t> d_slot, d_mcnt, d_tmp, d_flag, d_segid, d_seg_start, d_seg_info, d_counter;
    DevBuf<uint64_t> d_k0, d_v0, d_k1, d_v1;
    DevBuf<lesvchain::Seed> d_seeds;
    DevBuf<int> d_f, d_p, d_t, d_v;
    DevBuf<lesvchain::IntPair> d_u;
    DevBuf<lesvmap::DdfHit> d_hits;
    // results of the last batch
    int sorted_in = 0; uint32_t n_match = 0, n_seg = 0, n_kmer = 0;
    std::vector<lesvgpu_subseq_hit> windows;
    std::vector<int> win_first;
    std::vector<lesvgpu_kmer_match> matches_host;
    std::vector<lesvgpu_init_hit> init_hits;
    std::vector<int> hit_first;
    lesvgpu_map_timing tm = {};
};

