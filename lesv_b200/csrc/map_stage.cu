// map_stage.cu -- the candidate stage of the mapper (qx2map) on the GPU: reference k-mer index,
// query seeding, per-(read strand, contig) chaining, candidate windows.
//
// Replaces (paths relative to the reference's src/):
//   build_lookup_table                 algo/lookup_table.c:199-224 (+ :28-175, hash_list_bucket_sort.c)
//   hbn_find_kmer_matches_mt           algo/hbn_word_finder.c:337-353 (collect_kmer_matches_thread :273-335)
//   find_candidates_for_one_context    app/map/chain_and_extend_kmer_matches.c:97-186
//   chaining_find_candidates           algo/chain_dp.c:273-395 (chain_core.cuh)
//   find_candidate_subject_subseqs     app/map/hbn_find_subseq_hit.c:213-256 (host side, below)
//
// Layout in HBM.  Index: offset_list (u64 per kept k-mer, sorted by (hash, offset)), a directory of the
// distinct hashes (dir_hash u64, dir_start u32, one extra entry) and a bucket table over the top bits
// of the hash, so a lookup is one bucket read plus a scan of ~1 directory entry -- the reference
// instead merges the sorted query k-mers against the whole directory (hbn_word_finder.c:287-314).
// Batch: the query volume's .pac bytes of reads [qfrom, qto); one u32 directory slot per sampled query
// k-mer; matches as (key = context << soff_bits | soff, payload = qoff) records, emitted in (context,
// qoff, soff) order and sorted stably by key, which yields the (context, soff, qoff) order.
//
// Canonical choices where the reference's result depends on thread scheduling:
//   - the query-side occurrence filter (qc <= max_kmer_occ, :292-294) counts over the whole batch
//     [qfrom, qto), i.e. what the reference does with -num_threads 1;
//   - seeds with equal subject offset enter the chaining in ascending query offset (the reference
//     leaves that order to an unstable sort, sort_sr_hit_seeds.cpp:53-56).
#include "map_state.cuh"
#include "chain_warp.cuh"
#include "gpu_prims.cuh"

#include <algorithm>
#include <array>
#include <cstring>

using namespace lesvmap;

namespace {

constexpr int kSeedBlock = 300, kSeedSkip = 200;      // kSeedingSeqSize / kSeedingSeqStride, hbn_word_finder.c:8-9

__device__ __forceinline__ uint64_t load_kmer(const uint8_t* __restrict__ pac, uint64_t pos, int k)
{
    // bases pos .. pos+k-1 of a .pac stream (base l at byte l>>2, bits (~l&3)<<1, hbn_aux.h:125-126),
    // first base most significant -- the reference's `hash = (hash << 2) | c` loop
    const uint8_t* b = pac + (pos >> 2);
    const int sh = (int)(pos & 3) * 2;
    uint64_t hi = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) hi = (hi << 8) | b[i];
    const uint64_t v = sh ? (hi << sh) | ((uint64_t)b[8] >> (8 - sh)) : hi;
    return v >> (64 - 2 * k);
}

__device__ __forceinline__ uint64_t revcomp_kmer(uint64_t x, int k)
{
    x = ~x;                                            // complement: 3 - c
    x = __brevll(x);                                   // reverses the bits inside each base too ...
    x = ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);   // ... undo that
    return x >> (64 - 2 * k);
}

__device__ __forceinline__ uint32_t index_lookup(const IndexDev& ix, uint64_t hash)
{
    const uint32_t b = (uint32_t)(hash >> ix.bucket_shift);
    uint32_t lo = ix.bucket[b];
    const uint32_t end = ix.bucket[b + 1];
    uint32_t hi = end;
    while (hi - lo > 4) {                               // lower bound; buckets hold ~1 entry on average
        const uint32_t mid = (lo + hi) >> 1;
        if (ix.dir_hash[mid] < hash) lo = mid + 1; else hi = mid;
    }
    for (uint32_t i = lo; i < end && i <= hi; ++i) {
        const uint64_t h = ix.dir_hash[i];
        if (h == hash) return i;
        if (h > hash) break;
    }
    return kNoSlot;
}

// ---------------------------------------------------------------------------------------------
// index build
// ---------------------------------------------------------------------------------------------
__global__ void index_extract_kernel(const uint8_t* __restrict__ pac, const uint64_t* __restrict__ seq_off,
                                     const uint64_t* __restrict__ km_pre, int nseq, int k, int w, uint64_t nk,
                                     uint64_t* __restrict__ keys, uint64_t* __restrict__ vals)
{
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nk) return;
    int lo = 0, hi = nseq;                              // contig c with km_pre[c] <= g < km_pre[c+1]
    while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (km_pre[m] <= g) lo = m; else hi = m; }
    const uint64_t pos = seq_off[lo] + (g - km_pre[lo]) * (uint64_t)w;
    keys[g] = load_kmer(pac, pos, k);
    vals[g] = pos;
}

__global__ void run_flag_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ flag)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
// rid[i] = exclusive scan of flag; for heads rid is the run's id, run_start[id] = i
__global__ void run_start_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rid, uint64_t n,
                                 uint32_t* __restrict__ run_start)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) run_start[rid[i]] = (uint32_t)i;
}
__global__ void run_keep_kernel(const uint32_t* __restrict__ run_start, uint32_t n_runs, uint32_t n, uint32_t max_occ,
                                uint32_t* __restrict__ kept_len, uint32_t* __restrict__ kept_flag)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const uint32_t end = r + 1 < n_runs ? run_start[r + 1] : n;
    const uint32_t len = end - run_start[r];
    const bool keep = len <= max_occ;                  // remove_repetitive_kmers, lookup_table.c:95-138
    kept_len[r] = keep ? len : 0u;
    kept_flag[r] = keep ? 1u : 0u;
}
__global__ void index_compact_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals, uint64_t n,
                                     const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rid,
                                     const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ kept_flag,
                                     const uint32_t* __restrict__ new_start, const uint32_t* __restrict__ new_idx,
                                     uint64_t* __restrict__ off, uint64_t* __restrict__ dir_hash, uint32_t* __restrict__ dir_start)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = flag[i] ? rid[i] : rid[i] - 1;  // exclusive scan: non-heads see the count of heads so far
    if (!kept_flag[r]) return;
    off[new_start[r] + ((uint32_t)i - run_start[r])] = vals[i];
    if (flag[i]) { dir_hash[new_idx[r]] = keys[i]; dir_start[new_idx[r]] = new_start[r]; }
}
__global__ void bucket_count_kernel(const uint64_t* __restrict__ dir_hash, uint32_t n_dir, int shift, uint32_t* __restrict__ cnt)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_dir) atomicAdd(&cnt[(uint32_t)(dir_hash[i] >> shift)], 1u);
}

// ---------------------------------------------------------------------------------------------
// seeding
// ---------------------------------------------------------------------------------------------
// One CTA per (read, strand).  Sampling as collect_ddfkmer_subseq (hbn_word_finder.c:187-218): blocks of
// 300 bases, 200 skipped in between, every position inside a block that still holds a full k-mer.
template <int kPhase>      // 0: look every k-mer up, count per directory entry; 1: emit the matches
__global__ void __launch_bounds__(256)
seed_kernel(const uint8_t* __restrict__ qpac, const ReadDev* __restrict__ reads, IndexDev ix, int smem_cap,
            uint32_t* __restrict__ slot, const uint32_t* __restrict__ mpos, uint32_t max_occ, int soff_bits,
            uint64_t* __restrict__ mkey, uint64_t* __restrict__ mval)
{
    extern __shared__ uint8_t seq_s[];
    const uint32_t r = blockIdx.x >> 1, strand = blockIdx.x & 1;
    const ReadDev rd = reads[r];
    const int k = ix.k, per_block = kSeedBlock - k + 1;
    const uint32_t g0 = rd.km_base + strand * rd.nk;
    if (kPhase == 0) {
        const uint8_t* src = qpac + (rd.off >> 2);
        const uint32_t nbytes = (uint32_t)(((rd.off & 3) + rd.len + 3) >> 2) + 9;
        const uint8_t* seq = src;
        if ((int)nbytes <= smem_cap) {
            for (uint32_t x = threadIdx.x; x < nbytes; x += blockDim.x) seq_s[x] = src[x];
            __syncthreads();
            seq = seq_s;
        }
        const uint64_t base = rd.off & 3;
        for (uint32_t t = threadIdx.x; t < rd.nk; t += blockDim.x) {
            const uint32_t b = t / per_block, p = b * (kSeedBlock + kSeedSkip) + (t - b * per_block);
            uint64_t hash;
            if (!strand) hash = load_kmer(seq, base + p, k);
            else hash = revcomp_kmer(load_kmer(seq, base + (rd.len - p - k), k), k);
            const uint32_t si = index_lookup(ix, hash);
            slot[g0 + t] = si;
            if (si != kNoSlot) atomicAdd(&ix.qcount[si], 1u);
        }
    } else {
        const uint64_t ctx = (uint64_t)(2 * r + strand) << soff_bits;
        for (uint32_t t = threadIdx.x; t < rd.nk; t += blockDim.x) {
            const uint32_t si = slot[g0 + t];
            if (si == kNoSlot || ix.qcount[si] > max_occ) continue;
            const uint32_t b = t / per_block, p = b * (kSeedBlock + kSeedSkip) + (t - b * per_block);
            const uint32_t s0 = ix.dir_start[si], s1 = ix.dir_start[si + 1];
            uint32_t o = mpos[g0 + t];
            for (uint32_t x = s0; x < s1; ++x, ++o) { mkey[o] = ctx | ix.off[x]; mval[o] = p; }
        }
    }
}

__global__ void match_count_kernel(const uint32_t* __restrict__ slot, uint32_t n, IndexDev ix, uint32_t max_occ,
                                   uint32_t* __restrict__ mcnt)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const uint32_t si = slot[g];
    mcnt[g] = (si != kNoSlot && ix.qcount[si] <= max_occ) ? ix.dir_start[si + 1] - ix.dir_start[si] : 0u;
}
__global__ void qcount_reset_kernel(const uint32_t* __restrict__ slot, uint32_t n, uint32_t* __restrict__ qcount)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n && slot[g] != kNoSlot) qcount[slot[g]] = 0u;
}

// ---------------------------------------------------------------------------------------------
// segments = (context, contig) runs of the sorted matches; chaining
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int contig_of(const uint64_t* __restrict__ seq_off, int nseq, uint64_t soff)
{
    int lo = 0, hi = nseq;                              // seqdb_offset_to_seq_id, corelib/seqdb.c:223-237
    while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (seq_off[m] <= soff) lo = m; else hi = m; }
    return lo;
}
__global__ void seg_flag_kernel(const uint64_t* __restrict__ mkey, uint32_t n, int soff_bits,
                                const uint64_t* __restrict__ seq_off, int nseq, uint32_t* __restrict__ flag)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t f = 1;
    if (i) {
        const uint64_t a = mkey[i - 1], b = mkey[i], m = (1ull << soff_bits) - 1;
        f = (a >> soff_bits) != (b >> soff_bits) || contig_of(seq_off, nseq, a & m) != contig_of(seq_off, nseq, b & m);
    }
    flag[i] = f;
}

struct ChainScratch {
    lesvchain::Seed* seeds;
    int *f, *p, *t, *v;
    lesvchain::IntPair* u;
};

// One warp per (read strand, contig) segment: seed list on lane 0, chaining DP on the warp (chain_warp.cuh),
// backtracking and joining on lane 0.
__global__ void __launch_bounds__(128)
chain_kernel(const uint64_t* __restrict__ mkey, const uint64_t* __restrict__ mval, const uint32_t* __restrict__ seg_start,
             uint32_t n_seg, uint32_t n_match, int soff_bits, const uint64_t* __restrict__ seq_off, int nseq, int k,
             int min_cnt, int min_score, ChainScratch sc, DdfHit* __restrict__ hits, uint32_t hit_cap,
             uint32_t* __restrict__ hit_counter, uint32_t* __restrict__ seg_ctx_sid)
{
    const uint32_t sg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (sg >= n_seg) return;
    const uint32_t a = seg_start[sg], b = sg + 1 < n_seg ? seg_start[sg + 1] : n_match;
    const uint64_t m = (1ull << soff_bits) - 1;
    const uint32_t ctx = (uint32_t)(mkey[a] >> soff_bits);
    const int sid = contig_of(seq_off, nseq, mkey[a] & m);
    if (lane == 0) { seg_ctx_sid[2 * sg] = ctx; seg_ctx_sid[2 * sg + 1] = (uint32_t)sid; }
    if ((int)(b - a) < min_cnt) return;                                  // chain_and_extend_kmer_matches.c:167
    const uint64_t sbeg = seq_off[sid];
    lesvchain::Seed* seeds = sc.seeds + a;
    int n = 0;
    if (lane == 0) {
        for (uint32_t i = a; i < b;) {                                   // :105-125: drop offsets hit by > 4 query positions
            uint32_t j = i + 1;
            while (j < b && mkey[j] == mkey[i]) ++j;
            if (j - i <= 4)
                for (uint32_t x = i; x < j; ++x) {
                    seeds[n].q = (int)mval[x]; seeds[n].s = (int)((mkey[x] & m) - sbeg); seeds[n].len = k; ++n;
                }
            i = j;
        }
    }
    n = __shfl_sync(0xFFFFFFFFu, n, 0);
    if (n == 0) return;
    const lesvchain::Params P = lesvchain::make_params(min_cnt, min_score);
    lesvchain::score_seeds_warp(P, seeds, n, false, sc.f + a, sc.p + a, sc.t + a, sc.v + a);
    if (lane != 0) return;
    lesvchain::Hit pre[lesvchain::kMaxChains], out[lesvchain::kMaxJoined];
    int members[lesvchain::kMaxChains], n_pre = 0;
    const int nh = lesvchain::find_candidates_scored(P, seeds, n, sc.f + a, sc.p + a, sc.t + a, sc.v + a, sc.u + a,
                                                     pre, &n_pre, out, members);
    if (nh == 0) return;
    const uint32_t o = atomicAdd(hit_counter, (uint32_t)nh);
    if (o + nh > hit_cap) return;                                        // host checks the counter and retries
    for (int h = 0; h < nh; ++h) {
        DdfHit d;
        d.score = out[h].score; d.qbeg = out[h].qbeg; d.qend = out[h].qend; d.sbeg = out[h].sbeg; d.send = out[h].send;
        d.seg = (int)sg | (h << 26);                                     // order inside the segment: h (< 20)
        hits[o + h] = d;
    }
}

template <class T> cudaError_t dmalloc(T** p, size_t n) { return cudaMalloc((void**)p, (n ? n : 1) * sizeof(T)); }

}  // namespace

// ---------------------------------------------------------------------------------------------
// state
// ---------------------------------------------------------------------------------------------
void lesvmap_state_free(lesvmap_state* m)
{
    if (!m) return;
    if (m->ext) lesvmap_ext_free(m->ext);
    m->d_seq_off.release(); m->d_off.release(); m->d_dir_hash.release(); m->d_dir_start.release();
    m->d_bucket.release(); m->d_qcount.release(); m->d_qpac.release(); m->d_reads.release(); m->d_slot.release();
    m->d_mcnt.release(); m->d_tmp.release(); m->d_flag.release(); m->d_segid.release(); m->d_seg_start.release();
    m->d_seg_info.release(); m->d_counter.release(); m->d_k0.release(); m->d_v0.release(); m->d_k1.release();
    m->d_v1.release(); m->d_seeds.release(); m->d_f.release(); m->d_p.release(); m->d_t.release(); m->d_v.release();
    m->d_u.release(); m->d_hits.release();
    delete m;
}

static inline unsigned nblocks(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }
static inline int bit_length(uint64_t x) { int b = 0; while (x) { ++b; x >>= 1; } return b; }

// ---------------------------------------------------------------------------------------------
// C ABI: index
// ---------------------------------------------------------------------------------------------
extern "C" int lesvgpu_index_build(lesvgpu_ctx* ctx, const lesvgpu_seq* seqs, int nseq, int kmer_size, int kmer_window,
                                   int max_kmer_occ)
{
    if (!ctx || !seqs || nseq <= 0 || kmer_size < 4 || kmer_size > 31 || kmer_window < 1 || max_kmer_occ < 1) return LESVGPU_E_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return LESVGPU_E_CUDA;
    if (!ctx->subject_bytes) { ctx->err = "lesvgpu_subject_load has not been called"; return LESVGPU_E_ARG; }
    if (ctx->map) { lesvmap_state_free(ctx->map); ctx->map = nullptr; }
    lesvmap_state* M = new lesvmap_state;
    ctx->map = M;
    M->k = kmer_size; M->w = kmer_window; M->max_occ = max_kmer_occ;
    M->sseqs.assign(seqs, seqs + nseq);
    std::vector<uint64_t> off(nseq), pre(nseq + 1, 0);
    for (int i = 0; i < nseq; ++i) {
        off[i] = seqs[i].seq_offset;
        if (i && off[i] < off[i - 1]) { ctx->err = "contig offsets must ascend"; return LESVGPU_E_ARG; }
        if (((seqs[i].seq_offset + seqs[i].seq_size + 3) >> 2) > ctx->subject_bytes) { ctx->err = "contig outside the loaded .pac"; return LESVGPU_E_ARG; }
        const uint64_t n = seqs[i].seq_size < (uint64_t)kmer_size ? 0 : (seqs[i].seq_size - kmer_size) / kmer_window + 1;   // calc_num_kmers :9-26
        pre[i + 1] = pre[i] + n;
    }
    M->total_res = seqs[nseq - 1].seq_offset + seqs[nseq - 1].seq_size;
    M->soff_bits = std::max(1, bit_length(M->total_res));
    const uint64_t nk = pre[nseq];
    if (nk >= 0xFFFFFFF0ull) { ctx->err = "more than 2^32 reference k-mers in one volume"; return LESVGPU_E_ARG; }
    cudaStream_t s = ctx->stream;
    CK(M->d_seq_off.reserve(nseq));
    CK(cudaMemcpyAsync(M->d_seq_off.p, off.data(), sizeof(uint64_t) * nseq, cudaMemcpyHostToDevice, s));
    uint64_t *k0 = nullptr, *v0 = nullptr, *k1 = nullptr, *v1 = nullptr, *d_pre = nullptr;
    uint32_t *tmp = nullptr, *flag = nullptr, *rid = nullptr, *run_start = nullptr, *kept_len = nullptr, *kept_flag = nullptr, *d_tot = nullptr;
    auto cleanup = [&]() {
        cudaFree(k0); cudaFree(v0); cudaFree(k1); cudaFree(v1); cudaFree(d_pre); cudaFree(tmp); cudaFree(flag); cudaFree(rid);
        cudaFree(run_start); cudaFree(kept_len); cudaFree(kept_flag); cudaFree(d_tot);
    };
#define CKI(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); char b__[256]; snprintf(b__, sizeof b__, "%s:%d %s", __FILE__, __LINE__, cudaGetErrorString(e__)); ctx->err = b__; return LESVGPU_E_CUDA; } } while (0)
    uint32_t n_runs = 0, tot[2] = {0, 0};
    if (nk) {
        CKI(dmalloc(&k0, nk)); CKI(dmalloc(&v0, nk)); CKI(dmalloc(&k1, nk)); CKI(dmalloc(&v1, nk));
        CKI(dmalloc(&d_pre, (size_t)nseq + 1));
        CKI(dmalloc(&tmp, std::max(radix_tmp_words(nk), scan_tmp_words(nk))));
        CKI(dmalloc(&flag, nk)); CKI(dmalloc(&rid, nk)); CKI(dmalloc(&d_tot, 4));
        CKI(cudaMemcpyAsync(d_pre, pre.data(), sizeof(uint64_t) * (nseq + 1), cudaMemcpyHostToDevice, s));
        index_extract_kernel<<<nblocks(nk, 256), 256, 0, s>>>(ctx->d_subject.p, M->d_seq_off.p, d_pre, nseq, kmer_size, kmer_window, nk, k0, v0); ++g_lesv_launches;
        const int where = radix_sort_pairs(k0, v0, k1, v1, nk, 0, 2 * kmer_size, tmp, s);
        uint64_t* K = where ? k1 : k0; uint64_t* V = where ? v1 : v0;
        run_flag_kernel<<<nblocks(nk, 256), 256, 0, s>>>(K, nk, flag); ++g_lesv_launches;
        exclusive_scan_u32(flag, rid, nk, d_tot, tmp, s);
        CKI(cudaMemcpyAsync(&n_runs, d_tot, 4, cudaMemcpyDeviceToHost, s));
        CKI(cudaStreamSynchronize(s));
        CKI(dmalloc(&run_start, (size_t)n_runs + 1)); CKI(dmalloc(&kept_len, (size_t)n_runs + 1)); CKI(dmalloc(&kept_flag, (size_t)n_runs + 1));
        run_start_kernel<<<nblocks(nk, 256), 256, 0, s>>>(flag, rid, nk, run_start); ++g_lesv_launches;
        run_keep_kernel<<<nblocks(n_runs, 256), 256, 0, s>>>(run_start, n_runs, (uint32_t)nk, (uint32_t)max_kmer_occ, kept_len, kept_flag); ++g_lesv_launches;
        // exclusive scans in place: kept_len -> new_start, kept_flag -> new_idx (kept_flag is needed afterwards: copy first)
        uint32_t* keep_copy = nullptr;
        CKI(dmalloc(&keep_copy, (size_t)n_runs));
        cudaMemcpyAsync(keep_copy, kept_flag, sizeof(uint32_t) * n_runs, cudaMemcpyDeviceToDevice, s);
        exclusive_scan_u32(kept_len, kept_len, n_runs, d_tot, tmp, s);
        exclusive_scan_u32(kept_flag, kept_flag, n_runs, d_tot + 1, tmp, s);
        cudaError_t e1 = cudaMemcpyAsync(tot, d_tot, 8, cudaMemcpyDeviceToHost, s);
        cudaError_t e2 = cudaStreamSynchronize(s);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaFree(keep_copy); CKI(e1 != cudaSuccess ? e1 : e2); }
        M->n_off = tot[0]; M->n_dir = tot[1];
        cudaError_t e3 = M->d_off.reserve(M->n_off + 1);
        if (e3 == cudaSuccess) e3 = M->d_dir_hash.reserve((size_t)M->n_dir + 1);
        if (e3 == cudaSuccess) e3 = M->d_dir_start.reserve((size_t)M->n_dir + 2);
        if (e3 != cudaSuccess) { cudaFree(keep_copy); CKI(e3); }
        index_compact_kernel<<<nblocks(nk, 256), 256, 0, s>>>(K, V, nk, flag, rid, run_start, keep_copy, kept_len, kept_flag,
                                                              M->d_off.p, M->d_dir_hash.p, M->d_dir_start.p); ++g_lesv_launches;
        const uint32_t end = (uint32_t)M->n_off;
        cudaMemcpyAsync(M->d_dir_start.p + M->n_dir, &end, 4, cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);
        cudaFree(keep_copy);
    } else {
        CKI(M->d_off.reserve(1)); CKI(M->d_dir_hash.reserve(1)); CKI(M->d_dir_start.reserve(2));
        CKI(cudaMemsetAsync(M->d_dir_start.p, 0, 8, s));
    }
    // bucket table over the top bits of the hash
    int bb = bit_length(M->n_dir);
    bb = std::max(8, std::min(bb, std::min(2 * kmer_size, 28)));
    M->bucket_bits = bb;
    const size_t nb = ((size_t)1 << bb) + 1;
    CKI(M->d_bucket.reserve(nb));
    CKI(M->d_qcount.reserve((size_t)M->n_dir + 1));
    {
        uint32_t* cnt = nullptr; uint32_t* stmp = nullptr;
        CKI(dmalloc(&cnt, nb));
        cudaError_t e = dmalloc(&stmp, scan_tmp_words(nb));
        if (e != cudaSuccess) { cudaFree(cnt); CKI(e); }
        cudaMemsetAsync(cnt, 0, sizeof(uint32_t) * nb, s);
        if (M->n_dir) { bucket_count_kernel<<<nblocks(M->n_dir, 256), 256, 0, s>>>(M->d_dir_hash.p, M->n_dir, 2 * kmer_size - bb, cnt); ++g_lesv_launches; }
        exclusive_scan_u32(cnt, M->d_bucket.p, nb, nullptr, stmp, s);
        cudaMemsetAsync(M->d_qcount.p, 0, sizeof(uint32_t) * ((size_t)M->n_dir + 1), s);
        e = cudaStreamSynchronize(s);
        cudaFree(cnt); cudaFree(stmp);
        CKI(e);
    }
    CKI(cudaGetLastError());
    cleanup();
#undef CKI
    return LESVGPU_OK;
}

extern "C" int lesvgpu_index_stats(const lesvgpu_ctx* ctx, uint64_t* n_offsets, uint64_t* n_distinct)
{
    if (!ctx || !ctx->map) return LESVGPU_E_ARG;
    if (n_offsets) *n_offsets = ctx->map->n_off;
    if (n_distinct) *n_distinct = ctx->map->n_dir;
    return LESVGPU_OK;
}

extern "C" int lesvgpu_index_export(lesvgpu_ctx* ctx, uint64_t* offset_list, uint64_t* hashes, uint32_t* starts)
{
    if (!ctx || !ctx->map) return LESVGPU_E_ARG;
    lesvmap_state* M = ctx->map;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return LESVGPU_E_CUDA;
    if (offset_list && M->n_off) CK(cudaMemcpy(offset_list, M->d_off.p, sizeof(uint64_t) * M->n_off, cudaMemcpyDeviceToHost));
    if (hashes && M->n_dir) CK(cudaMemcpy(hashes, M->d_dir_hash.p, sizeof(uint64_t) * M->n_dir, cudaMemcpyDeviceToHost));
    if (starts) CK(cudaMemcpy(starts, M->d_dir_start.p, sizeof(uint32_t) * ((size_t)M->n_dir + 1), cudaMemcpyDeviceToHost));
    return LESVGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// host side of the candidate stage: find_candidate_subject_subseqs (app/map/hbn_find_subseq_hit.c)
// ---------------------------------------------------------------------------------------------
namespace {

struct HostHit { int qdir, qoff, sid, score; long long soff; };

// adjust_init_hit_subject_offset, hbn_find_subseq_hit.c:120-156 (task rm)
void window_of(const HostHit& h, int query_size, size_t subject_size, size_t* sfrom, size_t* sto)
{
    if (subject_size <= (size_t)query_size || subject_size - query_size <= 2000) { *sfrom = 0; *sto = subject_size; return; }
    const int kMaxGap = 30000;
    const size_t ql = (size_t)h.qoff, sl = (size_t)h.soff;
    if (sl <= ql) *sfrom = 0;
    else {
        size_t x = (size_t)(query_size * 1.3);
        if (x > ql + kMaxGap) x = ql + kMaxGap;
        if (x > sl) x = sl;
        *sfrom = (size_t)h.soff - x;
    }
    const size_t qr = (size_t)(query_size - h.qoff), sr = subject_size - (size_t)h.soff;
    if (sr <= qr) *sto = subject_size;
    else {
        size_t x = (size_t)(query_size * 1.3);
        if (x > qr + kMaxGap) x = qr + kMaxGap;
        if (x > sr) x = sr;
        *sto = (size_t)h.soff + x;
    }
}

struct WinSfromLt {       // subseq_hit_sfrom_lt, app/map/hbn_subseq_hit.c:8
    bool operator()(const lesvgpu_subseq_hit& a, const lesvgpu_subseq_hit& b) const
    { return a.sfrom < b.sfrom || (a.sfrom == b.sfrom && a.sto > b.sto); }
};

// merge_adjacent_subject_subseqs, hbn_find_subseq_hit.c:83-117 (stale fields of the slot kept, App. A #7)
void merge_windows(std::vector<lesvgpu_subseq_hit>& v, size_t max_merge_dist)
{
    const int n = (int)v.size();
    if (n < 2) return;
    lesvsort::introsort(v.data(), (size_t)n, WinSfromLt());
    int i = 0, idx = 0;
    while (i < n) {
        if (v[i].score == 0) { ++i; continue; }
        size_t last_send = v[i].sto;
        int j = i + 1, score = v[i].score;
        while (j < n) {
            if (v[j].score == 0) continue;        // (never true here: scores of unvisited entries are > 0)
            if (v[j].sfrom <= last_send || v[j].sfrom - last_send <= max_merge_dist) {
                last_send = std::max(last_send, (size_t)v[j].sto);
                score = std::max(score, v[j].score);
                v[j].score = 0;
                ++j;
            } else break;
        }
        v[idx].sfrom = v[i].sfrom; v[idx].sto = last_send; v[idx].score = score;
        ++idx;
        i = j;
    }
    v.resize(idx);
}

struct InitSidLt { bool operator()(const lesvgpu_init_hit& a, const lesvgpu_init_hit& b) const { return a.sid < b.sid; } };
struct SbjScoreGt { bool operator()(const std::array<int, 3>& a, const std::array<int, 3>& b) const { return a[2] > b[2]; } };

}  // namespace

static void candidate_windows(const lesvmap_state* M, std::vector<lesvgpu_init_hit>& hits, int qid, int qsize, int hitlist_size,
                              std::vector<lesvgpu_subseq_hit>& out)
{
    if (hits.empty()) return;
    // group_subject_init_hits_map_task, hbn_find_subseq_hit.c:50-81
    lesvsort::introsort(hits.data(), hits.size(), InitSidLt());
    std::vector<std::array<int, 3>> info;               // offset, count, best score
    for (size_t i = 0; i < hits.size();) {
        size_t j = i + 1;
        while (j < hits.size() && hits[i].sid == hits[j].sid) ++j;
        int best = 0;
        for (size_t x = i; x < j; ++x) best = std::max(best, hits[x].score);
        info.push_back({(int)i, (int)(j - i), best});
        i = j;
    }
    lesvsort::introsort(info.data(), info.size(), SbjScoreGt());
    for (int g = 0; g < hitlist_size && g < (int)info.size(); ++g) {
        const lesvgpu_init_hit* h0 = hits.data() + info[g][0];
        const int cnt = info[g][1];
        const size_t ssize = M->sseqs[h0->sid].seq_size;
        std::vector<lesvgpu_subseq_hit> fwd, rev;     // setup_candidates_for_one_subject :159-211
        for (int x = 0; x < cnt; ++x) {
            const lesvgpu_init_hit& ih = h0[x];
            HostHit h{ih.qdir, ih.qoff, ih.sid, ih.score, (long long)ih.soff};
            size_t sfrom = 0, sto = 0;
            window_of(h, qsize, ssize, &sfrom, &sto);
            lesvgpu_subseq_hit wv;
            wv.qid = qid; wv.qdir = ih.qdir; wv.qoff = ih.qoff; wv.sid = ih.sid;
            wv.soff = (int)((size_t)ih.soff - sfrom);
            wv.sfrom = sfrom; wv.sto = sto; wv.score = ih.score;
            (ih.qdir == 0 ? fwd : rev).push_back(wv);
        }
        merge_windows(fwd, 500);
        merge_windows(rev, 500);
        out.insert(out.end(), fwd.begin(), fwd.end());
        out.insert(out.end(), rev.begin(), rev.end());
    }
}

// ---------------------------------------------------------------------------------------------
// C ABI: candidate stage for one query batch
// ---------------------------------------------------------------------------------------------
static uint32_t sampled_kmers(uint32_t len, int k)
{
    uint32_t cnt = 0;
    for (uint32_t s = 0; s < len;) {                    // collect_ddfkmer_subseq, hbn_word_finder.c:187-218
        const uint32_t e = std::min(s + (uint32_t)kSeedBlock, len);
        if (e - s >= (uint32_t)k) cnt += e - s - k + 1;
        s = e + kSeedSkip;
    }
    return cnt;
}

extern "C" int lesvgpu_map_candidates(lesvgpu_ctx* ctx, const uint8_t* query_pac, const lesvgpu_seq* qseqs,
                                      int qfrom, int qto, const lesvgpu_map_opts* opts,
                                      const lesvgpu_subseq_hit** windows, const int** first)
{
    if (!ctx || !ctx->map || !query_pac || !qseqs || qto < qfrom || !opts) return LESVGPU_E_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return LESVGPU_E_CUDA;
    lesvmap_state* M = ctx->map;
    cudaStream_t s = ctx->stream;
    const int nr = qto - qfrom, k = M->k;
    M->windows.clear(); M->win_first.assign(nr + 1, 0); M->matches_host.clear();
    M->init_hits.clear(); M->hit_first.assign(nr + 1, 0);
    M->tm = lesvgpu_map_timing{};
    M->n_match = 0; M->n_seg = 0; M->n_kmer = 0;
    if (windows) *windows = nullptr;
    if (first) *first = M->win_first.data();
    if (nr == 0) return LESVGPU_OK;
    const int min_cnt = opts->min_ddfs, min_score = (int)(opts->min_ddfs * k * 0.8);   // chain_and_extend_kmer_matches.c:59
    cudaEvent_t* ev = ctx->ev;

    // ---- reads of the batch -------------------------------------------------------------------
    const uint64_t byte0 = qseqs[qfrom].seq_offset >> 2;
    const uint64_t byte1 = (qseqs[qto - 1].seq_offset + qseqs[qto - 1].seq_size + 3) >> 2;
    const size_t nbytes = (size_t)(byte1 - byte0);
    std::vector<ReadDev>& reads = M->h_reads;
    reads.assign(nr, ReadDev{});
    M->qfrom = qfrom;
    uint64_t nkm = 0; uint32_t max_read_bytes = 0;
    for (int i = 0; i < nr; ++i) {
        const lesvgpu_seq& q = qseqs[qfrom + i];
        if (q.seq_size >= 0x7FFFFFFFull) { ctx->err = "read longer than 2^31"; return LESVGPU_E_ARG; }
        reads[i].off = q.seq_offset - 4 * byte0;
        reads[i].len = (uint32_t)q.seq_size;
        reads[i].nk = sampled_kmers(reads[i].len, k);
        reads[i].km_base = (uint32_t)nkm;
        nkm += 2ull * reads[i].nk;
        max_read_bytes = std::max(max_read_bytes, (uint32_t)((reads[i].len + 6) >> 2) + 9);
    }
    if (nkm >= 0xFFFFFFF0ull) { ctx->err = "batch too large: more than 2^32 sampled k-mers (lower -query_batch_size)"; return LESVGPU_E_ARG; }
    if (2ull * nr >= (1ull << (63 - M->soff_bits))) { ctx->err = "batch too large for the match key"; return LESVGPU_E_ARG; }
    M->n_kmer = (uint32_t)nkm;
    CK(M->d_qpac.reserve(nbytes + 32));
    CK(M->d_reads.reserve(nr));
    CK(M->d_slot.reserve(nkm)); CK(M->d_mcnt.reserve(nkm + 1));
    CK(M->d_tmp.reserve(scan_tmp_words(nkm + 1)));
    CK(M->d_counter.reserve(8));
    CK(cudaEventRecord(ev[0], s));
    CK(cudaMemcpyAsync(M->d_qpac.p, query_pac + byte0, nbytes, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(M->d_qpac.p + nbytes, 0, 32, s));
    CK(cudaMemcpyAsync(M->d_reads.p, reads.data(), sizeof(ReadDev) * nr, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(ev[1], s));

    IndexDev ix;
    ix.off = M->d_off.p; ix.dir_hash = M->d_dir_hash.p; ix.dir_start = M->d_dir_start.p; ix.bucket = M->d_bucket.p;
    ix.qcount = M->d_qcount.p; ix.n_dir = M->n_dir; ix.k = k; ix.bucket_shift = 2 * k - M->bucket_bits;
    int smem_cap = (int)std::min<uint32_t>(max_read_bytes, 160 * 1024);
    CK(lesv_need_smem(seed_kernel<0>, (size_t)smem_cap));
    uint32_t n_match = 0;
    if (nkm) {
        seed_kernel<0><<<2 * nr, 256, smem_cap, s>>>(M->d_qpac.p, M->d_reads.p, ix, smem_cap, M->d_slot.p, nullptr,
                                                    (uint32_t)M->max_occ, M->soff_bits, nullptr, nullptr); ++g_lesv_launches;
        match_count_kernel<<<nblocks(nkm, 256), 256, 0, s>>>(M->d_slot.p, (uint32_t)nkm, ix, (uint32_t)M->max_occ, M->d_mcnt.p); ++g_lesv_launches;
        exclusive_scan_u32(M->d_mcnt.p, M->d_mcnt.p, nkm, M->d_counter.p, M->d_tmp.p, s);
        CK(cudaMemcpyAsync(&n_match, M->d_counter.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    M->n_match = n_match;
    CK(cudaEventRecord(ev[2], s));
    uint32_t n_seg = 0, n_hits = 0;
    std::vector<DdfHit> hhits;
    std::vector<uint32_t> seg_info;
    const int ctx_bits = std::max(1, bit_length((uint64_t)2 * nr - 1));
    if (n_match) {
        CK(M->d_k0.reserve(n_match)); CK(M->d_v0.reserve(n_match)); CK(M->d_k1.reserve(n_match)); CK(M->d_v1.reserve(n_match));
        CK(M->d_tmp.reserve(std::max(radix_tmp_words(n_match), scan_tmp_words((size_t)n_match + 1))));
        seed_kernel<1><<<2 * nr, 256, 0, s>>>(M->d_qpac.p, M->d_reads.p, ix, 0, M->d_slot.p, M->d_mcnt.p,
                                              (uint32_t)M->max_occ, M->soff_bits, M->d_k0.p, M->d_v0.p); ++g_lesv_launches;
        qcount_reset_kernel<<<nblocks(nkm, 256), 256, 0, s>>>(M->d_slot.p, (uint32_t)nkm, M->d_qcount.p); ++g_lesv_launches;
        CK(cudaEventRecord(ev[3], s));
        M->sorted_in = radix_sort_pairs(M->d_k0.p, M->d_v0.p, M->d_k1.p, M->d_v1.p, n_match, 0, M->soff_bits + ctx_bits, M->d_tmp.p, s);
        CK(cudaEventRecord(ev[4], s));
        const uint64_t* K = M->sorted_in ? M->d_k1.p : M->d_k0.p;
        const uint64_t* V = M->sorted_in ? M->d_v1.p : M->d_v0.p;
        // segments
        CK(M->d_flag.reserve(n_match)); CK(M->d_segid.reserve(n_match));
        seg_flag_kernel<<<nblocks(n_match, 256), 256, 0, s>>>(K, n_match, M->soff_bits, M->d_seq_off.p, (int)M->sseqs.size(), M->d_flag.p); ++g_lesv_launches;
        exclusive_scan_u32(M->d_flag.p, M->d_segid.p, n_match, M->d_counter.p, M->d_tmp.p, s);
        CK(cudaMemcpyAsync(&n_seg, M->d_counter.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        M->n_seg = n_seg;
        if (n_seg > 0x3FFFFFF) { ctx->err = "more than 2^26 segments in one batch"; return LESVGPU_E_ARG; }
        CK(M->d_seg_start.reserve((size_t)n_seg + 1)); CK(M->d_seg_info.reserve(2 * (size_t)n_seg));
        run_start_kernel<<<nblocks(n_match, 256), 256, 0, s>>>(M->d_flag.p, M->d_segid.p, n_match, M->d_seg_start.p); ++g_lesv_launches;
        // chaining
        CK(M->d_seeds.reserve(n_match)); CK(M->d_f.reserve(n_match)); CK(M->d_p.reserve(n_match));
        CK(M->d_t.reserve(n_match)); CK(M->d_v.reserve(n_match)); CK(M->d_u.reserve(n_match));
        const uint32_t hit_cap = std::max<uint32_t>(1024, std::min<uint64_t>((uint64_t)n_seg * 20, (uint64_t)nr * 64 + 4096));
        CK(M->d_hits.reserve(hit_cap));
        CK(cudaMemsetAsync(M->d_counter.p + 1, 0, 4, s));
        CK(cudaMemsetAsync(M->d_seg_info.p, 0, sizeof(uint32_t) * 2 * n_seg, s));
        ChainScratch sc{M->d_seeds.p, M->d_f.p, M->d_p.p, M->d_t.p, M->d_v.p, M->d_u.p};
        chain_kernel<<<nblocks((uint64_t)n_seg * 32, 128), 128, 0, s>>>(K, V, M->d_seg_start.p, n_seg, n_match, M->soff_bits, M->d_seq_off.p,
                                                         (int)M->sseqs.size(), k, min_cnt, min_score, sc, M->d_hits.p, hit_cap,
                                                         M->d_counter.p + 1, M->d_seg_info.p); ++g_lesv_launches;
        CK(cudaEventRecord(ev[5], s));
        CK(cudaMemcpyAsync(&n_hits, M->d_counter.p + 1, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (n_hits > hit_cap) { ctx->err = "candidate hit buffer overflow"; return LESVGPU_E_NOMEM; }
        hhits.resize(n_hits); seg_info.resize(2 * (size_t)n_seg);
        if (n_hits) CK(cudaMemcpyAsync(hhits.data(), M->d_hits.p, sizeof(DdfHit) * n_hits, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(seg_info.data(), M->d_seg_info.p, sizeof(uint32_t) * 2 * n_seg, cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(ev[6], s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        cudaEventElapsedTime(&M->tm.emit_ms, ev[2], ev[3]);
        cudaEventElapsedTime(&M->tm.sort_ms, ev[3], ev[4]);
        cudaEventElapsedTime(&M->tm.chain_ms, ev[4], ev[5]);
        cudaEventElapsedTime(&M->tm.d2h_ms, ev[5], ev[6]);
    }
    cudaEventElapsedTime(&M->tm.h2d_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&M->tm.lookup_ms, ev[1], ev[2]);
    M->tm.n_kmers = nkm; M->tm.n_matches = n_match; M->tm.n_segments = n_seg; M->tm.n_hits = n_hits;
    M->tm.h2d_bytes = nbytes + sizeof(ReadDev) * nr;
    M->tm.d2h_bytes = sizeof(DdfHit) * n_hits + 8ull * n_seg;

    // ---- host: order the hits as the reference appends them (context, contig, rank), then windows ----
    std::sort(hhits.begin(), hhits.end(), [](const DdfHit& a, const DdfHit& b) {
        const int sa = a.seg & 0x3FFFFFF, sb = b.seg & 0x3FFFFFF;
        return sa != sb ? sa < sb : (a.seg >> 26) < (b.seg >> 26);
    });
    size_t hp = 0;
    std::vector<lesvgpu_init_hit> per_read;
    for (int r = 0; r < nr; ++r) {
        per_read.clear();
        const int qsize = (int)reads[r].len;
        while (hp < hhits.size()) {
            const int sg = hhits[hp].seg & 0x3FFFFFF;
            const uint32_t c = seg_info[2 * sg];
            if ((int)(c >> 1) != r) break;
            lesvgpu_init_hit ih;
            ih.qid = qfrom + r; ih.qdir = (int)(c & 1); ih.sid = (int)seg_info[2 * sg + 1];
            ih.qbeg = hhits[hp].qbeg; ih.qend = hhits[hp].qend; ih.sbeg = hhits[hp].sbeg; ih.send = hhits[hp].send;
            ih.qoff = ih.qend; ih.soff = ih.send;       // fill_init_hit_info :91-92
            ih.score = hhits[hp].score;
            per_read.push_back(ih);
            ++hp;
        }
        M->init_hits.insert(M->init_hits.end(), per_read.begin(), per_read.end());
        M->hit_first[r + 1] = (int)M->init_hits.size();
        candidate_windows(M, per_read, qfrom + r, qsize, opts->hitlist_size, M->windows);
        M->win_first[r + 1] = (int)M->windows.size();
    }
    if (windows) *windows = M->windows.data();
    return LESVGPU_OK;
}

extern "C" int lesvgpu_map_last_init_hits(const lesvgpu_ctx* ctx, const lesvgpu_init_hit** hits, const int** first)
{
    if (!ctx || !ctx->map) return LESVGPU_E_ARG;
    if (hits) *hits = ctx->map->init_hits.data();
    if (first) *first = ctx->map->hit_first.data();
    return LESVGPU_OK;
}

extern "C" int lesvgpu_map_last_matches(lesvgpu_ctx* ctx, int qfrom, const lesvgpu_kmer_match** matches, size_t* n)
{
    if (!ctx || !ctx->map || !matches || !n) return LESVGPU_E_ARG;
    lesvmap_state* M = ctx->map;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return LESVGPU_E_CUDA;
    const uint32_t nm = M->n_match;
    M->matches_host.resize(nm);
    if (nm) {
        std::vector<uint64_t> K(nm), V(nm);
        CK(cudaMemcpy(K.data(), M->sorted_in ? M->d_k1.p : M->d_k0.p, sizeof(uint64_t) * nm, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(V.data(), M->sorted_in ? M->d_v1.p : M->d_v0.p, sizeof(uint64_t) * nm, cudaMemcpyDeviceToHost));
        const uint64_t m = (1ull << M->soff_bits) - 1;
        for (uint32_t i = 0; i < nm; ++i) {
            M->matches_host[i].context = (int)(K[i] >> M->soff_bits) + 2 * qfrom;
            M->matches_host[i].qoff = (int)V[i];
            M->matches_host[i].soff = (int64_t)(K[i] & m);
        }
    }
    *matches = M->matches_host.data(); *n = nm;
    return LESVGPU_OK;
}

extern "C" void lesvgpu_map_last_timing(const lesvgpu_ctx* ctx, lesvgpu_map_timing* t)
{
    if (ctx && ctx->map && t) *t = ctx->map->tm;
}
