// lesvgpu.cu -- host side of the C ABI (include/lesvgpu.h): contexts, batching, scheduling.
//
// The reference calls its aligner from T pthreads, one pair at a time.  Here one
// feeder thread hands a whole batch to one GPU: pairs are binned by band width into the thread
// counts the DP kernel is instantiated for, each bin is sorted by decreasing traceback size and
// pulled by persistent CTAs, and every CTA reuses one traceback arena in HBM.
#include "../../include/lesvgpu.h"
#include "ksw_kernels.cuh"
#include "ksw_core.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace lesvksw;

namespace {

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <class T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMallocHost((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

struct lesvgpu_ctx {
    int device = 0;
    int n_sm = 0;
    size_t arena_limit = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    std::string err;
    lesvgpu_timing timing = {};
    DevBuf<uint8_t> d_raw, d_arena;
    DevBuf<uint32_t> d_pool, d_cigar;
    DevBuf<KswJob> d_jobs;
    DevBuf<KswOut> d_outs;
    DevBuf<int> d_order, d_counter;
    PinBuf<uint8_t> h_raw;
    PinBuf<uint32_t> h_cigar;
    PinBuf<KswOut> h_outs;
};

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            char b_[512];                                                                     \
            snprintf(b_, sizeof b_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            ctx->err = b_;                                                                    \
            return LESVGPU_E_CUDA;                                                            \
        }                                                                                     \
    } while (0)

extern "C" int lesvgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int lesvgpu_ctx_create(int device, size_t arena_bytes, lesvgpu_ctx** out)
{
    if (!out) return LESVGPU_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return LESVGPU_E_NOGPU;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LESVGPU_E_NOGPU;
    if (prop.major != 10) return LESVGPU_E_NOGPU;          // sm_100a code only
    if (cudaSetDevice(device) != cudaSuccess) return LESVGPU_E_NOGPU;
    lesvgpu_ctx* ctx = new lesvgpu_ctx;
    ctx->device = device;
    ctx->n_sm = prop.multiProcessorCount;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    ctx->arena_limit = arena_bytes ? arena_bytes : (size_t)(free_b * 0.8);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LESVGPU_E_CUDA; }
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    *out = ctx;
    return LESVGPU_OK;
}

extern "C" void lesvgpu_ctx_destroy(lesvgpu_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->d_raw.release(); ctx->d_arena.release(); ctx->d_pool.release(); ctx->d_cigar.release();
    ctx->d_jobs.release(); ctx->d_outs.release(); ctx->d_order.release(); ctx->d_counter.release();
    ctx->h_raw.release(); ctx->h_cigar.release(); ctx->h_outs.release();
    for (auto& e : ctx->ev) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* lesvgpu_last_error(const lesvgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void lesvgpu_last_timing(const lesvgpu_ctx* ctx, lesvgpu_timing* t) { if (ctx && t) *t = ctx->timing; }

static inline int effective_band(int qlen, int tlen, int max_distance)
{
    // ksw2_wrapper.c:121: a zero band means 10 % of the longer sequence (double product, truncated)
    if (max_distance == 0) max_distance = (int)((qlen > tlen ? qlen : tlen) * 0.1);
    // ksw2_extd2_sse.c:82: negative = no band
    if (max_distance < 0) max_distance = qlen > tlen ? qlen : tlen;
    return max_distance;
}

static inline bool band_reaches_end(int qlen, int tlen, int w)
{
    // st0 <= en0 on every anti-diagonal (ksw2_extd2_sse.c:130-137)  <=>  |qlen - tlen| <= w, and a
    // zero band only survives a 1x1 matrix.
    const int d = qlen > tlen ? qlen - tlen : tlen - qlen;
    if (d > w) return false;
    if (w == 0 && qlen + tlen - 1 > 1) return false;
    return true;
}

extern "C" int64_t lesvgpu_ksw_cells(int qlen, int tlen, int max_distance, int* ok)
{
    const int w = effective_band(qlen, tlen, max_distance);
    int64_t cells = 0;
    if (ok) *ok = 1;
    for (int r = 0; r < qlen + tlen - 1; ++r) {
        const RowGeom g = row_geom(r, qlen, tlen, w);
        if (g.empty()) { if (ok) *ok = 0; break; }
        cells += g.en0 - g.st0 + 1;
    }
    return cells;
}

extern "C" int lesvgpu_ksw_batch(lesvgpu_ctx* ctx, const lesvgpu_ksw_pair* pairs, int n,
                                 lesvgpu_ksw_result* results, const uint32_t** cigar_pool)
{
    if (!ctx || n < 0 || (n && (!pairs || !results))) return LESVGPU_E_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return LESVGPU_E_CUDA;
    ctx->timing = lesvgpu_timing{};
    if (cigar_pool) *cigar_pool = nullptr;
    if (n == 0) return LESVGPU_OK;

    // ---- plan ---------------------------------------------------------------------------------
    static const int kClasses[4] = {64, 128, 256, 512};
    std::vector<KswJob> jobs;
    std::vector<int> job_pair;                      // job -> pair index
    std::vector<std::vector<int>> cls(4);
    std::vector<size_t> tb_bytes;
    jobs.reserve(n); job_pair.reserve(n); tb_bytes.reserve(n);
    size_t raw_bytes = 0, pool_words = 0, cigar_words = 0;
    for (int i = 0; i < n; ++i) {
        lesvgpu_ksw_result& R = results[i];
        R.ok = 0; R.n_cigar = 0; R.cigar_off = 0; R.aln_size = 0; R.n_match = 0;
        const lesvgpu_ksw_pair& P = pairs[i];
        if (P.qlen <= 0 || P.tlen <= 0) continue;              // ksw2_extd2_sse.c:68 -> n_cigar 0
        if (!P.query || !P.target) { ctx->err = "null sequence pointer"; return LESVGPU_E_ARG; }
        const int w = effective_band(P.qlen, P.tlen, P.max_distance);
        if (!band_reaches_end(P.qlen, P.tlen, w)) continue;    // partial CIGAR -> wrapper returns 0
        const int wb = window_blocks(P.tlen, w);
        int c = 0;
        while (c < 4 && kClasses[c] < wb) ++c;
        if (c == 4) {
            char b[160];
            snprintf(b, sizeof b, "pair %d: band of %d columns needs %d blocks (> 512): wide kernel not built yet", i, w, wb);
            ctx->err = b;
            return LESVGPU_E_NOMEM;
        }
        KswJob J;
        memset(&J, 0, sizeof J);
        J.qlen = P.qlen; J.tlen = P.tlen; J.w = w;
        J.tb_blocks = tb_row_blocks(P.qlen, P.tlen, w);
        J.seq_off_q = raw_bytes; raw_bytes += (size_t)P.qlen;
        J.seq_off_t = raw_bytes; raw_bytes += (size_t)P.tlen;
        J.tp_off = (uint32_t)pool_words; pool_words += tp_words(P.tlen);
        J.qrp_off = (uint32_t)pool_words; pool_words += qrp_words(P.qlen);
        J.cigar_off = cigar_words; J.cigar_cap = (uint32_t)(P.qlen + P.tlen + 2);
        cigar_words += J.cigar_cap;
        if (pool_words > 0xFFFFFFF0ull) { ctx->err = "batch too large (packed pool offsets are 32-bit)"; return LESVGPU_E_ARG; }
        const size_t tb = (size_t)(P.qlen + P.tlen - 1) * (size_t)J.tb_blocks * 16;
        if (tb > ctx->arena_limit) {
            ctx->err = "one alignment's traceback does not fit the arena";
            return LESVGPU_E_NOMEM;
        }
        cls[c].push_back((int)jobs.size());
        jobs.push_back(J); job_pair.push_back(i); tb_bytes.push_back(tb);
    }
    const int nj = (int)jobs.size();
    if (nj == 0) return LESVGPU_OK;

    // ---- upload -------------------------------------------------------------------------------
    CK(ctx->h_raw.reserve(raw_bytes));
    for (int k = 0; k < nj; ++k) {
        const lesvgpu_ksw_pair& P = pairs[job_pair[k]];
        memcpy(ctx->h_raw.p + jobs[k].seq_off_q, P.query, (size_t)P.qlen);
        memcpy(ctx->h_raw.p + jobs[k].seq_off_t, P.target, (size_t)P.tlen);
    }
    std::vector<int> order;
    order.reserve(nj);
    struct Launch { int nt, first, count, n_ctas; size_t stride, smem; };
    std::vector<Launch> launches;
    size_t arena_need = 0;
    for (int c = 0; c < 4; ++c) {
        if (cls[c].empty()) continue;
        std::sort(cls[c].begin(), cls[c].end(), [&](int a, int b) {
            return tb_bytes[a] != tb_bytes[b] ? tb_bytes[a] > tb_bytes[b] : a < b;
        });
        Launch L;
        L.nt = kClasses[c]; L.first = (int)order.size(); L.count = (int)cls[c].size();
        L.stride = (tb_bytes[cls[c][0]] + 255) & ~(size_t)255;
        int mq = 0, mt = 0;
        for (int k : cls[c]) { mq = std::max(mq, jobs[k].qlen); mt = std::max(mt, jobs[k].tlen); order.push_back(k); }
        L.smem = ksw_smem_bytes(mq, mt);
        int per_sm = ksw_ctas_per_sm(L.nt);
        const size_t smem_cap = 200 * 1024;
        while (per_sm > 1 && (L.smem + 4096) * per_sm > smem_cap) --per_sm;
        int ctas = std::min(L.count, ctx->n_sm * per_sm);
        const size_t by_mem = ctx->arena_limit / L.stride;
        if ((size_t)ctas > by_mem) ctas = (int)by_mem;
        if (ctas < 1) { ctx->err = "arena too small"; return LESVGPU_E_NOMEM; }
        L.n_ctas = ctas;
        arena_need = std::max(arena_need, L.stride * (size_t)ctas);
        launches.push_back(L);
    }
    CK(ctx->d_raw.reserve(raw_bytes));
    CK(ctx->d_pool.reserve(pool_words));
    CK(ctx->d_cigar.reserve(cigar_words));
    CK(ctx->d_jobs.reserve(nj));
    CK(ctx->d_outs.reserve(nj));
    CK(ctx->d_order.reserve(nj));
    CK(ctx->d_counter.reserve(launches.size()));
    CK(ctx->d_arena.reserve(arena_need));
    CK(ctx->h_outs.reserve(nj));
    CK(ctx->h_cigar.reserve(cigar_words));

    cudaStream_t s = ctx->stream;
    CK(cudaEventRecord(ctx->ev[0], s));
    CK(cudaMemcpyAsync(ctx->d_raw.p, ctx->h_raw.p, raw_bytes, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_jobs.p, jobs.data(), sizeof(KswJob) * nj, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_order.p, order.data(), sizeof(int) * nj, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->d_counter.p, 0, sizeof(int) * launches.size(), s));
    CK(cudaMemsetAsync(ctx->d_outs.p, 0, sizeof(KswOut) * nj, s));
    CK(cudaEventRecord(ctx->ev[1], s));
    launch_pack(ctx->d_jobs.p, nj, ctx->d_raw.p, ctx->d_pool.p, s);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[2], s));
    for (size_t li = 0; li < launches.size(); ++li) {
        const Launch& L = launches[li];
        launch_ksw(L.nt, L.n_ctas, L.smem, ctx->d_jobs.p, ctx->d_order.p + L.first, L.count, ctx->d_pool.p,
                   ctx->d_arena.p, L.stride, ctx->d_cigar.p, ctx->d_outs.p, ctx->d_counter.p + li, s);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ctx->ev[3], s));
    CK(cudaMemcpyAsync(ctx->h_outs.p, ctx->d_outs.p, sizeof(KswOut) * nj, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_cigar.p, ctx->d_cigar.p, sizeof(uint32_t) * cigar_words, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->ev[4], s));
    CK(cudaStreamSynchronize(s));

    for (int k = 0; k < nj; ++k) {
        lesvgpu_ksw_result& R = results[job_pair[k]];
        const KswOut& O = ctx->h_outs.p[k];
        R.n_cigar = O.n_cigar; R.cigar_off = jobs[k].cigar_off;
        R.aln_size = O.aln_size; R.n_match = O.n_match;
        R.ok = O.n_cigar > 0 && O.status == 0;
    }
    if (cigar_pool) *cigar_pool = ctx->h_cigar.p;
    lesvgpu_timing& T = ctx->timing;
    cudaEventElapsedTime(&T.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&T.pack_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&T.dp_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&T.d2h_ms, ctx->ev[3], ctx->ev[4]);
    T.dp_launches = (int)launches.size();
    T.kernel_launches = 1 + (int)launches.size();
    T.h2d_bytes = raw_bytes + sizeof(KswJob) * nj + sizeof(int) * nj;
    T.d2h_bytes = sizeof(KswOut) * nj + sizeof(uint32_t) * cigar_words;
    return LESVGPU_OK;
}
