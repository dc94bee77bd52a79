// ctx_internal.cuh -- the context behind the C ABI (include/lesvgpu.h), shared by the ksw and the
// mapper translation units.  Not part of the public interface.
#pragma once
#include "../../include/lesvgpu.h"
#include "ksw_kernels.cuh"

#include <cstdio>
#include <string>
#include <vector>

namespace lesvctx {

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
    cudaError_t reserve_exact(size_t n)              // no headroom: for the traceback arena, sized against free memory
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
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

}  // namespace lesvctx
using lesvctx::DevBuf;
using lesvctx::PinBuf;

constexpr int kSideStreams = 8;

struct lesvmap_state;                             // mapper stages (map_*.cu)
void lesvmap_state_free(lesvmap_state*);

struct lesvgpu_ctx {
    int device = 0;
    int n_sm = 0;
    size_t arena_limit = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side[kSideStreams] = {};        // launch classes of one batch run concurrently
    cudaEvent_t side_done[kSideStreams] = {};
    cudaEvent_t ev[8] = {};
    std::string err;
    lesvgpu_timing timing = {};
    std::vector<lesvgpu_class_timing> classes;    // launch classes of the last K-ksw batch
    DevBuf<uint8_t> d_raw, d_arena, d_subject;
    size_t subject_bytes = 0;
    DevBuf<uint32_t> d_pool, d_cigar, d_scratch;
    DevBuf<lesvksw::KswJob> d_jobs;
    DevBuf<lesvksw::KswOut> d_outs;
    DevBuf<lesvksw::KswGapRun> d_runs;
    DevBuf<int> d_order, d_counter;
    DevBuf<size_t> d_aoff;
    DevBuf<lesvksw::KswStripJob> d_sjobs;          // column strips of full-matrix alignments
    DevBuf<uint2> d_edges;
    DevBuf<int> d_sprog;
    PinBuf<uint8_t> h_raw;
    PinBuf<uint32_t> h_cigar;
    PinBuf<lesvksw::KswOut> h_outs;
    PinBuf<lesvksw::KswGapRun> h_runs;
    std::vector<lesvgpu_svsig> sigs;
    std::vector<int> sig_subject;                 // contig of every record in sigs (lesvgpu_svsig_batch_contigs)
    std::vector<lesvgpu_svread> svreads;          // result of lesvgpu_svr_batch
    lesvmap_state* map = nullptr;                 // index + scratch of the mapping stages
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

