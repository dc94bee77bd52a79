// gpu_prims.cu -- exclusive scan and stable LSD radix sort (see gpu_prims.cuh).
#include "gpu_prims.cuh"

std::atomic<unsigned long long> g_lesv_launches{0};   // several feeder threads (one per GPU) count here

#include <map>
#include <mutex>
#include <utility>
cudaError_t lesv_need_smem(const void* func, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> ceiling;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = ceiling[{dev, func}];
    if (bytes <= have) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

namespace lesvmap {

// ---------------------------------------------------------------------------------------------
// exclusive scan (u32)
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_tot, uint32_t* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, d);
            if (lane >= d) winc += t;
        }
        if (lane < (int)(blockDim.x >> 5)) warp_tot[lane] = winc - w;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    return inc - v + warp_tot[warp];
}

__global__ void __launch_bounds__(kScanThreads)
scan_local_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, uint32_t* __restrict__ sums)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t total;
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t x[kScanItems];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        x[k] = base + k < n ? in[base + k] : 0u;
        mine += x[k];
    }
    uint32_t pre = block_exclusive_scan(mine, warp_tot, &total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = pre;
        pre += x[k];
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void scan_add_kernel(uint32_t* __restrict__ out, size_t n, const uint32_t* __restrict__ offs)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += offs[i / kScanTile];
}

size_t scan_tmp_words(size_t n)
{
    size_t w = 0;
    while (n > 1) {
        n = (n + kScanTile - 1) / kScanTile;
        w += (n + 63) & ~(size_t)63;
        if (n == 1) break;
    }
    return w + 64;
}

void exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, size_t n, uint32_t* d_total,
                        uint32_t* d_tmp, cudaStream_t s)
{
    if (n == 0) {
        if (d_total) cudaMemsetAsync(d_total, 0, sizeof(uint32_t), s);
        return;
    }
    const size_t nb = (n + kScanTile - 1) / kScanTile;
    uint32_t* sums = d_tmp;
    scan_local_kernel<<<(unsigned)nb, kScanThreads, 0, s>>>(d_in, d_out, n, sums); ++g_lesv_launches;
    if (nb == 1) {
        if (d_total) cudaMemcpyAsync(d_total, sums, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
        return;
    }
    exclusive_scan_u32(sums, sums, nb, d_total, d_tmp + ((nb + 63) & ~(size_t)63), s);
    scan_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_out, n, sums); ++g_lesv_launches;
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass
// ---------------------------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsChunks = 32;                           // chunks of 256 records per block
constexpr int kRsTile = kRsThreads * kRsChunks;

__global__ void __launch_bounds__(kRsThreads)
rs_hist_kernel(const uint64_t* __restrict__ keys, size_t n, int shift, uint32_t mask,
               uint32_t* __restrict__ hist, unsigned nblk)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * kRsTile;
#pragma unroll 4
    for (int c = 0; c < kRsChunks; ++c) {
        const size_t i = base + (size_t)c * kRsThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];       // bin-major: scan order = output order
}

// Every chunk of 256 records is ranked in input order: a warp's lanes with the same digit are ranked
// by lane (__match_any_sync), warps are ordered by a 256-thread pass over the per-warp digit counts.
__global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals,
                  uint64_t* __restrict__ okeys, uint64_t* __restrict__ ovals, size_t n, int shift, uint32_t mask,
                  const uint32_t* __restrict__ offs, unsigned nblk)
{
    __shared__ uint32_t next[256];                      // where the next record of digit d goes
    __shared__ uint32_t wcnt[kRsThreads / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    next[threadIdx.x] = offs[(size_t)threadIdx.x * nblk + blockIdx.x];
    const size_t base = (size_t)blockIdx.x * kRsTile;
    for (int c = 0; c < kRsChunks; ++c) {
        const size_t i = base + (size_t)c * kRsThreads + threadIdx.x;
        if (base + (size_t)c * kRsThreads >= n) break;
#pragma unroll
        for (int w = 0; w < kRsThreads / 32; ++w) wcnt[w][threadIdx.x] = 0;
        __syncthreads();
        const bool valid = i < n;
        uint64_t k = 0, v = 0;
        uint32_t d = 0, rank = 0;
        const unsigned vm = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid) {
            k = keys[i]; v = vals ? vals[i] : 0ull;
            d = (uint32_t)(k >> shift) & mask;
            const unsigned same = __match_any_sync(vm, d);
            rank = __popc(same & ((1u << lane) - 1u));
            if (rank == 0) wcnt[warp][d] = __popc(same);
        }
        __syncthreads();
        {
            uint32_t acc = next[threadIdx.x];
#pragma unroll
            for (int w = 0; w < kRsThreads / 32; ++w) {
                const uint32_t t = wcnt[w][threadIdx.x];
                wcnt[w][threadIdx.x] = acc;
                acc += t;
            }
            next[threadIdx.x] = acc;
        }
        __syncthreads();
        if (valid) {
            const uint32_t pos = wcnt[warp][d] + rank;
            okeys[pos] = k;
            if (ovals) ovals[pos] = v;
        }
        __syncthreads();
    }
}

size_t radix_tmp_words(size_t n)
{
    const size_t nblk = (n + kRsTile - 1) / kRsTile;
    const size_t h = ((nblk * 256) + 63) & ~(size_t)63;
    return 2 * h + scan_tmp_words(nblk * 256) + 64;
}

int radix_sort_pairs(uint64_t* k0, uint64_t* v0, uint64_t* k1, uint64_t* v1, size_t n,
                     int bit_lo, int bit_hi, uint32_t* d_tmp, cudaStream_t s)
{
    if (n == 0 || bit_hi <= bit_lo) return 0;
    const unsigned nblk = (unsigned)((n + kRsTile - 1) / kRsTile);
    const size_t h = (((size_t)nblk * 256) + 63) & ~(size_t)63;
    uint32_t* hist = d_tmp;
    uint32_t* offs = d_tmp + h;
    uint32_t* stmp = d_tmp + 2 * h;
    int cur = 0;
    for (int lo = bit_lo; lo < bit_hi; lo += 8) {
        const int bits = bit_hi - lo < 8 ? bit_hi - lo : 8;
        const uint32_t mask = (1u << bits) - 1u;
        uint64_t* ki = cur ? k1 : k0; uint64_t* vi = cur ? v1 : v0;
        uint64_t* ko = cur ? k0 : k1; uint64_t* vo = cur ? v0 : v1;
        rs_hist_kernel<<<nblk, kRsThreads, 0, s>>>(ki, n, lo, mask, hist, nblk); ++g_lesv_launches;
        exclusive_scan_u32(hist, offs, (size_t)nblk * 256, nullptr, stmp, s);
        rs_scatter_kernel<<<nblk, kRsThreads, 0, s>>>(ki, vi, ko, vo, n, lo, mask, offs, nblk); ++g_lesv_launches;
        cur ^= 1;
    }
    return cur;
}

}  // namespace lesvmap
