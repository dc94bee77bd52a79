// mix.cu -- what the SM sub-partitions can issue on K-ksw's block update alone (sm_100a): the real block_row() of
// ksw_core.cuh in a loop on registers, no shared memory, no barrier, no traceback store, with 1..8 warps per
// sub-partition.  Reports warp-instructions-equivalent throughput as block rows per kilo-cycle per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../lesv_b200/csrc/ksw_core.cuh"
using namespace lesvksw;

// kMode bit 0: one CTA barrier per row; bit 1: column-15 values through shared memory (store own, load the left neighbour's);
// bit 2: 16 traceback bytes per thread per row to global memory
template <int kMaxT, int kMode> __global__ void __launch_bounds__(kMaxT, 1) mix(uint32_t* out, long long* cyc, int iters, uint32_t seed, uint4* tbuf)
{
    __shared__ __align__(16) uint2 carry_all[1024][2];
    uint32_t rd = (uint32_t)__cvta_generic_to_shared(&carry_all[threadIdx.x ? threadIdx.x - 1 : blockDim.x - 1][1]);
    uint32_t wr = (uint32_t)__cvta_generic_to_shared(&carry_all[threadIdx.x][0]);
    uint4* tp = tbuf + (size_t)blockIdx.x * 1024 * 64 + threadIdx.x;
    BlockState b;
    b.init();
    uint32_t S[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) S[j] = (threadIdx.x + j) & 1 ? kSMatch : kSMatch + kSStep * 0x00010001u;
    Carry cin = carry_init(), cout;
    uint32_t acc = seed;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        uint32_t tb[4];
        if (kMode & 2) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cin.xv), "=r"(cin.x2) : "r"(rd) : "memory");
        block_row(b, S, [&]() -> Carry { return cin; }, cout, tb);
        if (kMode & 4) tp[(size_t)(it & 63) * 1024] = make_uint4(tb[0], tb[1], tb[2], tb[3]);
        else acc ^= tb[0] ^ tb[1] ^ tb[2] ^ tb[3];
        if (kMode & 2) {
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(wr), "r"(cout.xv), "r"(cout.x2) : "memory");
            rd ^= 8u; wr ^= 8u;
        } else { cin.xv = cout.xv ^ (acc & 0x00010000u); cin.x2 = cout.x2; }
        if (kMode & 1) __syncthreads();
        S[it & 7] ^= (acc & 0x07000700u) * 0;      // keep S in registers without changing it
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ b.U[0] ^ b.V[3] ^ b.X[5] ^ b.Y2[7];
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + threadIdx.x / 32] = t1 - t0;
}

template <int kMode> static void run_mode(uint32_t* d_out, long long* d_cyc, uint4* d_tb, int iters, const char* what)
{
    printf("== %s\n", what);
    for (int wps = 1; wps <= 8; ++wps) {           // 128 registers per thread up to 4 warps per sub-partition, 80 up to 6, 64 up to 8
        const int nt = 128 * wps;
        for (int rep = 0; rep < 2; ++rep) {
            if (wps <= 4) mix<512, kMode><<<148, nt>>>(d_out, d_cyc, iters, 12345u, d_tb);
            else if (wps <= 6) mix<768, kMode><<<148, nt>>>(d_out, d_cyc, iters, 12345u, d_tb);
            else mix<1024, kMode><<<148, nt>>>(d_out, d_cyc, iters, 12345u, d_tb);
            cudaDeviceSynchronize();
        }
        long long h[32], mx = 0;
        cudaMemcpy(h, d_cyc, sizeof(long long) * (nt / 32), cudaMemcpyDeviceToHost);
        for (int i = 0; i < nt / 32; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%d warps/SMSP: %.1f cycles per block row per warp, %.3f block rows per kcycle per SMSP\n", wps,
               (double)mx / iters, 1000.0 * iters * wps / (double)mx);
    }
}

int main()
{
    uint32_t* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 148 * 1024 * 4); cudaMalloc(&d_cyc, 148 * 32 * 8);
    const int iters = 20000;
    uint4* d_tb; cudaMalloc(&d_tb, (size_t)148 * 1024 * 64 * 16);
    run_mode<0>(d_out, d_cyc, d_tb, iters, "block update only");
    run_mode<1>(d_out, d_cyc, d_tb, iters, "+ CTA barrier per row");
    run_mode<3>(d_out, d_cyc, d_tb, iters, "+ barrier + carry through shared memory");
    run_mode<7>(d_out, d_cyc, d_tb, iters, "+ barrier + carry + traceback store");
    run_mode<4>(d_out, d_cyc, d_tb, iters, "+ traceback store only");
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
