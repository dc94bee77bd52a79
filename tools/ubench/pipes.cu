// pipes.cu -- issue-rate microbenchmark for the packed 16x2 integer instructions K-ksw is built from (sm_100a).
// For each instruction (and pairs of them, interleaved) it reports warp-instructions per clock per SM sub-partition
// with 1, 2, 4 warps per sub-partition.  Two instructions that sustain ~1.0 together but ~0.5 alone sit on different pipes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { VADD, VSUB, VMAX3, VMIN, VADDRELU, VADDMAX, LOP, PRMT_, IMAD_, SHF_, IADD_, NOPS };
static const char* kNames[] = {"VIADD.16x2", "VSUB.16x2", "VIMNMX3.16x2", "VIMNMX.16x2", "VIADDMNMX.relu", "VIADDMNMX.16x2", "LOP3", "PRMT", "IMAD", "SHF", "IADD3"};

template <int OP> __device__ __forceinline__ uint32_t apply(uint32_t a, uint32_t b, uint32_t c)
{
    if (OP == VADD) return __vadd2(a, b);
    if (OP == VSUB) return __vsub2(a, b);
    if (OP == VMAX3) return __vimax3_s16x2(a, b, c);
    if (OP == VMIN) return __vmins2(a, b);
    if (OP == VADDRELU) return __viaddmax_s16x2_relu(a, b, 0u);
    if (OP == VADDMAX) return __viaddmax_s16x2(a, b, c);
    if (OP == LOP) return (a & b) ^ c;
    if (OP == PRMT_) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(r) : "r"(a), "r"(b)); return r; }
    if (OP == IMAD_) return a * b + c;
    if (OP == SHF_) return __funnelshift_r(a, b, 7);
    if (OP == IADD_) return a + b + c;
    return a;
}

template <int OPA, int OPB>
__global__ void bench(uint32_t* out, long long* cyc, int iters, uint32_t seed)
{
    uint32_t x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = seed * (i + 1) + threadIdx.x; y[i] = seed * (i + 9) ^ threadIdx.x; }
    const uint32_t b = seed | 0x00010001u, c = seed * 7u + 3u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = apply<OPA>(x[i], b, c);
            y[i] = apply<OPB>(y[i], c, b);
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int OPA, int OPB> static void run(uint32_t* d_out, long long* d_cyc)
{
    const int iters = 4096;
    printf("%-16s + %-16s :", kNames[OPA], kNames[OPB]);
    for (int wps = 1; wps <= 4; wps *= 2) {
        const int nt = 128 * wps;
        bench<OPA, OPB><<<148, nt>>>(d_out, d_cyc, iters, 12345u);
        cudaDeviceSynchronize();
        bench<OPA, OPB><<<148, nt>>>(d_out, d_cyc, iters, 12345u);
        cudaDeviceSynchronize();
        long long h[32];
        cudaMemcpy(h, d_cyc, sizeof(long long) * (nt / 32), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < nt / 32; ++i) mx = h[i] > mx ? h[i] : mx;
        const double inst = (double)iters * 16.0 * wps;       // warp instructions per sub-partition
        printf("  %dw/SMSP %.3f ipc", wps, inst / (double)mx);
    }
    printf("\n");
}

int main()
{
    uint32_t* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 148 * 512 * 4); cudaMalloc(&d_cyc, 148 * 16 * 8);
    run<VADD, VADD>(d_out, d_cyc);
    run<VSUB, VSUB>(d_out, d_cyc);
    run<VMAX3, VMAX3>(d_out, d_cyc);
    run<VMIN, VMIN>(d_out, d_cyc);
    run<VADDRELU, VADDRELU>(d_out, d_cyc);
    run<VADDMAX, VADDMAX>(d_out, d_cyc);
    run<LOP, LOP>(d_out, d_cyc);
    run<PRMT_, PRMT_>(d_out, d_cyc);
    run<IMAD_, IMAD_>(d_out, d_cyc);
    run<SHF_, SHF_>(d_out, d_cyc);
    run<IADD_, IADD_>(d_out, d_cyc);
    run<VADD, VMAX3>(d_out, d_cyc);
    run<VADD, VADDRELU>(d_out, d_cyc);
    run<VADD, LOP>(d_out, d_cyc);
    run<VADD, IMAD_>(d_out, d_cyc);
    run<VADD, PRMT_>(d_out, d_cyc);
    run<VMAX3, LOP>(d_out, d_cyc);
    run<VMAX3, IMAD_>(d_out, d_cyc);
    run<VADDRELU, LOP>(d_out, d_cyc);
    run<VADDRELU, IMAD_>(d_out, d_cyc);
    run<IMAD_, LOP>(d_out, d_cyc);
    run<IMAD_, PRMT_>(d_out, d_cyc);
    run<LOP, PRMT_>(d_out, d_cyc);
    run<SHF_, IMAD_>(d_out, d_cyc);
    run<IADD_, IMAD_>(d_out, d_cyc);
    run<IADD_, VADD>(d_out, d_cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
