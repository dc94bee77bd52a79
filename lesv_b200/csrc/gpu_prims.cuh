// gpu_prims.cuh -- device-wide building blocks of the mapper stages: exclusive scan and a stable
// LSD radix sort of (u64 key, u64 payload) records.  Hand-written for sm_100a; no CUB/Thrust.
//
// Why stable LSD: the reference's index is "k-mers sorted by hash, ties in ascending offset" because
// its own sort is a stable 8-bit LSD radix sort (reference src/algo/hash_list_bucket_sort.c:93-235),
// and the seed lists of a read are consumed in (subject offset, query offset) order.  Emitting
// records in payload order and sorting them stably by key reproduces both without a second key.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

// kernels launched by the mapper stages of this process (lesvgpu_launch_count, include/lesvgpu.h)
#include <atomic>
extern std::atomic<unsigned long long> g_lesv_launches;

// The dynamic-shared-memory ceiling of a kernel is a property of the FUNCTION, shared by every context and feeder thread of
// the process (one per device): a launch that sets it to what it needs can see another thread lower it again before the
// kernel is issued ("invalid argument", one run in ten with three feeders on a device).  lesv_need_smem only ever RAISES
// the ceiling of (device, function), under a lock, to the largest amount any launch has asked for so far -- not to the
// device maximum: the driver sizes the shared-memory carve-out by this attribute, and the maximum costs the kernels
// 1-2 % through the smaller L1.
cudaError_t lesv_need_smem(const void* func, size_t bytes);
template <class F> inline cudaError_t lesv_need_smem(F* func, size_t bytes) { return lesv_need_smem(reinterpret_cast<const void*>(func), bytes); }

namespace lesvmap {

// out[i] = sum_{j<i} in[j]; returns nothing, the grand total lands in *d_total (may be null).
// tmp must hold scan_tmp_words(n) uint32 words.
size_t scan_tmp_words(size_t n);
void exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, size_t n, uint32_t* d_total,
                        uint32_t* d_tmp, cudaStream_t s);

// Stable sort of n records by bits [bit_lo, bit_hi) of the key.  Ping-pongs between (k0,v0) and
// (k1,v1); returns 0 if the sorted records end up in (k0,v0), 1 if in (k1,v1).
// tmp: radix_tmp_words(n) uint32 words.  v0 == v1 == nullptr sorts keys only.
size_t radix_tmp_words(size_t n);
int radix_sort_pairs(uint64_t* k0, uint64_t* v0, uint64_t* k1, uint64_t* v1, size_t n,
                     int bit_lo, int bit_hi, uint32_t* d_tmp, cudaStream_t s);

}  // namespace lesvmap
