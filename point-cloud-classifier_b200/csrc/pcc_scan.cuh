// Exclusive prefix sums over int64 counters (segment offsets, CSR row pointers).
#pragma once
#include "pcc_common.cuh"

#define PCC_SCAN_SINGLE_MAX (1u << 20)

namespace pcc {

// block-wide exclusive scan of one value per thread (1024 threads); returns exclusive
// prefix, writes block total to *total (valid in every thread after the call).
__device__ __forceinline__ long long block_exclusive_scan_1024(long long v, long long* total) {
  __shared__ long long warp_tot[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    long long w = warp_tot[lane];
    long long winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tot[lane] = winc;  // inclusive over warps
  }
  __syncthreads();
  long long base = wid > 0 ? warp_tot[wid - 1] : 0;
  *total = warp_tot[31];
  long long res = base + inc - v;
  __syncthreads();
  return res;
}

// in-place exclusive scan by ONE block of 1024 threads, chunk by chunk with a carry.
static __global__ void __launch_bounds__(1024) exclusive_scan_single_block_kernel(int64_t* data, int64_t count) {
  long long carry = 0;
  for (int64_t base = 0; base < count; base += 1024) {
    int64_t i = base + threadIdx.x;
    long long v = (i < count) ? data[i] : 0;
    long long tot;
    long long ex = block_exclusive_scan_1024(v, &tot);
    if (i < count) data[i] = ex + carry;
    carry += tot;
  }
}

// multi-block: phase 1 scans 1024-element chunks in place and writes chunk totals
static __global__ void __launch_bounds__(1024) scan_chunks_kernel(int64_t* data, int64_t count, int64_t* chunk_tot) {
  int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  long long v = (i < count) ? data[i] : 0;
  long long tot;
  long long ex = block_exclusive_scan_1024(v, &tot);
  if (i < count) data[i] = ex;
  if (threadIdx.x == 0) chunk_tot[blockIdx.x] = tot;
}

static __global__ void __launch_bounds__(1024) scan_add_kernel(int64_t* data, int64_t count, const int64_t* chunk_pre) {
  int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  if (i < count) data[i] += chunk_pre[blockIdx.x];
}

// in-place exclusive scan of data[count]; ws must hold cdiv(count,1024) int64
inline void exclusive_scan_i64(int64_t* data, int64_t count, int64_t* ws, cudaStream_t st) {
  if (count <= 4096) {
    PCC_K(exclusive_scan_single_block_kernel)<<<1, 1024, 0, st>>>(data, count);
    return;
  }
  int64_t chunks = cdiv(count, 1024);
  PCC_K(scan_chunks_kernel)<<<(unsigned)chunks, 1024, 0, st>>>(data, count, ws);
  PCC_K(exclusive_scan_single_block_kernel)<<<1, 1024, 0, st>>>(ws, chunks);
  PCC_K(scan_add_kernel)<<<(unsigned)chunks, 1024, 0, st>>>(data, count, ws);
}

}  // namespace pcc
