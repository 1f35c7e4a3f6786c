// Segment bookkeeping and ragged pooling.
//   reference: /root/reference/models/deep_sets.py:91-106 (bincount/split/pool loop),
//   PyG global_mean_pool called at /root/reference/models/graph_net.py:92,96.
// HBM-bound kernels: coalesced 128 B row slabs, one pass over x.
#include "pcc_common.cuh"
#include "pcc_scan.cuh"

namespace pcc {

// ------------------------------------------------------------------ histogram of idx
__global__ void zero_i64_kernel(int64_t* p, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0;
}

__global__ void histogram_kernel(const int64_t* __restrict__ idx, int64_t n, int64_t B,
                                 unsigned long long* __restrict__ counts) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t v = (i < n) ? idx[i] : -1;
  bool valid = (v >= 0 && v < B);
  // warp-aggregate: sorted idx makes most warps uniform
  unsigned mask = __match_any_sync(0xffffffffu, valid ? v : -1 - (int64_t)(threadIdx.x & 31));
  int leader = __ffs(mask) - 1;
  if (valid && (int)(threadIdx.x & 31) == leader) atomicAdd(&counts[v], (unsigned long long)__popc(mask));
}

// histogram + exclusive scan in ONE launch: every block adds its warp-aggregated counts into offsets[0..B) (zeroed
// by a memset node), takes a ticket in offsets[B], and the last block to finish scans the counts in place
// (offsets[b] = number of rows with idx < b, offsets[B] = n) — the reference's bincount + split bookkeeping
// (/root/reference/models/deep_sets.py:91-92) without a host sync and without three dependent launches.
__global__ void __launch_bounds__(1024) segment_offsets_fused_kernel(const int64_t* __restrict__ idx, int64_t n, int64_t B,
                                                                     int64_t* __restrict__ offsets) {
  pdl_enter();
  __shared__ bool s_last;
  unsigned long long* counts = reinterpret_cast<unsigned long long*>(offsets);
  for (int64_t i0 = (int64_t)blockIdx.x * 1024; i0 < n; i0 += (int64_t)gridDim.x * 1024) {
    const int64_t i = i0 + threadIdx.x;
    const int64_t v = (i < n) ? idx[i] : -1;
    const bool valid = (v >= 0 && v < B);
    const unsigned mask = __match_any_sync(0xffffffffu, valid ? v : -1 - (int64_t)(threadIdx.x & 31));
    const int leader = __ffs(mask) - 1;
    if (valid && (int)(threadIdx.x & 31) == leader) atomicAdd(&counts[v], (unsigned long long)__popc(mask));
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&counts[B], 1ull) == (unsigned long long)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  long long carry = 0;
  for (int64_t base = 0; base < B; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const long long v = (i < B) ? (long long)*reinterpret_cast<volatile unsigned long long*>(&counts[i]) : 0;
    long long tot;
    const long long ex = block_exclusive_scan_1024(v, &tot);
    if (i < B) offsets[i] = ex + carry;
    carry += tot;
  }
  if (threadIdx.x == 0) offsets[B] = carry;
}

__global__ void index_max_kernel(const int64_t* __restrict__ idx, int64_t n, long long* __restrict__ out) {
  long long m = LLONG_MIN;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    long long v = idx[i];
    m = v > m ? v : m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    long long t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__global__ void set_i64_kernel(long long* p, long long v) { *p = v; }

// ------------------------------------------------------------------ pooling forward
// block = 32 columns x 8 row lanes; grid = (B, ceil(H/32))
__global__ void __launch_bounds__(256) segment_pool_fwd_kernel(const float* __restrict__ x,
                                                               const int64_t* __restrict__ offsets, int64_t H,
                                                               int pooling, float* __restrict__ pooled,
                                                               int32_t* __restrict__ argmax) {
  const int b = blockIdx.x;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.y * 32 + cx;
  const int64_t s = offsets[b], e = offsets[b + 1];
  __shared__ float sv[8][33];
  __shared__ int si[8][33];
  float acc = (pooling == PCC_POOL_MAX) ? -INFINITY : 0.f;
  int arg = 0x7fffffff;
  if (c < H) {
    if (pooling == PCC_POOL_MAX) {
      for (int64_t r = s + ry; r < e; r += 8) {
        float v = __ldg(x + r * H + c);
        if (v > acc || arg == 0x7fffffff) { acc = v; arg = (int)r; }
      }
    } else {
      for (int64_t r = s + ry; r < e; r += 8) acc += __ldg(x + r * H + c);
    }
  }
  sv[ry][cx] = acc;
  si[ry][cx] = arg;
  __syncthreads();
  if (ry == 0 && c < H) {
    if (pooling == PCC_POOL_MAX) {
      float best = sv[0][cx];
      int bi = si[0][cx];
#pragma unroll
      for (int j = 1; j < 8; ++j) {
        float v = sv[j][cx];
        int i = si[j][cx];
        if (i != 0x7fffffff && (bi == 0x7fffffff || v > best || (v == best && i < bi))) { best = v; bi = i; }
      }
      if (bi == 0x7fffffff) { best = 0.f; bi = -1; }  // empty segment
      pooled[(int64_t)b * H + c] = best;
      argmax[(int64_t)b * H + c] = bi;
    } else {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += sv[j][cx];
      const float n = (float)(e - s);
      if (pooling == PCC_POOL_SUM) t = (e > s) ? t / sqrtf(n) : 0.f;
      else if (pooling == PCC_POOL_MEAN) t = (e > s) ? t / n : 0.f;
      pooled[(int64_t)b * H + c] = t;
    }
  }
}

// ------------------------------------------------------------------ pooling backward
// block handles 32 rows; set id per row by binary search over offsets.
__global__ void __launch_bounds__(256) segment_pool_bwd_kernel(const float* __restrict__ dpooled,
                                                               const int64_t* __restrict__ offsets,
                                                               const int32_t* __restrict__ argmax, int64_t n,
                                                               int64_t B, int64_t H, int pooling,
                                                               float* __restrict__ dx) {
  __shared__ int s_set[32];
  __shared__ float s_scale[32];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  if (threadIdx.x < 32) {
    int64_t r = r0 + threadIdx.x;
    int set = -1;
    float scale = 0.f;
    if (r < n) {
      int64_t lo = 0, hi = B;  // find b with offsets[b] <= r < offsets[b+1]
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (offsets[mid + 1] <= r) lo = mid + 1; else hi = mid;
      }
      if (lo < B && offsets[lo] <= r) {
        set = (int)lo;
        float cnt = (float)(offsets[lo + 1] - offsets[lo]);
        scale = pooling == PCC_POOL_SUM ? 1.f / sqrtf(cnt) : pooling == PCC_POOL_MEAN ? 1.f / cnt : 1.f;
      }
    }
    s_set[threadIdx.x] = set;
    s_scale[threadIdx.x] = scale;
  }
  __syncthreads();
  const int64_t total = 32 * H;
  for (int64_t t = threadIdx.x; t < total; t += blockDim.x) {
    int lr = (int)(t / H);
    int64_t c = t - (int64_t)lr * H;
    int64_t r = r0 + lr;
    if (r >= n) break;
    int set = s_set[lr];
    float g = 0.f;
    if (set >= 0) {
      float dp = __ldg(dpooled + (int64_t)set * H + c);
      if (pooling == PCC_POOL_MAX) g = (__ldg(argmax + (int64_t)set * H + c) == (int)r) ? dp : 0.f;
      else g = dp * s_scale[lr];
    }
    dx[r * H + c] = g;
  }
}

// ------------------------------------------------------------------ fused loss + row gather
// BCEWithLogitsLoss(reduction=mean) forward and its gradient in one pass (wrapper.py:38,64-67):
// loss = mean(max(z,0) - z*y + log1p(exp(-|z|))), dlogits = (sigmoid(z) - y) / count
__global__ void __launch_bounds__(1024) bce_logits_kernel(const float* __restrict__ z, const float* __restrict__ y,
                                                         int64_t count, float* __restrict__ loss,
                                                         float* __restrict__ dz) {
  pdl_enter();
  __shared__ float red[32];
  const float inv = 1.f / (float)count;
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float zi = z[i], yi = y[i];
    s += fmaxf(zi, 0.f) - zi * yi + log1pf(expf(-fabsf(zi)));
    dz[i] = (1.f / (1.f + expf(-zi)) - yi) * inv;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) t += red[j];
    if (gridDim.x == 1) *loss = t * inv;  // one block: plain store, no zero-fill before the launch, deterministic
    else atomicAdd(loss, t * inv);
  }
}

// out[i, :] = x[clamp(idx[i], 0, n-1), :]  (rows of d floats)
__global__ void gather_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, int64_t count, int d,
                                   int64_t n, float* __restrict__ out, const float* __restrict__ g_in,
                                   float* __restrict__ g_out) {
  pdl_enter();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= count * d) return;
  const int64_t i = t / d;
  const int j = (int)(t - i * d);
  int64_t r = idx[i];
  // a negative index marks "no row" (argmax of an empty set): its gradient entry is zeroed so that the clamped
  // row contributes nothing downstream
  if (g_out && j == 0) g_out[i] = (r >= 0) ? g_in[i] : 0.f;
  r = r < 0 ? 0 : (r >= n ? n - 1 : r);
  out[t] = __ldg(x + r * d + j);
}

}  // namespace pcc

using namespace pcc;

extern "C" int pcc_bce_logits(const float* logits, const float* target, int64_t count, float* loss, float* dlogits,
                              int device, void* stream) {
  PCC_ENTER(device);
  PCC_REQUIRE(count > 0, "empty logits");
  cudaStream_t st = (cudaStream_t)stream;
  // up to 4 K logits (the train step's [B, out] is 256 .. 2560): ONE block, so the launch needs no memset node
  // in front of it and stays a programmatic dependent of the head kernel before it
  const int blocks = (count <= 4096) ? 1 : (int)(cdiv(count, 256) < 148 ? cdiv(count, 256) : 148);
  if (blocks > 1) PCC_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  launch_dep(bce_logits_kernel, dim3(blocks), dim3(blocks == 1 ? 1024 : 256), 0, st, logits, target, count, loss, dlogits);
  return check_launch(__func__);
}

extern "C" int pcc_gather_rows(const float* x, const int32_t* idx, int64_t count, int d, int64_t n, float* out,
                               const float* g_in, float* g_out,
                               int device, void* stream) {
  PCC_ENTER(device);
  if (count == 0 || d == 0) return 0;
  PCC_REQUIRE(n > 0, "gather from an empty tensor");
  launch_dep(gather_rows_kernel, dim3((unsigned)cdiv(count * d, 256)), dim3(256), 0, (cudaStream_t)stream, x, idx, count, d, n, out, g_in, g_out);
  return check_launch(__func__);
}

extern "C" int pcc_segment_offsets(const int64_t* idx, int64_t n, int64_t B, int64_t* offsets, int device,
                                   void* stream) {
  PCC_ENTER(device);
  PCC_REQUIRE(n >= 0 && B >= 0, "negative size");
  PCC_REQUIRE(B + 1 <= (int64_t)PCC_SCAN_SINGLE_MAX, "too many segments for the single-block scan");
  cudaStream_t st = (cudaStream_t)stream;
  PCC_CUDA(cudaMemsetAsync(offsets, 0, (size_t)(B + 1) * sizeof(int64_t), st));
  int64_t blocks = cdiv(n, 1024);
  if (blocks < 1) blocks = 1;
  if (blocks > 592) blocks = 592;
  launch_dep(segment_offsets_fused_kernel, dim3((unsigned)blocks), dim3(1024), 0, st, idx, n, B, offsets);
  return check_launch(__func__);
}

extern "C" int pcc_index_max(const int64_t* idx, int64_t n, int64_t* out_max, int device, void* stream) {
  PCC_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  PCC_K(set_i64_kernel)<<<1, 1, 0, st>>>((long long*)out_max, -1);
  if (n > 0) {
    int blocks = (int)(cdiv(n, 256) < 592 ? cdiv(n, 256) : 592);
    PCC_K(index_max_kernel)<<<blocks, 256, 0, st>>>(idx, n, (long long*)out_max);
  }
  return check_launch(__func__);
}

extern "C" int pcc_segment_pool_fwd(const float* x, const int64_t* offsets, int64_t n, int64_t B, int64_t H,
                                    int pooling, float* pooled, int32_t* argmax, int device, void* stream) {
  PCC_ENTER(device);
  PCC_REQUIRE(pooling >= PCC_POOL_SUM && pooling <= PCC_POOL_ADD, "bad pooling id");
  PCC_REQUIRE(pooling != PCC_POOL_MAX || argmax != nullptr, "argmax buffer required for max pooling");
  PCC_REQUIRE(n < (int64_t)0x7fffffff, "row count exceeds int32 argmax range");
  if (B == 0 || H == 0) return 0;
  dim3 grid((unsigned)B, (unsigned)cdiv(H, 32));
  PCC_K(segment_pool_fwd_kernel)<<<grid, 256, 0, (cudaStream_t)stream>>>(x, offsets, H, pooling, pooled, argmax);
  return check_launch(__func__);
}

extern "C" int pcc_segment_pool_bwd(const float* dpooled, const int64_t* offsets, const int32_t* argmax, int64_t n,
                                    int64_t B, int64_t H, int pooling, float* dx, int device, void* stream) {
  PCC_ENTER(device);
  PCC_REQUIRE(pooling >= PCC_POOL_SUM && pooling <= PCC_POOL_ADD, "bad pooling id");
  PCC_REQUIRE(pooling != PCC_POOL_MAX || argmax != nullptr, "argmax buffer required for max pooling");
  if (n == 0 || H == 0) return 0;
  PCC_K(segment_pool_bwd_kernel)<<<(unsigned)cdiv(n, 32), 256, 0, (cudaStream_t)stream>>>(dpooled, offsets, argmax, n, B, H,
                                                                                     pooling, dx);
  return check_launch(__func__);
}

// ====================================================================== device-side collate helpers
//   reference: /root/reference/utils/data.py:651-663 (_collate_sparse: idx = cat(full((n_i,), i))) and :1228-1261
//   (_graph_collate: membership = cat(full((n_i,), i)), edges_i + node offset).  The per-sample python loops become
//   two launches over the concatenated arrays.
namespace pcc {

// idx[i] = segment of row i (last s with offsets[s] <= i); offsets[B+1]
__global__ void expand_segments_kernel(const int64_t* __restrict__ offsets, int64_t B, int64_t n, int64_t* __restrict__ idx) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t lo = 0, hi = B;
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(offsets + mid) <= i) lo = mid; else hi = mid;
  }
  idx[i] = lo;
}

// edges[2,E] hold per-graph LOCAL node ids, graphs back to back (edge_offsets[G+1]); out = edges + node_offsets[graph]
__global__ void offset_edges_kernel(const int64_t* __restrict__ edges, int64_t E, const int64_t* __restrict__ edge_offsets,
                                    const int64_t* __restrict__ node_offsets, int64_t G, int64_t* __restrict__ out) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t lo = 0, hi = G;
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(edge_offsets + mid) <= e) lo = mid; else hi = mid;
  }
  const int64_t off = __ldg(node_offsets + lo);
  out[e] = edges[e] + off;
  out[E + e] = edges[E + e] + off;
}

}  // namespace pcc

extern "C" int pcc_expand_segments(const int64_t* offsets, int64_t B, int64_t n, int64_t* idx, int device, void* stream) {
  PCC_ENTER(device);
  PCC_REQUIRE(B >= 1 || n == 0, "rows without a segment");
  if (n == 0) return 0;
  PCC_K(expand_segments_kernel)<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(offsets, B, n, idx);
  return check_launch(__func__);
}

extern "C" int pcc_offset_edges(const int64_t* edges, int64_t E, const int64_t* edge_offsets, const int64_t* node_offsets,
                                int64_t G, int64_t* out, int device, void* stream) {
  PCC_ENTER(device);
  PCC_REQUIRE(G >= 1 || E == 0, "edges without a graph");
  if (E == 0) return 0;
  PCC_K(offset_edges_kernel)<<<(unsigned)cdiv(E, 256), 256, 0, (cudaStream_t)stream>>>(edges, E, edge_offsets, node_offsets,
                                                                                  G, out);
  return check_launch(__func__);
}
