// Fused rho head (set encoder): [Linear + act] x (0..3) + Linear, M = batch rows.
//   reference: /root/reference/models/deep_sets.py:112 (self.rho(pooled)) with the layer stack of :59-72
//   (no LayerNorm variant) and its autograd.
// The head is tiny (0.07 GFLOP per train step at the yaml shape) and pure launch latency when run layer by
// layer (13 launches); here the whole forward is one launch and the whole backward another.  Each CTA owns 8
// rows through ALL layers (rows are independent), activations live in shared memory, weights stream from L2
// with 128-bit loads (one warp per output unit in the forward, one thread per input column in the dgrad),
// weight gradients are combined across CTAs with vector atomics (red.global.add.v4.f32).
#include "pcc_common.cuh"

namespace pcc {

constexpr int kHeadRows = 4;   // rows per CTA (64 CTAs at B = 256)
constexpr int kHeadUnits = 4;  // output units a warp processes together (independent weight-load streams)
constexpr int kHeadMaxDim = 1024;
constexpr int kHeadMaxLayers = 4;

struct HeadParams {
  int L;                        // layers incl. the final Linear
  int dims[kHeadMaxLayers + 1]; // dims[0] = input width, dims[l+1] = output width of layer l
  int zoff[kHeadMaxLayers];     // column offset of hidden layer l inside zsave
  int zwidth;                   // sum of hidden widths
  int act;
  const float* w[kHeadMaxLayers];
  const float* b[kHeadMaxLayers];
  float* dw[kHeadMaxLayers];
  float* db[kHeadMaxLayers];
  const float* x;               // [M, dims[0]]
  float* y;                     // [M, dims[L]]
  float* zsave;                 // [M, zwidth] pre-activations of the hidden layers
  const float* dy;              // [M, dims[L]]
  float* dx;                    // [M, dims[0]] or null
  int64_t M;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(256) head_fwd_kernel(const HeadParams p) {
  extern __shared__ __align__(16) float hs[];
  float* bufA = hs;                              // [8][maxdim]
  float* bufB = hs + kHeadRows * kHeadMaxDim;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * kHeadRows;
  const int nrow = (int)((p.M - r0) < kHeadRows ? (p.M - r0) : kHeadRows);
  const int K0 = p.dims[0];
  for (int i = threadIdx.x; i < kHeadRows * K0; i += 256) {
    const int r = i / K0, k = i % K0;
    bufA[r * kHeadMaxDim + k] = (r < nrow) ? __ldg(p.x + (r0 + r) * K0 + k) : 0.f;
  }
  __syncthreads();
  float* in = bufA;
  float* out = bufB;
  for (int l = 0; l < p.L; ++l) {
    const int K = p.dims[l], N = p.dims[l + 1];
    const bool hidden = l < p.L - 1;
    for (int u0 = warp * kHeadUnits; u0 < N; u0 += 8 * kHeadUnits) {
      float acc[kHeadUnits][kHeadRows];
#pragma unroll
      for (int j = 0; j < kHeadUnits; ++j)
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r) acc[j][r] = 0.f;
      for (int k0 = lane * 4; k0 < K; k0 += 128) {
        float4 wv[kHeadUnits];
#pragma unroll
        for (int j = 0; j < kHeadUnits; ++j)
          wv[j] = (u0 + j < N) ? __ldg(reinterpret_cast<const float4*>(p.w[l] + (int64_t)(u0 + j) * K + k0))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(in + r * kHeadMaxDim + k0);
#pragma unroll
          for (int j = 0; j < kHeadUnits; ++j)
            acc[j][r] = fmaf(a.x, wv[j].x, fmaf(a.y, wv[j].y, fmaf(a.z, wv[j].z, fmaf(a.w, wv[j].w, acc[j][r]))));
        }
      }
#pragma unroll
      for (int j = 0; j < kHeadUnits; ++j)
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r) acc[j][r] = warp_sum(acc[j][r]);
      // lane (j * kHeadRows + r) finishes unit u0 + j of row r
      if (lane < kHeadUnits * kHeadRows) {
        const int j = lane / kHeadRows, r = lane % kHeadRows;
        const int u = u0 + j;
        float z = 0.f;
#pragma unroll
        for (int jj = 0; jj < kHeadUnits; ++jj)
#pragma unroll
          for (int rr = 0; rr < kHeadRows; ++rr) z = (jj == j && rr == r) ? acc[jj][rr] : z;
        if (u < N) {
          z += __ldg(p.b[l] + u);
          if (hidden) {
            if (r < nrow) p.zsave[(r0 + r) * p.zwidth + p.zoff[l] + u] = z;
            out[r * kHeadMaxDim + u] = act_fwd(p.act, z);
          } else if (r < nrow) {
            p.y[(r0 + r) * N + u] = z;
          }
        }
      }
    }
    __syncthreads();
    float* t = in; in = out; out = t;
  }
}

// backward: dy -> (dz_l, dW_l, db_l) for l = L-1 .. 0, dx
__global__ void __launch_bounds__(256) head_bwd_kernel(const HeadParams p) {
  extern __shared__ __align__(16) float hs[];
  float* ain = hs;                                   // [8][maxdim] input activations of the current layer
  float* dcur = hs + kHeadRows * kHeadMaxDim;        // [8][maxdim] gradient w.r.t. the current layer's output / dz
  float* dprev = hs + 2 * kHeadRows * kHeadMaxDim;   // [8][maxdim] gradient w.r.t. its input
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * kHeadRows;
  const int nrow = (int)((p.M - r0) < kHeadRows ? (p.M - r0) : kHeadRows);
  const int NO = p.dims[p.L];
  for (int i = threadIdx.x; i < kHeadRows * NO; i += 256) {
    const int r = i / NO, u = i % NO;
    dcur[r * kHeadMaxDim + u] = (r < nrow) ? __ldg(p.dy + (r0 + r) * NO + u) : 0.f;
  }
  for (int l = p.L - 1; l >= 0; --l) {
    const int K = p.dims[l], N = p.dims[l + 1];
    // input activations of layer l: x (l == 0) or act(z_{l-1}); dz_l = dcur * act'(z_l) for hidden layers
    for (int i = threadIdx.x; i < kHeadRows * K; i += 256) {
      const int r = i / K, k = i % K;
      float v = 0.f;
      if (r < nrow) v = (l == 0) ? __ldg(p.x + (r0 + r) * K + k)
                                 : act_fwd(p.act, __ldg(p.zsave + (r0 + r) * p.zwidth + p.zoff[l - 1] + k));
      ain[r * kHeadMaxDim + k] = v;
    }
    if (l < p.L - 1) {
      __syncthreads();  // dcur fully written by the previous dgrad
      for (int i = threadIdx.x; i < kHeadRows * N; i += 256) {
        const int r = i / N, u = i % N;
        const float z = (r < nrow) ? __ldg(p.zsave + (r0 + r) * p.zwidth + p.zoff[l] + u) : 0.f;
        dcur[r * kHeadMaxDim + u] *= act_grad(p.act, z);
      }
    }
    __syncthreads();
    // ---- db_l and dW_l: warp w owns units w, w+8, ...; lanes own 4 consecutive input columns
    for (int u = warp; u < N; u += 8) {
      float dz[kHeadRows];
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) { dz[r] = dcur[r * kHeadMaxDim + u]; s += dz[r]; }
      if (lane == 0) atomicAdd(p.db[l] + u, s);
      float* dwr = p.dw[l] + (int64_t)u * K;
      for (int k0 = lane * 4; k0 < K; k0 += 128) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(ain + r * kHeadMaxDim + k0);
          g.x = fmaf(dz[r], a.x, g.x); g.y = fmaf(dz[r], a.y, g.y); g.z = fmaf(dz[r], a.z, g.z); g.w = fmaf(dz[r], a.w, g.w);
        }
        red_add_v4(dwr + k0, g.x, g.y, g.z, g.w);
      }
    }
    // ---- dgrad: dprev[r, k] = sum_u dz[r, u] * W[u, k]   (thread per input column, coalesced over k)
    if (l > 0 || p.dx) {
      for (int k = threadIdx.x; k < K; k += 256) {
        float acc[kHeadRows];
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r) acc[r] = 0.f;
#pragma unroll 8
        for (int u = 0; u < N; ++u) {
          const float wv = __ldg(p.w[l] + (int64_t)u * K + k);
#pragma unroll
          for (int r = 0; r < kHeadRows; ++r) acc[r] = fmaf(dcur[r * kHeadMaxDim + u], wv, acc[r]);
        }
        if (l > 0) {
#pragma unroll
          for (int r = 0; r < kHeadRows; ++r) dprev[r * kHeadMaxDim + k] = acc[r];
        } else {
#pragma unroll
          for (int r = 0; r < kHeadRows; ++r)
            if (r < nrow) p.dx[(r0 + r) * K + k] = acc[r];
        }
      }
    }
    __syncthreads();
    float* t = dcur; dcur = dprev; dprev = t;
  }
}

__global__ void head_zero_kernel(HeadParams p) {
  const int l = blockIdx.y;
  if (l >= p.L) return;
  const int64_t nw = (int64_t)p.dims[l] * p.dims[l + 1], nb = p.dims[l + 1];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nw + nb; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < nw) p.dw[l][i] = 0.f; else p.db[l][i - nw] = 0.f;
  }
}

static int check_head(const pcc_head_desc* d, const char* where) {
  if (!d) return fail(where, "null descriptor");
  if (d->n_layers < 1 || d->n_layers > kHeadMaxLayers) return fail(where, "head needs 1..4 layers");
  for (int l = 0; l <= d->n_layers; ++l) {
    if (d->dims[l] < 1 || d->dims[l] > kHeadMaxDim) return fail(where, "head widths must be in [1,1024]");
    if (l < d->n_layers && d->dims[l] % 4 != 0) return fail(where, "head input widths must be multiples of 4");
  }
  if (d->act != PCC_ACT_RELU && d->act != PCC_ACT_GELU && d->act != PCC_ACT_SILU && d->act != PCC_ACT_TANH)
    return fail(where, "head activation must be relu/gelu/silu/tanh");
  return 0;
}

static HeadParams make_params(const pcc_head_desc* d, int64_t M) {
  HeadParams p{};
  p.L = d->n_layers; p.act = d->act; p.M = M;
  int off = 0;
  for (int l = 0; l <= d->n_layers; ++l) p.dims[l] = d->dims[l];
  for (int l = 0; l < d->n_layers; ++l) {
    p.w[l] = d->w[l]; p.b[l] = d->b[l];
    if (l < d->n_layers - 1) { p.zoff[l] = off; off += d->dims[l + 1]; }
  }
  p.zwidth = off;
  return p;
}

}  // namespace pcc

using namespace pcc;

extern "C" int pcc_mlp_head_supported(const pcc_head_desc* d) { return check_head(d, __func__); }

extern "C" int pcc_mlp_head_fwd(const pcc_head_desc* d, const float* x, float* y, float* zsave, int64_t M, int device,
                                void* stream) {
  PCC_ENTER(device);
  if (check_head(d, __func__) != 0) return -1;
  if (M == 0) return 0;
  HeadParams p = make_params(d, M);
  p.x = x; p.y = y; p.zsave = zsave;
  const int smem = 2 * kHeadRows * kHeadMaxDim * (int)sizeof(float);
  PCC_CUDA(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  PCC_K(head_fwd_kernel)<<<(unsigned)cdiv(M, kHeadRows), 256, smem, (cudaStream_t)stream>>>(p);
  return check_launch(__func__);
}

extern "C" int pcc_mlp_head_bwd(const pcc_head_desc* d, const float* x, const float* zsave, const float* dy, float* dx,
                                float* const* dw, float* const* db, int64_t M, int device, void* stream) {
  PCC_ENTER(device);
  if (check_head(d, __func__) != 0) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  HeadParams p = make_params(d, M);
  p.x = x; p.zsave = const_cast<float*>(zsave); p.dy = dy; p.dx = dx;
  for (int l = 0; l < d->n_layers; ++l) { p.dw[l] = dw[l]; p.db[l] = db[l]; }
  PCC_K(head_zero_kernel)<<<dim3(64, d->n_layers), 256, 0, st>>>(p);
  if (M > 0) {
    const int smem = 3 * kHeadRows * kHeadMaxDim * (int)sizeof(float);
    PCC_CUDA(cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PCC_K(head_bwd_kernel)<<<(unsigned)cdiv(M, kHeadRows), 256, smem, st>>>(p);
  }
  return check_launch(__func__);
}
