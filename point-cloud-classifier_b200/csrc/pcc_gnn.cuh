// Shared definitions of the fused bf16 GraphNet path (pcc_gnn*.cu): the neighbour stage of
// /root/reference/models/graph_net.py:65-104 (GraphConv -> act -> BatchNorm1d, twice; fc1 -> act -> bn3 ->
// global_mean_pool) with tcgen05 / TMEM GEMMs.
//
// Data flow (C = hidden_dim = 128, M nodes, train mode):
//   conv1  (CUDA cores, K = 2F <= 16) : x[M,F] -> agg1[M,F] fp32, z1[M,C] fp32, partial sums of act(z1), act(z1)^2
//   bn fin : partials -> scale s = gamma*invstd, shift t = beta - mean*s (+ running statistics)
//   apply  : h1 = bf16( act(z1)*s1 + t1 )                                     [M,C] bf16 (stays in L2: 67 MB)
//   conv2  (ONE tcgen05 kernel)       : CSR gather-reduce of h1 rows -> bf16 A image [agg2 | h1] in shared memory
//                                       -> [128 x 2C] x [2C x C] MMA -> z2 fp32 + BatchNorm partial sums; agg2 kept (bf16)
//   apply  : h2 = bf16( act(z2)*s2 + t2 )
//   fc1    (ONE tcgen05 kernel)       : z3 = h2 Wfc1^T + b; a3 = act(z3); BatchNorm partial sums + per-graph sums of a3
//                                       (mean pooling commutes with the BatchNorm affine); z3 never reaches HBM
// and the mirrored backward (pcc_gnn_bwd.cu).  Activations between kernels: fp32 pre-activations z (needed by act'
// and the BatchNorm backward), bf16 normalised activations h (the GEMM / gather operands).
#pragma once
#include "pcc_common.cuh"
#include "pcc_tc.cuh"

namespace pcc {
namespace gnn {
using namespace tc;

constexpr int kC = 128;              // hidden width of the fused path
constexpr int kFc = 256;             // fc1 width (hard-coded in the reference, graph_net.py:61)
constexpr int kTile = 128;           // nodes per tile (= UMMA M)
constexpr uint32_t kSlab = kTile * 128u;   // one 64-column slab of a 128-row SW128 image: 16 KB

// Activation math of the bf16 path: ONE MUFU op per element (the kernels are issue bound, libm tanhf / erff cost ~30
// instructions each and dominated the first version of every epilogue):
//   tanh: tanh.approx.f32 (abs error ~5e-4, below the bf16 rounding of the stored activation);
//   gelu: 0.5 z (1 + tanh(sqrt(2/pi) (z + 0.044715 z^3))) with tanh.approx — within 5e-4 of the reference's exact-erf
//         nn.GELU() (graph_net.py:43); the fp32 mode of the module keeps libm (pcc_common.cuh).
// Forward and backward use the same functions, so act' is consistent with the act the forward applied.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int ACT>
__device__ __forceinline__ float actf(float z) {
  if (ACT == PCC_ACT_RELU) return fmaxf(z, 0.f);
  if (ACT == PCC_ACT_TANH) return tanh_fast(z);
  if (ACT == PCC_ACT_GELU) return 0.5f * z * (1.f + tanh_fast(0.7978845608028654f * z * fmaf(0.044715f, z * z, 1.f)));
  return z;
}
// act'(z), given a = act(z) where that is cheaper
template <int ACT>
__device__ __forceinline__ float actg(float z, float a) {
  if (ACT == PCC_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (ACT == PCC_ACT_TANH) return 1.f - a * a;
  if (ACT == PCC_ACT_GELU) {
    const float z2 = z * z;
    const float t = tanh_fast(0.7978845608028654f * z * fmaf(0.044715f, z2, 1.f));
    return 0.5f * (1.f + t) + 0.5f * z * (1.f - t * t) * (0.7978845608028654f * fmaf(0.134145f, z2, 1.f));
  }
  return 1.f;
}

// Sum over the 32 lanes of a warp of 32 per-lane values, transposed: afterwards v[0] of lane l holds the sum over
// all lanes of their v[l] (31 shuffles instead of 160).
template <int HALF>
__device__ __forceinline__ void tr_step(float* v, int lane) {
  const bool up = (lane & HALF) != 0;
#pragma unroll
  for (int i = 0; i < HALF; ++i) {
    const float send = up ? v[i] : v[i + HALF];
    const float keep = up ? v[i + HALF] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, HALF);
  }
}
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
  tr_step<16>(v, lane);
  tr_step<8>(v, lane);
  tr_step<4>(v, lane);
  tr_step<2>(v, lane);
  tr_step<1>(v, lane);
  return v[0];
}

// 16 values per lane: afterwards lanes l and l ^ 16 hold the sum over all 32 lanes of their v[l & 15]
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
  tr_step<8>(v, lane);
  tr_step<4>(v, lane);
  tr_step<2>(v, lane);
  tr_step<1>(v, lane);
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// byte offset of the 16-byte chunk holding columns [col0, col0+8) of row r in a 128-row SW128 image
__device__ __forceinline__ uint32_t img_chunk_off(int r, int col0) {
  return (uint32_t)(col0 >> 6) * kSlab + (uint32_t)r * 128u + ((uint32_t)(((col0 & 63) >> 3) ^ (r & 7)) << 4);
}

__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) { return make_float2(bf16_lo(u), bf16_hi(u)); }

// bounded mbarrier wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000ll) __trap();
  }
}

struct GnnGraph {
  const int64_t* rowptr;   // [M+1] CSR by target (forward) or by source (backward)
  const int32_t* col;      // [E] neighbour node per CSR slot
  const float* w;          // [E] edge weight per CSR slot, or null
  int mean;                // forward: divide the aggregate by the in-degree
};

}  // namespace gnn
}  // namespace pcc
