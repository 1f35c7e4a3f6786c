// Tile problems of the small fp32 GEMM kernel (pcc_head.cu): C[i][j] = sum_kk A(i,kk) B(j,kk) with elementwise
// transforms folded into the operand loads.  Shared with the fused phi kernels, which use it for the [B, H]-sized
// GEMMs around the pooled activations (sum / mean pooling commuted with the final Linear).
#pragma once
#include "pcc_common.cuh"

namespace pcc {

// operand element (i, kk): i = output index of the tile problem, kk = contraction index
struct HeadOperand {
  const float* p;
  const float* q;     // act' argument (mode 2)
  int64_t si, sk;     // element strides of p along i / kk (one of them is 1)
  int64_t qi, qk;
  int mode;           // 0: p   1: act(p)   2: p * act'(q)
};
// C[i][j] = sum_kk A(i,kk) B(j,kk) (+ bias[j]);  colsum[i] = sum_kk A(i,kk)
struct HeadTileProb {
  HeadOperand A, B;
  int I, J, KK;
  float* C;
  int64_t ldc;
  const float* bias;
  float* colsum;
  const float* row_scale;   // [I] or null:  C[i][j] = row_scale[i] * acc + bias_scale[i] * bias[j]
  const float* bias_scale;  // [I] or null
  const float* colsum_w;    // [KK] or null: colsum[i] = sum_kk colsum_w[kk] * A(i,kk)
  int tiles_j, tiles;
};
struct HeadTileParams {
  HeadTileProb prob[2];
  int act;
};

// launches `tiles` CTAs of the tile kernel (act: activation id used by operand modes 1 / 2)
void launch_head_tiles(const HeadTileParams& hp, int tiles, cudaStream_t st);
void head_set_tiles(HeadTileProb& p);

}  // namespace pcc
