// Diagnostics: one-CTA tcgen05 GEMMs on exactly representable integer data, covering the
// three operand descriptor conventions the fused kernels rely on (see pcc_tc.cuh).
//   mode 0: D[i][j] = sum_k A[i][k] * B[j][k]      A, B K-major blobs          (forward)
//   mode 1: D[i][j] = sum_k A[i][k] * W[k][j]      B = MN-major view of W blob (dgrad)
//   mode 2: D[i][j] = sum_p G[p][i] * Hh[p][j]     both MN-major views         (wgrad)
//   mode 3: as mode 0 with K = 128 and SWIZZLE_128B images
//   mode 4: as mode 2 with SWIZZLE_128B images
#include "pcc_common.cuh"
#include "pcc_tc.cuh"

namespace pcc {
using namespace tc;

__device__ __host__ inline float st_a(int i, int k) { return (float)((i * 3 + k * 5) % 7 - 3); }
__device__ __host__ inline float st_b(int j, int k) { return (float)((j * 2 + k) % 5 - 2); }

// blob [C/8][128][8]: element (row, col)
__device__ inline void blob_store(__nv_bfloat16* blob, int row, int col, float v) {
  blob[((col >> 3) * 128 + row) * 8 + (col & 7)] = __float2bfloat16(v);
}

__global__ void __launch_bounds__(128, 1) selftest_umma_kernel(int mode, float* out) {
  extern __shared__ __align__(1024) uint8_t dyn_smem[];
  __nv_bfloat16* blobA = reinterpret_cast<__nv_bfloat16*>(dyn_smem);
  __nv_bfloat16* blobB = blobA + 128 * 128;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  const int N = 64;
  // mode 0: A[128 x 64] rows=i cols=k ; B[64(rows j, padded to 128 rows) x 64]
  // mode 1: A[128 x 128] rows=i cols=k ; W blob rows=k(128) cols=j(64)
  // mode 2: G blob rows=p(128) cols=i(128) ; Hh blob rows=p(128) cols=j(64)
  const bool swz = mode >= 3;
  for (int e = t; e < 128 * 128; e += 128) {
    const int row = e / 128, col = e % 128;
    float a = 0.f, b = 0.f;
    if (mode == 3) { a = st_a(row, col); b = row < N ? st_b(row, col) : 0.f; }
    if (mode == 4) { a = st_a(col, row); b = col < N ? st_b(col, row) : 0.f; }
    if (swz) {
      blobA[sw128_off(row, col, 128) / 2] = __float2bfloat16(a);
      blobB[sw128_off(row, col, 128) / 2] = __float2bfloat16(b);
      continue;
    }
    if (mode == 0) { a = col < 64 ? st_a(row, col) : 0.f; b = (col < 64 && row < N) ? st_b(row, col) : 0.f; }
    if (mode == 1) { a = st_a(row, col); b = col < N ? st_b(col, row) : 0.f; }   // W[k=row][j=col] = st_b(j,k)
    if (mode == 2) { a = st_a(col, row); b = col < N ? st_b(col, row) : 0.f; }   // G[p][i] = st_a(i,p); Hh[p][j] = st_b(j,p)
    blob_store(blobA, row, col, a);
    blob_store(blobB, row, col, b);
  }
  if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<64>(&tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    const uint32_t a0 = smem_u32(blobA), b0 = smem_u32(blobB);
    if (mode == 0) {
      const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem, make_smem_desc(a0 + ks * 2 * 2048, 2048, 128), make_smem_desc(b0 + ks * 2 * 2048, 2048, 128),
                  idesc, ks > 0);
    } else if (mode == 1) {
      const uint32_t idesc = make_idesc_bf16(128, N, 0, 1);
      for (int ks = 0; ks < 8; ++ks)  // K step = 16 rows of the W blob = 256 B
        umma_bf16(tmem, make_smem_desc(a0 + ks * 2 * 2048, 2048, 128), make_smem_desc(b0 + ks * 256, 128, 2048), idesc,
                  ks > 0);
    } else if (mode == 2) {
      const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
      for (int ks = 0; ks < 8; ++ks)
        umma_bf16(tmem, make_smem_desc(a0 + ks * 256, 128, 2048), make_smem_desc(b0 + ks * 256, 128, 2048), idesc,
                  ks > 0);
    } else if (mode == 3) {
      const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      for (int ks = 0; ks < 8; ++ks) {  // slab = ks / 4 (64 K values each), 32 B per K step inside a slab
        const uint32_t off = (ks >> 2) * (128 * 128) + (ks & 3) * 32;
        umma_bf16(tmem, make_smem_desc_sw128_k(a0 + off), make_smem_desc_sw128_k(b0 + off), idesc, ks > 0);
      }
    } else {
      const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
      for (int ks = 0; ks < 8; ++ks)  // K step = 16 rows = two 1024 B row groups
        umma_bf16(tmem, make_smem_desc_sw128_mn(a0 + ks * 2048, 128 * 128), make_smem_desc_sw128_mn(b0 + ks * 2048, 128 * 128),
                  idesc, ks > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < N / 32; ++c) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out[t * N + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}
}  // namespace pcc

using namespace pcc;

extern "C" int pcc_selftest_umma(int mode, float* out, int device, void* stream) {
  PCC_ENTER(device);
  PCC_REQUIRE(mode >= 0 && mode <= 4, "mode must be 0..4");
  const int smem = 2 * 128 * 128 * 2;
  PCC_CUDA(cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  selftest_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mode, out);
  return check_launch(__func__);
}

// ---- FP32 FMA peak (the roofline denominator of the kNN kernel, SURVEY.md section 8d: "must be measured on the box"):
// 8 independent FMA chains per thread, 2 x iters x 8 flops per thread, grid = 8 blocks of 256 threads per SM.
namespace pcc {
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  for (int i = 0; i < iters; ++i) {
    x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
    x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
  }
  out[blockIdx.x * 256 + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace pcc

extern "C" int pcc_selftest_fp32_peak(float* out, int blocks, int iters, int device, void* stream) {
  PCC_ENTER(device);
  fp32_peak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters, 0.999f, 0.001f);
  return check_launch(__func__);
}
