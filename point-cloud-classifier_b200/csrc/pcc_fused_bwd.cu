// Fused DeepSets phi backward on tcgen05 / TMEM (sm_100a): autograd of
// /root/reference/models/deep_sets.py:89-106 in two kernels.
//
//  K1 "chain" (persistent, 128-point tiles): recompute the forward per tile (activations
//     never stored by the forward), build dZ of the final Linear from the pooled gradient
//     (+ argmax rows for max pooling), then walk the layers backwards:
//        dZ_l = dH_l * act'(z_l)            (epilogue: TMEM x TMEM -> bf16 blob in smem)
//        dH_{l-1} = dZ_l W_l (+ dH_l for a ResidualBlock: MMA accumulates onto dH_l in TMEM)
//     The operand blobs needed by the weight gradients (h_{l-1}, dZ_l) are pushed to a
//     staging area with cp.async.bulk stores straight from the smem operand images.
//     (TMEM holds exactly one H x H fp32 matrix, so the weight gradients of several
//     layers cannot stay resident next to the chain's accumulators)
//  K2 "wgrad": dW_l = sum_tiles dZ_l^T h_{l-1} with both operands read as MN-major views of
//     the staged blobs, fp32 accumulation in TMEM across all tiles of a CTA, one partial
//     per CTA, column sums (db) on the epilogue warps; a small kernel reduces the partials.
#include "pcc_fused.cuh"

namespace pcc {

constexpr int kRingB = 5;  // weight slabs in flight in the chain kernel

struct BwdParams {
  const float* x;
  const int64_t* offsets;
  int64_t n, B, num_tiles;
  int d, L, pooling, res_mask;
  const uint8_t* wpack;
  uint32_t w_off[kMaxLayers], wt_off[kMaxLayers];
  const float* bias[kMaxLayers];
  const float* dpooled;
  const int32_t* argmax;
  uint8_t* stage_h[kMaxLayers];  // h_l blobs,  l = 0..L-2, [num_tiles][H/8][128][8] bf16
  uint8_t* stage_g[kMaxLayers];  // dZ_l blobs, l = 0..L-1
  float* part_w[kMaxLayers];     // per-CTA partial dW_l [grid][H][K_l]  (K_0 = 16)
  float* part_b[kMaxLayers];     // per-CTA partial db_l [grid][H]
};

struct BwdSmem {
  uint32_t bufG, bufH, bufX, ring, bias, bars, total;
};
__host__ __device__ inline BwdSmem bwd_smem(int H, int L) {
  BwdSmem s;
  uint32_t o = 0;
  s.bufG = o; o += kTileM * H * 2;
  s.bufH = o; o += kTileM * H * 2;
  s.bufX = o; o += kTileM * kK0 * 2;
  s.ring = o; o += kRingB * (uint32_t)(64 * H);
  s.bias = o; o += (uint32_t)L * H * 4;
  s.bars = o; o += 256;
  s.total = o;
  return s;
}

// ====================================================================== K1: chain
template <int H, int ACT>
__global__ void __launch_bounds__(kThreads, 1) phi_bwd_chain_kernel(const BwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const BwdSmem lay = bwd_smem(H, p.L);
  uint8_t* bufG = smem + lay.bufG;
  uint8_t* bufH = smem + lay.bufH;
  uint8_t* bufX = smem + lay.bufX;
  uint8_t* ring = smem + lay.ring;
  float* biasS = reinterpret_cast<float*>(smem + lay.bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kRingB;
  uint64_t* a_ready = bars + 2 * kRingB;
  uint64_t* acc_ready = bars + 2 * kRingB + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRingB + 2);

  constexpr uint32_t SLAB = 64 * H;
  constexpr uint32_t A_LBO = kTileM * 16;
  constexpr uint32_t W_LBO = H * 16;
  constexpr uint32_t BLOB = kTileM * H * 2;
  constexpr uint32_t ACC_A = 0, ACC_B = 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;
  const bool recompute_z0 = (L >= 3);  // z_0 is overwritten by z_1 during the forward sweep

  for (int i = threadIdx.x; i < L * H; i += kThreads) biasS[i] = __ldg(p.bias[i / H] + (i % H));
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRingB; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(a_ready, 128);
    mbar_init(acc_ready, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ===================== producer: weight slabs in consumption order
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      auto push = [&](const uint8_t* src, uint32_t bytes) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], bytes);
        bulk_g2s(ring + stage * SLAB, src, bytes, &full[stage]);
        if (++stage == kRingB) { stage = 0; phase ^= 1; }
      };
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int l = 0; l <= L - 2; ++l) {
          if (l == 0) push(p.wpack + p.w_off[0], (kK0 / 8) * W_LBO);
          else for (int s = 0; s < H / 32; ++s) push(p.wpack + p.w_off[l] + (size_t)s * SLAB, SLAB);
        }
        for (int l = L - 1; l >= 1; --l) {
          for (int s = 0; s < H / 32; ++s) push(p.wpack + p.wt_off[l] + (size_t)s * SLAB, SLAB);
          if (l == 1 && recompute_z0) push(p.wpack + p.w_off[0], (kK0 / 8) * W_LBO);
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc_bf16(128, H, 0, 0);
      uint32_t stage = 0, phase = 0, a_phase = 0;
      const uint32_t g_base = smem_u32(bufG), h_base = smem_u32(bufH), x_base = smem_u32(bufX), r_base = smem_u32(ring);
      // one GEMM: D[acc] (+)= act[128 x K] * slab-streamed blob^T, K = nslab * (2 or 1 K-steps)
      auto gemm = [&](uint32_t act_base, uint32_t acc_col, int nslab, int ksteps, bool accumulate_first) {
        for (int s = 0; s < nslab; ++s) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t w_slab = r_base + stage * SLAB;
          for (int ks = 0; ks < ksteps; ++ks) {
            const int kglob = s * 2 + ks;
            umma_bf16(tmem + acc_col, make_smem_desc(act_base + kglob * 2 * A_LBO, A_LBO, 128),
                      make_smem_desc(w_slab + ks * 2 * W_LBO, W_LBO, 128), IDESC, (kglob > 0) || accumulate_first);
          }
          umma_commit(&empty[stage]);
          if (++stage == kRingB) { stage = 0; phase ^= 1; }
        }
      };
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int l = 0; l <= L - 2; ++l) {  // forward sweep: z_l -> accA
          mbar_wait(a_ready, a_phase); a_phase ^= 1; tc_fence_after();
          if (l == 0) gemm(x_base, ACC_A, 1, kK0 / 16, false);
          else gemm(h_base, ACC_A, H / 32, 2, false);
          umma_commit(acc_ready);
        }
        for (int l = L - 1; l >= 1; --l) {  // backward sweep: dH_{l-1} -> accB
          mbar_wait(a_ready, a_phase); a_phase ^= 1; tc_fence_after();
          const bool res = (p.res_mask >> l) & 1;
          gemm(g_base, ACC_B, H / 32, 2, res);
          if (l == 1 && recompute_z0) gemm(x_base, ACC_A, 1, kK0 / 16, false);
          umma_commit(acc_ready);
        }
      }
    }
  } else {
    // ===================== epilogue warps 0-3
    const int r = warp * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t acc_phase = 0;
    const int d = p.d;
    float xr[kK0];
    auto load_x = [&](int64_t tile) {
      const int64_t row = tile * kTileM + r;
#pragma unroll
      for (int j = 0; j < kK0; ++j) xr[j] = (j < d && row < p.n && tile < p.num_tiles) ? __ldg(p.x + row * d + j) : 0.f;
    };
    // buffer hand-over with the bulk stores issued by thread 0
    auto acquire = [&]() {  // every earlier bulk store has finished reading its smem source
      if (r == 0) bulk_wait_read0();
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    auto store_blob = [&](uint8_t* gdst, const uint8_t* sbuf) {
      fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (r == 0) { bulk_s2g(gdst, sbuf, BLOB); bulk_commit(); }
    };
    auto wait_acc = [&]() { mbar_wait(acc_ready, acc_phase); acc_phase ^= 1; tc_fence_after(); };
    auto arrive_a = [&]() { tc_fence_before(); fence_proxy_async(); mbar_arrive(a_ready); };

    load_x(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int64_t r0 = tile * kTileM;
      const int64_t row = r0 + r;
      {
        uint4 c0 = make_uint4(pack_bf16x2(xr[0], xr[1]), pack_bf16x2(xr[2], xr[3]), pack_bf16x2(xr[4], xr[5]),
                              pack_bf16x2(xr[6], xr[7]));
        uint4 c1 = make_uint4(pack_bf16x2(xr[8], xr[9]), pack_bf16x2(xr[10], xr[11]), pack_bf16x2(xr[12], xr[13]),
                              pack_bf16x2(xr[14], xr[15]));
        *reinterpret_cast<uint4*>(bufX + r * 16) = c0;
        *reinterpret_cast<uint4*>(bufX + A_LBO + r * 16) = c1;
      }
      arrive_a();
      load_x(tile + gridDim.x);
      // set of this thread's row (for the pooled-gradient scatter)
      int64_t myset = -1;
      float scale = 0.f;
      if (row < p.n) {
        int64_t lo = 0, hi = p.B;
        while (lo < hi) {
          int64_t mid = (lo + hi) >> 1;
          if (__ldg(p.offsets + mid + 1) <= row) lo = mid + 1; else hi = mid;
        }
        if (lo < p.B && __ldg(p.offsets + lo) <= row) {
          myset = lo;
          const float cnt = (float)(__ldg(p.offsets + lo + 1) - __ldg(p.offsets + lo));
          scale = p.pooling == PCC_POOL_SUM ? rsqrtf(cnt) : (p.pooling == PCC_POOL_MEAN ? 1.f / cnt : 1.f);
        }
      }

      // ---- forward sweep epilogues: h_l = [h_{l-1} +] act(z_l + b_l) -> bufH, staged
      for (int l = 0; l <= L - 2; ++l) {
        wait_acc();
        acquire();
        const bool res = (p.res_mask >> l) & 1;
        const float* bl = biasS + l * H;
#pragma unroll 1
        for (int c = 0; c < H / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(lane_base + ACC_A + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint8_t* dst = bufH + (uint32_t)(c * 4 + q) * A_LBO + r * 16;
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = act_t<ACT>(__uint_as_float(v[q * 8 + j]) + bl[c * 32 + q * 8 + j]);
            if (res) {
              const uint4 old = *reinterpret_cast<const uint4*>(dst);
              o[0] += bf16_lo(old.x); o[1] += bf16_hi(old.x); o[2] += bf16_lo(old.y); o[3] += bf16_hi(old.y);
              o[4] += bf16_lo(old.z); o[5] += bf16_hi(old.z); o[6] += bf16_lo(old.w); o[7] += bf16_hi(old.w);
            }
            *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                        pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          }
        }
        store_blob(p.stage_h[l] + (size_t)tile * BLOB, bufH);
        if (l < L - 2) arrive_a();
      }

      // ---- dZ of the final Linear from the pooled gradient (autograd of deep_sets.py:96-106)
      acquire();
      {
        const float* g = p.dpooled + (myset >= 0 ? myset : 0) * H;
        const int32_t* am = p.argmax ? p.argmax + (myset >= 0 ? myset : 0) * H : nullptr;
#pragma unroll 1
        for (int kc = 0; kc < H / 8; ++kc) {
          float o[8];
          if (myset >= 0) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + kc * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + kc * 8 + 4));
            o[0] = g0.x; o[1] = g0.y; o[2] = g0.z; o[3] = g0.w; o[4] = g1.x; o[5] = g1.y; o[6] = g1.z; o[7] = g1.w;
            if (p.pooling == PCC_POOL_MAX) {
              const int4 a0 = __ldg(reinterpret_cast<const int4*>(am + kc * 8));
              const int4 a1 = __ldg(reinterpret_cast<const int4*>(am + kc * 8 + 4));
              const int rr = (int)row;
              o[0] = a0.x == rr ? o[0] : 0.f; o[1] = a0.y == rr ? o[1] : 0.f; o[2] = a0.z == rr ? o[2] : 0.f;
              o[3] = a0.w == rr ? o[3] : 0.f; o[4] = a1.x == rr ? o[4] : 0.f; o[5] = a1.y == rr ? o[5] : 0.f;
              o[6] = a1.z == rr ? o[6] : 0.f; o[7] = a1.w == rr ? o[7] : 0.f;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] *= scale;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = 0.f;
          }
          *reinterpret_cast<uint4*>(bufG + (uint32_t)kc * A_LBO + r * 16) =
              make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        }
      }
      store_blob(p.stage_g[L - 1] + (size_t)tile * BLOB, bufG);
      arrive_a();

      // ---- backward sweep epilogues: dZ_l = dH_l * act'(z_l + b_l) -> bufG, staged
      for (int l = L - 2; l >= 0; --l) {
        wait_acc();
        acquire();
        const float* bl = biasS + l * H;
#pragma unroll 1
        for (int c = 0; c < H / 32; ++c) {
          uint32_t z[32], g[32];
          tmem_ld32(lane_base + ACC_A + c * 32, z);
          tmem_ld32(lane_base + ACC_B + c * 32, g);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = __uint_as_float(g[q * 8 + j]) * act_grad_t<ACT>(__uint_as_float(z[q * 8 + j]) + bl[c * 32 + q * 8 + j]);
            *reinterpret_cast<uint4*>(bufG + (uint32_t)(c * 4 + q) * A_LBO + r * 16) =
                make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          }
        }
        store_blob(p.stage_g[l] + (size_t)tile * BLOB, bufG);
        if (l >= 1) arrive_a();
      }
    }
    if (r == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem);
}

// ====================================================================== K2: wgrad
constexpr int kSlots = 3;

template <int H>
__global__ void __launch_bounds__(kThreads, 1) phi_wgrad_kernel(const BwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr uint32_t BLOB = kTileM * H * 2;
  constexpr uint32_t XBLOB = kTileM * kK0 * 2;
  constexpr int HALVES = H / 128;
  constexpr int CH = H / 8;          // 16-byte feature chunks per blob row
  constexpr int CPW = CH / 4;        // chunks per epilogue warp for the db column sums
  uint8_t* slots = smem;                                   // kSlots blobs
  uint8_t* bufX = smem + kSlots * BLOB;                    // 2 x-blobs (tile parity)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSlots * BLOB + 2 * XBLOB);
  uint64_t* full = bars;               // [kSlots]
  uint64_t* empty = bars + kSlots;     // [kSlots]   count 1 (MMA commit) + 128 (epilogue readers)
  uint64_t* x_full = bars + 2 * kSlots;      // [2] count 128
  uint64_t* x_empty = bars + 2 * kSlots + 2; // [2] count 1
  uint64_t* acc_ready = bars + 2 * kSlots + 4;
  uint64_t* acc_free = bars + 2 * kSlots + 5;  // count 128
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kSlots + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 129); }
    for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 128); mbar_init(&x_empty[i], 1); }
    mbar_init(acc_ready, 1);
    mbar_init(acc_free, 128);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      auto push = [&](const uint8_t* src) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], BLOB);
        bulk_g2s(slots + stage * BLOB, src, BLOB, &full[stage]);
        if (++stage == kSlots) { stage = 0; phase ^= 1; }
      };
      for (int l = 0; l < L; ++l)
        for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
          push(p.stage_g[l] + (size_t)tile * BLOB);
          if (l >= 1) push(p.stage_h[l - 1] + (size_t)tile * BLOB);
        }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, xpar = 0, xphase = 0 /* bit i = phase of x buffer i */, free_phase = 0;
      const uint32_t s_base = smem_u32(slots), x_base = smem_u32(bufX);
      for (int l = 0; l < L; ++l) {
        const int Np = (l == 0) ? kK0 : H;
        const uint32_t idesc = make_idesc_bf16(128, Np, 1, 1);
        if (l > 0) { mbar_wait(acc_free, free_phase); free_phase ^= 1; tc_fence_after(); }
        bool first = true;
        for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
          const uint32_t g_stage = stage;
          mbar_wait(&full[stage], phase);
          if (++stage == kSlots) { stage = 0; phase ^= 1; }
          uint32_t in_addr, in_stage = 0;
          if (l >= 1) {
            in_stage = stage;
            mbar_wait(&full[stage], phase);
            if (++stage == kSlots) { stage = 0; phase ^= 1; }
            in_addr = s_base + in_stage * BLOB;
          } else {
            mbar_wait(&x_full[xpar], (xphase >> xpar) & 1);
            xphase ^= 1u << xpar;
            in_addr = x_base + xpar * XBLOB;
          }
          tc_fence_after();
          const uint32_t g_addr = s_base + g_stage * BLOB;
#pragma unroll
          for (int o = 0; o < HALVES; ++o)
            for (int ks = 0; ks < kTileM / 16; ++ks)
              umma_bf16(tmem + o * Np, make_smem_desc(g_addr + o * (16 * kTileM * 16) + ks * 256, 128, kTileM * 16),
                        make_smem_desc(in_addr + ks * 256, 128, kTileM * 16), idesc, !(first && ks == 0));
          first = false;
          umma_commit(&empty[g_stage]);
          if (l >= 1) umma_commit(&empty[in_stage]);
          else { umma_commit(&x_empty[xpar]); xpar ^= 1; }
        }
        umma_commit(acc_ready);
      }
    }
  } else {
    // ===================== epilogue warps: db column sums, x blobs for layer 0, TMEM flush
    const int r = warp * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t stage = 0, phase = 0, xpar = 0, xphase = 0, acc_phase = 0;
    const int d = p.d;
    for (int l = 0; l < L; ++l) {
      const int Np = (l == 0) ? kK0 : H;
      float dbacc[CPW][8];
#pragma unroll
      for (int i = 0; i < CPW; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dbacc[i][j] = 0.f;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        if (l == 0) {  // x tile -> [2][128][8] bf16 blob
          mbar_wait(&x_empty[xpar], ((xphase >> xpar) & 1) ^ 1);
          xphase ^= 1u << xpar;
          const int64_t row = tile * kTileM + r;
          float xr[kK0];
#pragma unroll
          for (int j = 0; j < kK0; ++j) xr[j] = (j < d && row < p.n) ? __ldg(p.x + row * d + j) : 0.f;
          uint8_t* xb = bufX + xpar * XBLOB;
          *reinterpret_cast<uint4*>(xb + r * 16) = make_uint4(pack_bf16x2(xr[0], xr[1]), pack_bf16x2(xr[2], xr[3]),
                                                              pack_bf16x2(xr[4], xr[5]), pack_bf16x2(xr[6], xr[7]));
          *reinterpret_cast<uint4*>(xb + kTileM * 16 + r * 16) =
              make_uint4(pack_bf16x2(xr[8], xr[9]), pack_bf16x2(xr[10], xr[11]), pack_bf16x2(xr[12], xr[13]),
                         pack_bf16x2(xr[14], xr[15]));
          fence_proxy_async();
          mbar_arrive(&x_full[xpar]);
          xpar ^= 1;
        }
        // column sums of the dZ blob: warp w owns chunks w, w+4, ...; lane owns rows lane, lane+32, ...
        mbar_wait(&full[stage], phase);
        const uint8_t* gb = slots + stage * BLOB;
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
          const int c = warp + 4 * i;
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const uint4 u = *reinterpret_cast<const uint4*>(gb + (uint32_t)c * (kTileM * 16) + (lane + 32 * rr) * 16);
            dbacc[i][0] += bf16_lo(u.x); dbacc[i][1] += bf16_hi(u.x); dbacc[i][2] += bf16_lo(u.y); dbacc[i][3] += bf16_hi(u.y);
            dbacc[i][4] += bf16_lo(u.z); dbacc[i][5] += bf16_hi(u.z); dbacc[i][6] += bf16_lo(u.w); dbacc[i][7] += bf16_hi(u.w);
          }
        }
        mbar_arrive(&empty[stage]);
        if (++stage == kSlots) { stage = 0; phase ^= 1; }
        if (l >= 1) {  // the input blob slot is only read by the tensor core; wait for THIS use of the slot
          // to be filled before releasing it, otherwise the arrival could land in the previous phase
          mbar_wait(&full[stage], phase);
          mbar_arrive(&empty[stage]);
          if (++stage == kSlots) { stage = 0; phase ^= 1; }
        }
      }
      // ---- db partial: reduce over lanes, lane 0 writes
      float* pb = p.part_b[l] + (size_t)blockIdx.x * H;
#pragma unroll
      for (int i = 0; i < CPW; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float s = warp_sum(dbacc[i][j]);
          if (lane == 0) pb[(warp + 4 * i) * 8 + j] = s;
        }
      // ---- dW partial: TMEM -> global
      mbar_wait(acc_ready, acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      float* pw = p.part_w[l] + (size_t)blockIdx.x * H * Np;
#pragma unroll 1
      for (int o = 0; o < HALVES; ++o) {
#pragma unroll 1
        for (int c = 0; c < Np / 32 || (c == 0 && Np < 32); ++c) {
          uint32_t v[32];
          tmem_ld32(lane_base + o * Np + c * 32, v);
          tmem_wait_ld();
          float* dst = pw + (size_t)(o * 128 + r) * Np + c * 32;
          const int nq = (Np < 32) ? Np / 4 : 8;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q < nq)
              *reinterpret_cast<float4*>(dst + q * 4) =
                  make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]), __uint_as_float(v[q * 4 + 2]),
                              __uint_as_float(v[q * 4 + 3]));
        }
      }
      tc_fence_before();
      mbar_arrive(acc_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem);
}

// sum the per-CTA partials: dW_l[H, K] (K = real in-features), db_l[H]
__global__ void wgrad_reduce_kernel(const float* __restrict__ part_w, const float* __restrict__ part_b, int grid_ctas,
                                    int H, int Kp, int K, float* __restrict__ dw, float* __restrict__ db) {
  const int total = H * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total + H; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    if (i < total) {
      const int row = i / K, col = i % K;
      for (int c = 0; c < grid_ctas; ++c) s += __ldg(part_w + ((size_t)c * H + row) * Kp + col);
      dw[i] = s;
    } else {
      const int f = i - total;
      for (int c = 0; c < grid_ctas; ++c) s += __ldg(part_b + (size_t)c * H + f);
      db[f] = s;
    }
  }
}

__global__ void zero_f32_kernel_b(float* p, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

// ------------------------------------------------------------------ host side
struct BwdWs {
  uint32_t w_off[kMaxLayers], wt_off[kMaxLayers];
  int64_t stage_h[kMaxLayers], stage_g[kMaxLayers], part_w[kMaxLayers], part_b[kMaxLayers];
  int64_t total;
  int grid;
};
constexpr int kGridCap = 160;  // upper bound on persistent CTAs used for sizing the partial buffers
static BwdWs bwd_ws(const pcc_phi_desc* d, int64_t n, int sms) {
  BwdWs w{};
  const int H = d->hidden, L = d->n_layers;
  const int64_t tiles = cdiv(n, kTileM);
  if (sms > kGridCap) sms = kGridCap;
  w.grid = (int)(tiles < sms ? tiles : sms);
  if (w.grid < 1) w.grid = 1;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t at = o; o = (o + bytes + 255) / 256 * 256; return at; };
  for (int l = 0; l < L; ++l) w.w_off[l] = (uint32_t)take((int64_t)((l == 0) ? kK0 : H) * H * 2);
  for (int l = 1; l < L; ++l) w.wt_off[l] = (uint32_t)take((int64_t)H * H * 2);
  const int64_t blob = (int64_t)kTileM * H * 2;
  for (int l = 0; l <= L - 2; ++l) w.stage_h[l] = take(tiles * blob);
  for (int l = 0; l < L; ++l) w.stage_g[l] = take(tiles * blob);
  for (int l = 0; l < L; ++l) {
    w.part_w[l] = take((int64_t)kGridCap * H * ((l == 0) ? kK0 : H) * 4);
    w.part_b[l] = take((int64_t)kGridCap * H * 4);
  }
  w.total = o;
  return w;
}

int64_t phi_bwd_workspace_bytes(const pcc_phi_desc* d, int64_t n) { return bwd_ws(d, n, kGridCap).total; }

template <int H, int ACT>
static int launch_chain(const BwdParams& p, int grid, cudaStream_t st) {
  const BwdSmem lay = bwd_smem(H, p.L);
  auto kern = phi_bwd_chain_kernel<H, ACT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total);
  if (e != cudaSuccess) return fail("pcc_deepsets_phi_pool_bwd", cudaGetErrorString(e));
  {
    ProfScope prof(1, st);
    PCC_K(kern)<<<grid, kThreads, lay.total, st>>>(p);
  }
  return 0;
}
template <int H>
static int launch_wgrad(const BwdParams& p, int grid, cudaStream_t st) {
  const int smem_bytes = kSlots * kTileM * H * 2 + 2 * kTileM * kK0 * 2 + 256;
  auto kern = phi_wgrad_kernel<H>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return fail("pcc_deepsets_phi_pool_bwd", cudaGetErrorString(e));
  {
    ProfScope prof(2, st);
    PCC_K(kern)<<<grid, kThreads, smem_bytes, st>>>(p);
  }
  return 0;
}

}  // namespace pcc

using namespace pcc;

extern "C" int pcc_deepsets_phi_pool_bwd(const pcc_phi_desc* d, const float* x, const int64_t* offsets, int64_t n,
                                         int64_t B, const float* dpooled, const int32_t* argmax, float* const* dw,
                                         float* const* db, void* ws, int device, void* stream) {
  PCC_ENTER(device);
  if (check_phi_desc(d, __func__) != 0) return -1;
  PCC_REQUIRE(d->pooling != PCC_POOL_MAX || argmax != nullptr, "argmax required for max pooling");
  cudaStream_t st = (cudaStream_t)stream;
  const int H = d->hidden, L = d->n_layers;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const BwdWs wl = bwd_ws(d, n, sms);
  uint8_t* wsb = (uint8_t*)ws;
  const int64_t tiles = cdiv(n, kTileM);
  if (tiles == 0) {
    for (int l = 0; l < L; ++l) {
      const int64_t cnt = (int64_t)H * (l == 0 ? d->input_dim : H);
      PCC_K(zero_f32_kernel_b)<<<(unsigned)cdiv(cnt, 256), 256, 0, st>>>(dw[l], cnt);
      PCC_K(zero_f32_kernel_b)<<<(unsigned)cdiv(H, 256), 256, 0, st>>>(db[l], H);
    }
    return check_launch(__func__);
  }

  PackParams pk{};
  for (int l = 0; l < L; ++l) { pk.w[l] = d->w[l]; pk.w_off[l] = wl.w_off[l]; pk.wt_off[l] = wl.wt_off[l]; }
  pk.wpack = wsb; pk.d = d->input_dim; pk.H = H; pk.L = L;
  PCC_K(pack_weights_kernel)<<<dim3(32, L, 2), 256, 0, st>>>(pk);

  BwdParams p{};
  p.x = x; p.offsets = offsets; p.n = n; p.B = B; p.num_tiles = tiles;
  p.d = d->input_dim; p.L = L; p.pooling = d->pooling; p.res_mask = d->residual_mask;
  p.wpack = wsb; p.dpooled = dpooled; p.argmax = argmax;
  for (int l = 0; l < L; ++l) {
    p.w_off[l] = wl.w_off[l]; p.wt_off[l] = wl.wt_off[l]; p.bias[l] = d->b[l];
    p.stage_g[l] = wsb + wl.stage_g[l];
    if (l <= L - 2) p.stage_h[l] = wsb + wl.stage_h[l];
    p.part_w[l] = (float*)(wsb + wl.part_w[l]);
    p.part_b[l] = (float*)(wsb + wl.part_b[l]);
  }
  int rc = 0;
#define PCC_DISPATCH(HH)                                                                 \
  switch (d->act) {                                                                      \
    case PCC_ACT_RELU: rc = launch_chain<HH, PCC_ACT_RELU>(p, wl.grid, st); break;       \
    case PCC_ACT_GELU: rc = launch_chain<HH, PCC_ACT_GELU>(p, wl.grid, st); break;       \
    default: rc = launch_chain<HH, PCC_ACT_SILU>(p, wl.grid, st); break;                 \
  }
  if (H == 256) { PCC_DISPATCH(256) } else { PCC_DISPATCH(128) }
#undef PCC_DISPATCH
  if (rc != 0) return rc;
  rc = (H == 256) ? launch_wgrad<256>(p, wl.grid, st) : launch_wgrad<128>(p, wl.grid, st);
  if (rc != 0) return rc;
  for (int l = 0; l < L; ++l) {
    const int K = (l == 0) ? d->input_dim : H, Kp = (l == 0) ? kK0 : H;
    PCC_K(wgrad_reduce_kernel)<<<(unsigned)cdiv((int64_t)H * K + H, 256), 256, 0, st>>>(p.part_w[l], p.part_b[l], wl.grid, H,
                                                                                Kp, K, dw[l], db[l]);
  }
  return check_launch(__func__);
}
