// Fused DeepSets phi-MLP + ragged pooling on tcgen05 / TMEM (sm_100a).
//   reference: /root/reference/models/deep_sets.py:89 (phi), :91-106 (split + pool loop)
//   and their autograd.  bf16 operands, fp32 accumulation in TMEM.
//
// Forward kernel (persistent, one CTA per SM, 128-point tiles):
//   x tile -> bf16 operand blob in smem -> [tcgen05.mma -> TMEM -> epilogue(bias, act,
//   residual) -> bf16 blob in smem] per hidden layer -> final Linear computed TRANSPOSED
//   (M = output features, N = points) so that each epilogue thread owns one feature and
//   pools over the points of its TMEM lane without any cross-thread traffic -> partial
//   sums / packed (value,row) maxima combined across tiles with atomics in a [B,H]
//   accumulator.  Per-point activations never leave the SM.
//   Weights: pre-packed bf16 blobs (see pcc_tc.cuh) streamed from L2 through an
//   mbarrier ring with cp.async.bulk (TMA engine).
// Warp roles: warps 0-3 epilogue (TMEM lane quarter = warp id), warp 4 bulk-copy
// producer, warp 5 TMEM allocator + MMA issuer (one elected thread).
#include "pcc_fused.cuh"

namespace pcc {

struct SmemLayout {
  uint32_t bufA, bufX, ring, bias, bars, total;
};
__host__ __device__ inline SmemLayout smem_layout(int H, int L) {
  SmemLayout s;
  uint32_t o = 0;
  s.bufA = o; o += kTileM * H * 2;
  s.bufX = o; o += kTileM * kK0 * 2;
  s.ring = o; o += kRing * (uint32_t)(64 * H);  // slab = 4 K-chunks * H rows * 16 B
  s.bias = o; o += (uint32_t)L * H * 4;
  s.bars = o; o += 256;
  s.total = o;
  return s;
}

// ------------------------------------------------------------------ forward kernel
template <int H, int ACT>
__global__ void __launch_bounds__(kThreads, 1) phi_pool_fwd_kernel(const PhiParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const SmemLayout lay = smem_layout(H, p.L);
  uint8_t* bufA = smem + lay.bufA;
  uint8_t* bufX = smem + lay.bufX;
  uint8_t* ring = smem + lay.ring;
  float* biasS = reinterpret_cast<float*>(smem + lay.bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bars);
  uint64_t* full = bars;               // [kRing]
  uint64_t* empty = bars + kRing;      // [kRing]
  uint64_t* a_ready = bars + 2 * kRing;
  uint64_t* acc_ready = bars + 2 * kRing + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRing + 2);
  int* seg_first = reinterpret_cast<int*>(bars + 2 * kRing + 3);  // [2], by tile parity

  constexpr uint32_t SLAB = 64 * H;            // bytes
  constexpr uint32_t A_LBO = kTileM * 16;      // K-chunk stride of a 128-row blob
  constexpr uint32_t W_LBO = H * 16;           // K-chunk stride of an H-row blob
  constexpr int HALVES = H / 128;              // M halves of the transposed final layer
  constexpr uint32_t ACC_T = 256;              // TMEM column base of the transposed accumulators

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;

  for (int i = threadIdx.x; i < L * H; i += kThreads) biasS[i] = __ldg(p.bias[i / H] + (i % H));
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
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
    // ===================== producer: stream weight slabs through the ring
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int l = 0; l < L; ++l) {
          const int nslab = (l == 0) ? 1 : H / 32;
          const uint32_t bytes = (l == 0) ? (kK0 / 8) * W_LBO : SLAB;
          for (int s = 0; s < nslab; ++s) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], bytes);
            bulk_g2s(ring + stage * SLAB, p.wpack + p.w_off[l] + (size_t)s * SLAB, bytes, &full[stage]);
            if (++stage == kRing) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC_N = make_idesc_bf16(128, H, 0, 0);    // points x features
      constexpr uint32_t IDESC_T = make_idesc_bf16(128, 128, 0, 0);  // features(128) x points
      uint32_t stage = 0, phase = 0, a_phase = 0;
      const uint32_t a_base = smem_u32(bufA), x_base = smem_u32(bufX), r_base = smem_u32(ring);
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int l = 0; l < L; ++l) {
          mbar_wait(a_ready, a_phase);
          a_phase ^= 1;
          tc_fence_after();
          const bool last = (l == L - 1);
          const int nslab = (l == 0) ? 1 : H / 32;
          const int ksteps = (l == 0) ? kK0 / 16 : 2;  // UMMA K steps per slab
          const uint32_t act_base = (l == 0) ? x_base : a_base;
          for (int s = 0; s < nslab; ++s) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t w_slab = r_base + stage * SLAB;
            for (int ks = 0; ks < ksteps; ++ks) {
              const int kglob = s * 2 + ks;  // K step index inside the layer
              const uint64_t act_desc = make_smem_desc(act_base + kglob * 2 * A_LBO, A_LBO, 128);
              if (!last) {
                const uint64_t w_desc = make_smem_desc(w_slab + ks * 2 * W_LBO, W_LBO, 128);
                umma_bf16(tmem, act_desc, w_desc, IDESC_N, kglob > 0);
              } else {
#pragma unroll
                for (int h = 0; h < HALVES; ++h) {
                  const uint64_t w_desc = make_smem_desc(w_slab + ks * 2 * W_LBO + h * 128 * 16, W_LBO, 128);
                  umma_bf16(tmem + ACC_T + h * 128, w_desc, act_desc, IDESC_T, kglob > 0);
                }
              }
            }
            umma_commit(&empty[stage]);
            if (++stage == kRing) { stage = 0; phase ^= 1; }
          }
          umma_commit(acc_ready);
        }
      }
    }
  } else {
    // ===================== epilogue warps 0-3: thread = TMEM lane = tile row (or feature)
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
    load_x(blockIdx.x);
    int par = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, par ^= 1) {
      const int64_t r0 = tile * kTileM;
      // ---- stage the x tile as the layer-0 operand blob [2][128][8] bf16
      {
        uint4 c0 = make_uint4(pack_bf16x2(xr[0], xr[1]), pack_bf16x2(xr[2], xr[3]), pack_bf16x2(xr[4], xr[5]),
                              pack_bf16x2(xr[6], xr[7]));
        uint4 c1 = make_uint4(pack_bf16x2(xr[8], xr[9]), pack_bf16x2(xr[10], xr[11]), pack_bf16x2(xr[12], xr[13]),
                              pack_bf16x2(xr[14], xr[15]));
        *reinterpret_cast<uint4*>(bufX + r * 16) = c0;
        *reinterpret_cast<uint4*>(bufX + A_LBO + r * 16) = c1;
      }
      if (r == 0) {  // first set intersecting this tile (binary search over offsets)
        int64_t lo = 0, hi = p.B;
        while (lo < hi) {
          int64_t mid = (lo + hi) >> 1;
          if (__ldg(p.offsets + mid + 1) <= r0) lo = mid + 1; else hi = mid;
        }
        seg_first[par] = (int)lo;
      }
      fence_proxy_async();
      mbar_arrive(a_ready);
      load_x(tile + gridDim.x);  // prefetch the next tile's rows into registers

      // ---- hidden layers: TMEM -> bias/act/residual -> bf16 blob (in place)
      for (int l = 0; l < L - 1; ++l) {
        mbar_wait(acc_ready, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        const bool res = (p.res_mask >> l) & 1;
        const float* bl = biasS + l * H;
#pragma unroll 1
        for (int c = 0; c < H / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(lane_base + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint8_t* dst = bufA + (uint32_t)(c * 4 + q) * A_LBO + r * 16;
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
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(a_ready);
      }

      // ---- final layer (transposed): thread = feature, TMEM columns = the tile's points
      mbar_wait(acc_ready, acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      asm volatile("bar.sync 1, 128;" ::: "memory");  // seg_first visible to all epilogue threads
      const int b_first = seg_first[par];
      const int64_t tile_end = (r0 + kTileM < p.n) ? r0 + kTileM : p.n;
#pragma unroll 1
      for (int h = 0; h < HALVES; ++h) {
        const int f = h * 128 + r;
        int64_t b = b_first;
        int64_t seg_lo = 0, seg_hi = 0;
        if (b < p.B) { seg_lo = __ldg(p.offsets + b); seg_hi = __ldg(p.offsets + b + 1); }
        float acc = (p.pooling == PCC_POOL_MAX) ? -INFINITY : 0.f;
        int arg = -1;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(lane_base + ACC_T + h * 128 + c * 32, v);
          tmem_wait_ld();
          const int64_t col0 = r0 + c * 32;
          while (b < p.B && seg_lo < tile_end && seg_lo < col0 + 32) {
            // columns of this chunk that belong to set b: [lo, hi)
            const int lo = (int)((seg_lo > col0 ? seg_lo : col0) - col0);
            const int64_t hi64 = (seg_hi < col0 + 32 ? seg_hi : col0 + 32) - col0;
            const int hi = (int)hi64;
            if (p.pooling == PCC_POOL_MAX) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float val = __uint_as_float(v[j]);
                if (j >= lo && j < hi && (val > acc || arg < 0)) { acc = val; arg = c * 32 + j; }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc += (j >= lo && j < hi) ? __uint_as_float(v[j]) : 0.f;
            }
            if (seg_hi > col0 + 32) break;  // set continues in the next chunk / tile
            // set b ends inside this chunk: flush and advance
            if (p.pooling == PCC_POOL_MAX) {
              if (arg >= 0) {
                unsigned long long key = ((unsigned long long)float_ordered(acc) << 32) |
                                         (unsigned long long)(0xFFFFFFFFu - (uint32_t)(r0 + arg));
                atomicMax(reinterpret_cast<unsigned long long*>(p.pool_acc) + b * H + f, key);
              }
              acc = -INFINITY; arg = -1;
            } else {
              atomicAdd(reinterpret_cast<float*>(p.pool_acc) + b * H + f, acc);
              acc = 0.f;
            }
            ++b;
            if (b < p.B) { seg_lo = seg_hi; seg_hi = __ldg(p.offsets + b + 1); }
          }
        }
        // partial of the set that continues past this tile
        if (b < p.B && seg_lo < tile_end) {
          if (p.pooling == PCC_POOL_MAX) {
            if (arg >= 0) {
              unsigned long long key = ((unsigned long long)float_ordered(acc) << 32) |
                                       (unsigned long long)(0xFFFFFFFFu - (uint32_t)(r0 + arg));
              atomicMax(reinterpret_cast<unsigned long long*>(p.pool_acc) + b * H + f, key);
            }
          } else {
            atomicAdd(reinterpret_cast<float*>(p.pool_acc) + b * H + f, acc);
          }
        }
      }
      tc_fence_before();
    }
  }

  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem);
}

// pool accumulator -> pooled[B,H] (+ argmax): adds the final bias after pooling
// (max(z+b) = max(z)+b, mean(z+b) = mean(z)+b, sum(z+b)/sqrt(n) = sum(z)/sqrt(n) + b*sqrt(n))
__global__ void pool_finalize_kernel(const void* __restrict__ pool_acc, const int64_t* __restrict__ offsets,
                                     const float* __restrict__ bias, int64_t B, int H, int pooling,
                                     float* __restrict__ pooled, int32_t* __restrict__ argmax) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int64_t b = i / H;
  const int f = (int)(i % H);
  const float n = (float)(offsets[b + 1] - offsets[b]);
  if (pooling == PCC_POOL_MAX) {
    const unsigned long long key = reinterpret_cast<const unsigned long long*>(pool_acc)[i];
    if (key == 0ull) {
      pooled[i] = 0.f;
      argmax[i] = -1;
    } else {
      pooled[i] = ordered_float((uint32_t)(key >> 32)) + bias[f];
      argmax[i] = (int32_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
    }
  } else {
    const float s = reinterpret_cast<const float*>(pool_acc)[i];
    if (n <= 0.f) pooled[i] = 0.f;
    else if (pooling == PCC_POOL_SUM) pooled[i] = s / sqrtf(n) + bias[f] * sqrtf(n);
    else pooled[i] = s / n + bias[f];
  }
}

// ------------------------------------------------------------------ host side
struct WsLayout {
  uint32_t w_off[kMaxLayers];
  int64_t wpack_bytes, pool_off, total;
};
static WsLayout ws_layout(const pcc_phi_desc* d, int64_t B) {
  WsLayout w{};
  int64_t o = 0;
  for (int l = 0; l < d->n_layers; ++l) {
    w.w_off[l] = (uint32_t)o;
    o += (int64_t)((l == 0) ? kK0 : d->hidden) * d->hidden * 2;
  }
  w.wpack_bytes = o;
  o = (o + 255) / 256 * 256;
  w.pool_off = o;
  o += B * d->hidden * 8;
  w.total = (o + 255) / 256 * 256;
  return w;
}

int check_phi_desc(const pcc_phi_desc* d, const char* where) {
  if (!d) return fail(where, "null descriptor");
  // the backward chain keeps z of the last hidden layer in TMEM and recomputes only z_0, which
  // covers phi = Linear(d,H) [+ one H x H hidden layer] + final Linear(H,H)
  if (d->n_layers < 2 || d->n_layers > 3) return fail(where, "fused path needs 1 or 2 hidden phi layers (+ final Linear)");
  if (d->input_dim < 1 || d->input_dim > kK0) return fail(where, "fused path needs input_dim <= 16");
  if (d->hidden != 128 && d->hidden != 256) return fail(where, "fused path needs hidden width 128 or 256");
  if (d->act != PCC_ACT_RELU && d->act != PCC_ACT_GELU && d->act != PCC_ACT_SILU)
    return fail(where, "fused path needs relu/gelu/silu");
  if (d->pooling != PCC_POOL_SUM && d->pooling != PCC_POOL_MEAN && d->pooling != PCC_POOL_MAX)
    return fail(where, "bad pooling id");
  if (d->residual_mask & 1) return fail(where, "layer 0 cannot be a residual block");
  if ((d->residual_mask >> (d->n_layers - 1)) & 1) return fail(where, "the final Linear cannot be a residual block");
  return 0;
}

template <int H, int ACT>
static int launch_fwd(const PhiParams& p, cudaStream_t st) {
  const SmemLayout lay = smem_layout(H, p.L);
  auto kern = phi_pool_fwd_kernel<H, ACT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total);
  if (e != cudaSuccess) return fail("pcc_deepsets_phi_pool_fwd", cudaGetErrorString(e));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(p.num_tiles < sms ? p.num_tiles : sms);
  {
    ProfScope prof(0, st);
    PCC_K(kern)<<<grid, kThreads, lay.total, st>>>(p);
  }
  return 0;
}

}  // namespace pcc

using namespace pcc;

extern "C" int pcc_phi_fused_supported(const pcc_phi_desc* d) { return check_phi_desc(d, __func__); }

extern "C" int64_t pcc_phi_fused_workspace_bytes(const pcc_phi_desc* d, int64_t n, int64_t B) {
  if (check_phi_desc(d, __func__) != 0) return -1;
  const int64_t fwd = ws_layout(d, B).total, bwd = phi_bwd_workspace_bytes(d, n);
  return fwd > bwd ? fwd : bwd;  // one query serves both directions
}

extern "C" int pcc_deepsets_phi_pool_fwd(const pcc_phi_desc* d, const float* x, const int64_t* offsets, int64_t n,
                                         int64_t B, float* pooled, int32_t* argmax, void* ws, int device,
                                         void* stream) {
  PCC_ENTER(device);
  if (check_phi_desc(d, __func__) != 0) return -1;
  PCC_REQUIRE(d->pooling != PCC_POOL_MAX || argmax != nullptr, "argmax buffer required for max pooling");
  PCC_REQUIRE(n < (int64_t)0x7fffffff, "row count exceeds int32 argmax range");
  cudaStream_t st = (cudaStream_t)stream;
  const int H = d->hidden, L = d->n_layers;
  const WsLayout wl = ws_layout(d, B);
  uint8_t* wsb = (uint8_t*)ws;

  PackParams pk{};
  for (int l = 0; l < L; ++l) { pk.w[l] = d->w[l]; pk.w_off[l] = wl.w_off[l]; }
  pk.wpack = wsb; pk.d = d->input_dim; pk.H = H; pk.L = L;
  PCC_K(pack_weights_kernel)<<<dim3(32, L), 256, 0, st>>>(pk);
  if (B * H > 0) PCC_K(zero_u64_kernel)<<<(unsigned)cdiv(B * H, 256), 256, 0, st>>>((unsigned long long*)(wsb + wl.pool_off), B * H);

  PhiParams p{};
  p.x = x; p.offsets = offsets; p.n = n; p.B = B; p.num_tiles = cdiv(n, kTileM);
  p.d = d->input_dim; p.L = L; p.pooling = d->pooling; p.res_mask = d->residual_mask;
  p.wpack = wsb;
  for (int l = 0; l < L; ++l) { p.w_off[l] = wl.w_off[l]; p.bias[l] = d->b[l]; }
  p.pool_acc = wsb + wl.pool_off;
  if (p.num_tiles > 0) {
    int rc = 0;
#define PCC_DISPATCH(HH)                                                              \
    switch (d->act) {                                                                 \
      case PCC_ACT_RELU: rc = launch_fwd<HH, PCC_ACT_RELU>(p, st); break;             \
      case PCC_ACT_GELU: rc = launch_fwd<HH, PCC_ACT_GELU>(p, st); break;             \
      default: rc = launch_fwd<HH, PCC_ACT_SILU>(p, st); break;                       \
    }
    if (H == 256) { PCC_DISPATCH(256) } else { PCC_DISPATCH(128) }
#undef PCC_DISPATCH
    if (rc != 0) return rc;
  }
  if (B * H > 0)
    PCC_K(pool_finalize_kernel)<<<(unsigned)cdiv(B * H, 256), 256, 0, st>>>(wsb + wl.pool_off, offsets, d->b[L - 1], B, H,
                                                                     d->pooling, pooled, argmax);
  return check_launch(__func__);
}

