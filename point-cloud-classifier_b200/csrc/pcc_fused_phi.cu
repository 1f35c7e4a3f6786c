// Fused DeepSets phi-MLP + ragged pooling on tcgen05 / TMEM (sm_100a).
//   reference: /root/reference/models/deep_sets.py:89 (phi), :91-106 (split + pool loop)
//   and their autograd.  bf16 operands, fp32 accumulation in TMEM.
//
// Forward kernel (persistent, one CTA per SM, 128-point tiles):
//   x tile -> bf16 operand image in smem -> [tcgen05.mma -> TMEM -> epilogue(bias, act,
//   residual) -> bf16 SWIZZLE_128B image in smem] per hidden layer -> final Linear computed
//   TRANSPOSED (M = output features, N = points) so that each epilogue thread owns one feature
//   and pools over the points of its TMEM lane without any cross-thread traffic -> partial
//   sums / packed (value,row) maxima combined across tiles with atomics in a [B,H]
//   accumulator.  Per-point activations never leave the SM.
//   Weights: pre-packed bf16 images (pcc_fused.cuh) streamed from L2 through an mbarrier ring
//   with cp.async.bulk (TMA engine), one K=64 slab (H rows x 128 B) per slot.
// Warp roles: warps 0-7 epilogue (TMEM lane quarter = warp & 3; the two warps of a quarter split
// the accumulator columns / the two feature halves), warp 8 bulk-copy producer, warp 9 TMEM
// allocator + MMA issuer (one elected thread).
#include "pcc_fused.cuh"

namespace pcc {

constexpr int kRingF = 4;  // weight slabs in flight (forward)

struct SmemLayout {
  uint32_t bufA, ring, bufX, bias, bars, total;
};
__host__ __device__ inline SmemLayout smem_layout(int H, int L) {
  SmemLayout s;
  uint32_t o = 0;
  s.bufA = o; o += kTileM * H * 2;             // activation image, 1024-aligned slabs
  s.ring = o; o += kRingF * w_slab_bytes(H);   // weight slabs, 1024-aligned
  s.bufX = o; o += kTileM * kK0 * 2;           // layer-0 operand (un-swizzled)
  s.bias = o; o += (uint32_t)L * H * 4;
  s.bars = o; o += 256;
  s.total = o;
  return s;
}

// debug trace: CTA 0 only, role 0 = epilogue thread 0, role 1 = MMA thread; 2 x 4096 slots
__device__ __forceinline__ void trace_ev(long long* trace, int role, int& n, int id) {
  if (trace && blockIdx.x == 0 && n < 2047) {
    trace[role * 4096 + 2 * n] = id;
    trace[role * 4096 + 2 * n + 1] = clock64();
    ++n;
  }
}

__device__ __forceinline__ float max8(const uint32_t* v) {
  return fmaxf(fmaxf(fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]))),
               fmaxf(fmaxf(__uint_as_float(v[4]), __uint_as_float(v[5])), fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7]))));
}
__device__ __forceinline__ float sum8(const uint32_t* v) {
  return ((__uint_as_float(v[0]) + __uint_as_float(v[1])) + (__uint_as_float(v[2]) + __uint_as_float(v[3]))) +
         ((__uint_as_float(v[4]) + __uint_as_float(v[5])) + (__uint_as_float(v[6]) + __uint_as_float(v[7])));
}

// one 32-column accumulator chunk -> bias, activation, residual -> bf16 -> SW128 activation image
template <int ACT>
__device__ __forceinline__ void epi_store_chunk(const uint32_t (&v)[32], uint8_t* bufA, const float* bl, int r, int c,
                                                bool res) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint8_t* dst = bufA + act_chunk_off(r, c * 32 + q * 8);
    const float4 b0 = *reinterpret_cast<const float4*>(bl + c * 32 + q * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(bl + c * 32 + q * 8 + 4);
    float o[8];
    o[0] = act_t<ACT>(__uint_as_float(v[q * 8 + 0]) + b0.x); o[1] = act_t<ACT>(__uint_as_float(v[q * 8 + 1]) + b0.y);
    o[2] = act_t<ACT>(__uint_as_float(v[q * 8 + 2]) + b0.z); o[3] = act_t<ACT>(__uint_as_float(v[q * 8 + 3]) + b0.w);
    o[4] = act_t<ACT>(__uint_as_float(v[q * 8 + 4]) + b1.x); o[5] = act_t<ACT>(__uint_as_float(v[q * 8 + 5]) + b1.y);
    o[6] = act_t<ACT>(__uint_as_float(v[q * 8 + 6]) + b1.z); o[7] = act_t<ACT>(__uint_as_float(v[q * 8 + 7]) + b1.w);
    if (res) {
      const uint4 old = *reinterpret_cast<const uint4*>(dst);
      o[0] += bf16_lo(old.x); o[1] += bf16_hi(old.x); o[2] += bf16_lo(old.y); o[3] += bf16_hi(old.y);
      o[4] += bf16_lo(old.z); o[5] += bf16_hi(old.z); o[6] += bf16_lo(old.w); o[7] += bf16_hi(old.w);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
  }
}

// ------------------------------------------------------------------ forward kernel
template <int H, int ACT>
__global__ void __launch_bounds__(kThreads, 1) phi_pool_fwd_kernel(const PhiParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const SmemLayout lay = smem_layout(H, p.L);
  uint8_t* bufA = smem + lay.bufA;
  uint8_t* ring = smem + lay.ring;
  uint8_t* bufX = smem + lay.bufX;
  float* biasS = reinterpret_cast<float*>(smem + lay.bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bars);
  uint64_t* full = bars;                 // [kRingF]
  uint64_t* empty = bars + kRingF;       // [kRingF]
  uint64_t* x_ready = bars + 2 * kRingF;           // layer-0 operand staged (count: all epilogue threads)
  uint64_t* acc_ready = bars + 2 * kRingF + 1;      // one MMA phase (layer of a tile) complete
  uint64_t* slab_ready = bars + 2 * kRingF + 2;     // [4] 64-column slab of the activation image written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRingF + 6);

  constexpr uint32_t SLAB = w_slab_bytes(H);   // K = 64 slab of a weight image
  constexpr uint32_t X_LBO = kTileM * 16;      // un-swizzled layer-0 images: K-chunk strides
  constexpr uint32_t W0_LBO = H * 16;
  constexpr int HALVES = H / 128;              // M halves of the transposed final layer
  constexpr int NSLAB = H / 64;                // slabs per H x H layer
  constexpr int NCHUNK = H / 32;               // 32-column accumulator chunks of a hidden layer
  // TMEM: two 256-column accumulators used alternately by successive MMA phases, so the MMA of
  // layer l+1 can start on the first activation slabs while the epilogue still drains layer l

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;

  for (int i = threadIdx.x; i < L * H; i += kThreads) biasS[i] = __ldg(p.bias[i / H] + (i % H));
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRingF; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(x_ready, kEpiWarps);
    mbar_init(acc_ready, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&slab_ready[i], kEpiWarps);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kProdWarp) {
    // ===================== producer: stream weight slabs through the ring
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int l = 0; l < L; ++l) {
          const int nslab = (l == 0) ? 1 : NSLAB;
          const uint32_t bytes = (l == 0) ? (kK0 / 8) * W0_LBO : SLAB;
          for (int s = 0; s < nslab; ++s) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], bytes);
            bulk_g2s(ring + stage * SLAB, p.wpack + p.w_off[l] + (size_t)s * SLAB, bytes, &full[stage]);
            if (++stage == kRingF) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC_N = make_idesc_bf16(128, H, 0, 0);    // points x features
      constexpr uint32_t IDESC_T = make_idesc_bf16(128, 128, 0, 0);  // features(128) x points
      uint32_t stage = 0, phase = 0, x_phase = 0, sl_phase = 0, ph = 0;
      int tn = 0;
      const uint32_t a_base = smem_u32(bufA), x_base = smem_u32(bufX), r_base = smem_u32(ring);
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int l = 0; l < L; ++l, ++ph) {
          const uint32_t acc_col = tmem + (ph & 1) * 256;
          const bool last = (l == L - 1);
          trace_ev(p.trace, 1, tn, 100 + l);
          if (l == 0) {  // K = 16, un-swizzled images
            mbar_wait(x_ready, x_phase);
            x_phase ^= 1;
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            trace_ev(p.trace, 1, tn, 110 + l);
            umma_bf16(acc_col, make_smem_desc(x_base, X_LBO, 128), make_smem_desc(r_base + stage * SLAB, W0_LBO, 128),
                      IDESC_N, 0);
            umma_commit(&empty[stage]);
            if (++stage == kRingF) { stage = 0; phase ^= 1; }
          } else {
            for (int s = 0; s < NSLAB; ++s) {
              mbar_wait(&slab_ready[s], sl_phase);  // activation slab s written by the epilogue
              mbar_wait(&full[stage], phase);        // weight slab s landed
              tc_fence_after();
              if (s == 0) trace_ev(p.trace, 1, tn, 120 + l);
              if (s == NSLAB - 1) trace_ev(p.trace, 1, tn, 130 + l);
              const uint32_t w_slab = r_base + stage * SLAB;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t act_desc = make_smem_desc_sw128_k(a_base + s * kActSlab + ks * 32);
                const uint32_t acc = (s | ks) != 0;
                if (!last) {
                  umma_bf16(acc_col, act_desc, make_smem_desc_sw128_k(w_slab + ks * 32), IDESC_N, acc);
                } else {
#pragma unroll
                  for (int h = 0; h < HALVES; ++h)
                    umma_bf16(acc_col + h * 128, make_smem_desc_sw128_k(w_slab + h * (128 * 128) + ks * 32), act_desc,
                              IDESC_T, acc);
                }
              }
              umma_commit(&empty[stage]);
              if (++stage == kRingF) { stage = 0; phase ^= 1; }
            }
            sl_phase ^= 1;
          }
          umma_commit(acc_ready);
          trace_ev(p.trace, 1, tn, 140 + l);
        }
      }
    }
  } else {
    // ===================== epilogue warps 0-7
    const int quarter = warp & 3, grp = warp >> 2;
    const int r = quarter * 32 + lane;  // tile row (hidden layers) / feature inside a half (final layer)
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    uint32_t acc_phase = 0;
    const int d = p.d;
    float xr[kK0];
    auto load_x = [&](int64_t tile) {
      const int64_t row = tile * kTileM + r;
#pragma unroll
      for (int j = 0; j < kK0; ++j)
        xr[j] = (grp == 0 && j < d && row < p.n && tile < p.num_tiles) ? __ldg(p.x + row * d + j) : 0.f;
    };
    auto stage_x = [&]() {  // registers -> un-swizzled [2][128][8] layer-0 operand image (group 0 owns the rows)
      if (grp == 0) {
        *reinterpret_cast<uint4*>(bufX + r * 16) = make_uint4(pack_bf16x2(xr[0], xr[1]), pack_bf16x2(xr[2], xr[3]),
                                                              pack_bf16x2(xr[4], xr[5]), pack_bf16x2(xr[6], xr[7]));
        *reinterpret_cast<uint4*>(bufX + X_LBO + r * 16) = make_uint4(pack_bf16x2(xr[8], xr[9]), pack_bf16x2(xr[10], xr[11]),
                                                                      pack_bf16x2(xr[12], xr[13]), pack_bf16x2(xr[14], xr[15]));
      }
      fence_proxy_async();
      mbar_arrive_warp(x_ready);
    };
    load_x(blockIdx.x);
    int tn = 0;
    uint32_t ph = 0;
    const bool tr0 = (threadIdx.x == 0);
    if (blockIdx.x < p.num_tiles) {
      stage_x();
      load_x(blockIdx.x + gridDim.x);
    }
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int64_t r0 = tile * kTileM;
      if (tr0) trace_ev(p.trace, 0, tn, 0);
      const int b_first = __ldg(p.tile_first + tile);  // first set intersecting this tile (precomputed)

      // ---- hidden layers: TMEM -> bias/act/residual -> bf16 image (in place); group g takes the
      //      chunks g, g+2, ...; TMEM loads run one chunk ahead of the math; every finished 64-column
      //      slab is handed to the MMA warp at once so the next layer's MMAs overlap this epilogue
      for (int l = 0; l < L - 1; ++l, ++ph) {
        mbar_wait(acc_ready, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        if (tr0) trace_ev(p.trace, 0, tn, 10 + l);
        const bool res = (p.res_mask >> l) & 1;
        const float* bl = biasS + l * H;
        const uint32_t acc_base = lane_base + (ph & 1) * 256;
        uint32_t va[32], vb[32];
        tmem_ld32(acc_base + grp * 32, va);
#pragma unroll 1
        for (int c = grp; c < NCHUNK; c += 4) {
          tmem_wait_ld();
          if (c + 2 < NCHUNK) tmem_ld32(acc_base + (c + 2) * 32, vb);
          epi_store_chunk<ACT>(va, bufA, bl, r, c, res);
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive_warp(&slab_ready[c >> 1]);
          if (c + 2 < NCHUNK) {
            tmem_wait_ld();
            if (c + 4 < NCHUNK) tmem_ld32(acc_base + (c + 4) * 32, va);
            epi_store_chunk<ACT>(vb, bufA, bl, r, c + 2, res);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive_warp(&slab_ready[(c + 2) >> 1]);
          }
        }
        if (tr0) trace_ev(p.trace, 0, tn, 20 + l);
      }
      // ---- the next tile's layer-0 operand can be staged already (its MMA uses the other accumulator)
      if (tile + gridDim.x < p.num_tiles) {
        stage_x();
        load_x(tile + 2 * (int64_t)gridDim.x);
      }

      // ---- final layer (transposed): thread = feature, TMEM columns = the tile's points
      mbar_wait(acc_ready, acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      if (tr0) trace_ev(p.trace, 0, tn, 30);
      const uint32_t accT = lane_base + (ph & 1) * 256;
      ++ph;
      const int64_t tile_end = (r0 + kTileM < p.n) ? r0 + kTileM : p.n;
      if (grp < HALVES) {
        const int h = grp;
        const int f = h * 128 + r;
        const bool is_max = (p.pooling == PCC_POOL_MAX);
        int64_t b = b_first;
        int64_t seg_lo = 0, seg_hi = 0;
        if (b < p.B) { seg_lo = __ldg(p.offsets + b); seg_hi = __ldg(p.offsets + b + 1); }
        float acc = is_max ? -INFINITY : 0.f;
        int arg = -1;       // column of the running max inside this tile (first occurrence)
        auto flush = [&](int64_t set) {
          if (is_max) {
            if (arg >= 0) {
              unsigned long long key = ((unsigned long long)float_ordered(acc) << 32) |
                                       (unsigned long long)(0xFFFFFFFFu - (uint32_t)(r0 + arg));
              atomicMax(reinterpret_cast<unsigned long long*>(p.pool_acc) + set * H + f, key);
            }
            acc = -INFINITY; arg = -1;
          } else {
            atomicAdd(reinterpret_cast<float*>(p.pool_acc) + set * H + f, acc);
            acc = 0.f;
          }
        };
        uint32_t va[32], vb[32];
        tmem_ld32(accT + h * 128, va);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_wait_ld();
          uint32_t (&v)[32] = (c & 1) ? vb : va;
          if (c + 1 < 4) tmem_ld32(accT + h * 128 + (c + 1) * 32, (c & 1) ? va : vb);
          const int64_t col0 = r0 + c * 32;
          while (b < p.B && seg_lo < tile_end && seg_lo < col0 + 32) {
            const int lo = (int)((seg_lo > col0 ? seg_lo : col0) - col0);
            const int hi = (int)((seg_hi < col0 + 32 ? seg_hi : col0 + 32) - col0);
            if (lo == 0 && hi == 32) {  // whole chunk inside the set: tree reduction, no predicates
              if (is_max) {
                const float m = fmaxf(fmaxf(max8(v), max8(v + 8)), fmaxf(max8(v + 16), max8(v + 24)));
                if (m > acc || arg < 0) {  // the chunk improves the maximum: locate its first occurrence
                  int j0 = 31;
#pragma unroll
                  for (int j = 30; j >= 0; --j) j0 = (__uint_as_float(v[j]) == m) ? j : j0;
                  acc = m;
                  arg = c * 32 + j0;
                }
              } else {
                acc += (sum8(v) + sum8(v + 8)) + (sum8(v + 16) + sum8(v + 24));
              }
            } else if (is_max) {
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
            flush(b);                        // set b ends inside this chunk
            ++b;
            if (b < p.B) { seg_lo = seg_hi; seg_hi = __ldg(p.offsets + b + 1); }
          }
        }
        if (b < p.B && seg_lo < tile_end) flush(b);  // partial of the set that continues past this tile
      }
      tc_fence_before();
      if (tr0) trace_ev(p.trace, 0, tn, 40);
    }
  }

  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem);
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
  int64_t pool_off, tile_first_off, total;
};
static WsLayout ws_layout(const pcc_phi_desc* d, int64_t n, int64_t B) {
  WsLayout w{};
  int64_t o = 0;
  w.pool_off = o;
  o += B * d->hidden * 8;
  o = (o + 255) / 256 * 256;
  w.tile_first_off = o;
  o += cdiv(n, kTileM) * 4;
  w.total = (o + 255) / 256 * 256;
  return w;
}

// ONE launch for everything the forward kernel needs prepared: blockIdx.y < 2L packs weight image
// (layer y>>1, transposed if y&1); y == 2L zeroes the pool accumulator; y == 2L+1 computes tile_first
__global__ void fwd_prep_kernel(PackParams pk, unsigned long long* pool, int64_t pool_count, const int64_t* offsets,
                                int64_t B, int64_t num_tiles, int32_t* tile_first) {
  const int y = blockIdx.y;
  if (y < 2 * pk.L) {
    const int l = y >> 1;
    const bool tr = y & 1;
    if (tr && l == 0) return;
    const int K = (l == 0) ? pk.d : pk.H;
    const int Kp = (l == 0) ? kK0 : pk.H;
    const int total = (Kp / 8) * pk.H;
    uint8_t* dst = pk.wpack + (tr ? pk.wt_off[l] : pk.w_off[l]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int kc = i / pk.H, row = i % pk.H;
      uint32_t q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k0 = kc * 8 + 2 * j;
        float a, b;
        if (!tr) {
          a = (k0 < K) ? __ldg(pk.w[l] + (int64_t)row * K + k0) : 0.f;
          b = (k0 + 1 < K) ? __ldg(pk.w[l] + (int64_t)row * K + k0 + 1) : 0.f;
        } else {
          a = __ldg(pk.w[l] + (int64_t)k0 * K + row);
          b = __ldg(pk.w[l] + (int64_t)(k0 + 1) * K + row);
        }
        q[j] = pack_bf16x2(a, b);
      }
      const uint32_t off = (l == 0) ? (uint32_t)i * 16u
                                    : (uint32_t)(kc >> 3) * w_slab_bytes(pk.H) + (uint32_t)row * 128u +
                                          ((uint32_t)((kc & 7) ^ (row & 7)) << 4);
      *reinterpret_cast<uint4*>(dst + off) = make_uint4(q[0], q[1], q[2], q[3]);
    }
  } else if (y == 2 * pk.L) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pool_count; i += (int64_t)gridDim.x * blockDim.x)
      pool[i] = 0ull;
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < num_tiles; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r0 = i * kTileM;
      int64_t lo = 0, hi = B;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid + 1) <= r0) lo = mid + 1; else hi = mid;
      }
      tile_first[i] = (int32_t)lo;
    }
  }
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

static void* g_trace_buf = nullptr;  // set through pcc_debug_set_trace (diagnostics only)
void* debug_trace_buffer() { return g_trace_buf; }

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

extern "C" int pcc_debug_set_trace(void* device_buf_2x4096_i64) {
  pcc::g_trace_buf = device_buf_2x4096_i64;
  return 0;
}

extern "C" int pcc_phi_fused_supported(const pcc_phi_desc* d) { return check_phi_desc(d, __func__); }

extern "C" int64_t pcc_phi_fused_workspace_bytes(const pcc_phi_desc* d, int64_t n, int64_t B) {
  if (check_phi_desc(d, __func__) != 0) return -1;
  const int64_t fwd = ws_layout(d, n, B).total, bwd = phi_bwd_workspace_bytes(d, n);
  return fwd > bwd ? fwd : bwd;  // one query serves both directions
}

extern "C" int64_t pcc_phi_packed_bytes(const pcc_phi_desc* d) {
  if (check_phi_desc(d, __func__) != 0) return -1;
  return pack_layout(d->n_layers, d->hidden).total;
}

extern "C" int pcc_deepsets_phi_pool_fwd(const pcc_phi_desc* d, const float* x, const int64_t* offsets, int64_t n,
                                         int64_t B, float* pooled, int32_t* argmax, void* ws, void* wpack, int device,
                                         void* stream) {
  PCC_ENTER(device);
  if (check_phi_desc(d, __func__) != 0) return -1;
  PCC_REQUIRE(d->pooling != PCC_POOL_MAX || argmax != nullptr, "argmax buffer required for max pooling");
  PCC_REQUIRE(n < (int64_t)0x7fffffff, "row count exceeds int32 argmax range");
  cudaStream_t st = (cudaStream_t)stream;
  const int H = d->hidden, L = d->n_layers;
  const WsLayout wl = ws_layout(d, n, B);
  uint8_t* wsb = (uint8_t*)ws;

  PCC_REQUIRE(wpack != nullptr, "packed-weight buffer required (pcc_phi_packed_bytes)");
  const PackLayout pl = pack_layout(L, H);
  PackParams pk{};
  for (int l = 0; l < L; ++l) { pk.w[l] = d->w[l]; pk.w_off[l] = pl.w_off[l]; pk.wt_off[l] = pl.wt_off[l]; }
  pk.wpack = (uint8_t*)wpack; pk.d = d->input_dim; pk.H = H; pk.L = L;

  PhiParams p{};
  p.x = x; p.offsets = offsets; p.n = n; p.B = B; p.num_tiles = cdiv(n, kTileM);
  p.d = d->input_dim; p.L = L; p.pooling = d->pooling; p.res_mask = d->residual_mask;
  p.wpack = (const uint8_t*)wpack;
  for (int l = 0; l < L; ++l) { p.w_off[l] = pl.w_off[l]; p.bias[l] = d->b[l]; }
  p.pool_acc = wsb + wl.pool_off;
  p.trace = (long long*)g_trace_buf;
  p.tile_first = (const int32_t*)(wsb + wl.tile_first_off);
  PCC_K(fwd_prep_kernel)<<<dim3(32, 2 * L + 2), 256, 0, st>>>(pk, (unsigned long long*)(wsb + wl.pool_off), B * H, offsets, B,
                                                             p.num_tiles, (int32_t*)(wsb + wl.tile_first_off));
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

