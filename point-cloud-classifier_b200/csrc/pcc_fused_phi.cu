// Fused DeepSets phi-MLP + ragged pooling on tcgen05 / TMEM (sm_100a).
//   reference: /root/reference/models/deep_sets.py:89 (phi), :91-106 (split + pool loop)
//   and their autograd.  bf16 operands, fp32 accumulation in TMEM.
//
// Forward kernel (persistent, one CTA per SM, 128-point tiles), three concurrent roles per tile:
//   hidden warps (0-7)  : layer 0 (K = input_dim <= 15) on the FP32 pipe straight from x into the bf16
//                         SWIZZLE_128B activation image (no TMEM round trip: a TMEM read of a 128 x 256
//                         fp32 accumulator costs as much as the MMA that produced it), then the epilogue
//                         of the H x H hidden layer (TMEM -> bias/act/residual -> image, in place);
//                         every finished 64-column slab is handed to the MMA warp at once
//   MMA warp (17)       : hidden layer into accumulator A; final Linear computed TRANSPOSED
//                         (M = output features, N = points) into accumulator B
//   pool warps (8-15)   : thread = output feature, TMEM columns = the tile's points: in-lane masked
//                         max / sum over the ragged sets, partials combined across tiles with atomics
//                         in a [B,H] accumulator — overlapped with the NEXT tile's layers
//   producer warp (16)  : weight slabs (H rows x 128 B) from L2 through an mbarrier ring with
//                         cp.async.bulk (TMA engine).
// Per-point activations never leave the SM.
#include "pcc_fused.cuh"
#include "pcc_head.cuh"
#include <stdlib.h>

namespace pcc {

constexpr int kRingF = 4;  // weight slabs in flight (forward)
constexpr int kFwdHidWarps = 8, kFwdPoolWarps = 8;
constexpr int kFwdProdWarp = 16, kFwdMmaWarp = 17;
constexpr int kFwdThreads = 18 * 32;  // register cap 96: warps are allocated in groups of 4 (20 x 32 x 96 <= 64 K)

constexpr uint32_t kIndBytes = 2 * 128 * 128;  // set-indicator operand of the pooling MMA: [2 slabs][<=128 sets][64 points] bf16
struct SmemLayout {
  uint32_t bufA, ring, ind, ones, bimg, w0, xs, bars, total;
};
__host__ __device__ inline SmemLayout smem_layout(int H, int L, int Q, int poolh) {
  SmemLayout s;
  uint32_t o = 0;
  s.bufA = o; o += kTileM * H * 2;             // activation image, 1024-aligned slabs
  s.ring = o; o += (poolh ? kRingF - 1 : kRingF) * w_slab_bytes(H);   // weight slabs, 1024-aligned
  s.ind = o;  o += poolh ? kIndBytes : 0;
  s.ones = o; o += kTileM * kK0 * 2;           // un-swizzled [2][128][8] bf16 image: columns 0,1 = 1, rest 0
  s.bimg = o; o += (uint32_t)H * kK0 * 2;      // un-swizzled [2][H][8] bf16 image: hidden-layer bias as (hi, lo, 0 ...)
  s.w0 = o;   o += (uint32_t)H * 4 * Q * 4;    // layer-0 table (see fwd_prep_kernel)
  s.xs = o;   o += 2u * kTileM * 4 * Q * 4;    // layer-0 inputs of two tiles: [128][4Q] fp32 = {-, x_0 .. x_{4Q-2}}
  s.bars = o; o += 256;
  s.total = o;
  return s;
}

// debug trace: CTA 0 only, role 0 = hidden thread 0, role 1 = MMA thread, role 2 = pool thread 0; 3 x 4096 slots
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
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// one 32-column accumulator chunk (bias already inside: it enters through an extra MMA K step) -> activation,
// residual -> bf16 -> SW128 activation image
template <int ACT>
__device__ __forceinline__ void epi_store_chunk(const uint32_t (&v)[32], uint8_t* bufA, int r, int c, bool res) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint8_t* dst = bufA + act_chunk_off(r, c * 32 + q * 8);
    if (ACT == PCC_ACT_RELU && !res) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(
          pack_bf16x2_relu(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1])),
          pack_bf16x2_relu(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3])),
          pack_bf16x2_relu(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])),
          pack_bf16x2_relu(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7])));
      continue;
    }
    uint4 old = make_uint4(0u, 0u, 0u, 0u);
    if (res) old = *reinterpret_cast<const uint4*>(dst);
    const uint32_t oo[4] = {old.x, old.y, old.z, old.w};
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // packed-pair math: two columns per instruction
      uint64_t o = act2<ACT>(f32x2(__uint_as_float(v[q * 8 + 2 * j]), __uint_as_float(v[q * 8 + 2 * j + 1])));
      if (res) o = fadd2(o, bf16x2_to_f32x2(oo[j]));
      pk[j] = pack_bf16x2_pair(o);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// layer 0 of one 128-point tile on the FP32 pipe, slabs [s0, s1) of 64 columns: h_0 = act(W_0 x + b_0) -> bf16 activation
// image.  256 threads (t = 0..255): thread = 8 columns (cg) x 4 rows (rw + 32 i) of a slab, so the weights sit in
// registers and every shared-memory read is shared by 8 (x) or 4 (weights) lanes; column pairs as packed fp32 pairs
// (FFMA2 does two columns per issue slot with the row's input broadcast).  Every finished slab is handed to the MMA warp
// at once (8 warp arrivals on slab_ready[s]).
template <int ACT, int Q, int NSLAB>
__device__ __forceinline__ void layer0_slabs(const float* xT, const ulonglong2* w0S, uint8_t* bufA, uint64_t* slab_ready, int s0,
                                             int s1, int t, long long* trace, int tn, bool tr0) {
  constexpr int XW = 4 * Q;
  const int cg = t & 7, rw = t >> 3;
#pragma unroll 1
  for (int s = s0; s < s1; ++s) {
    uint64_t z[4][4];
#pragma unroll
    for (int qq = 0; qq < Q; ++qq) {
      ulonglong2 w[8];  // [pair][half]: {t0, t1} / {t2, t3} of quad qq, each a packed column pair
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] = w0S[(qq * NSLAB + s) * 64 + e * 8 + cg];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 xv = *reinterpret_cast<const float4*>(xT + (rw + 32 * i) * XW + 4 * qq);
        const uint64_t x0 = f32x2(xv.x, xv.x), x1 = f32x2(xv.y, xv.y), x2 = f32x2(xv.z, xv.z), x3 = f32x2(xv.w, xv.w);
#pragma unroll
        for (int pi = 0; pi < 4; ++pi) {
          if (qq == 0) z[i][pi] = ffma2(w[2 * pi].y, x1, ffma2(w[2 * pi + 1].x, x2, ffma2(w[2 * pi + 1].y, x3, w[2 * pi].x)));
          else z[i][pi] = ffma2(w[2 * pi].x, x0, ffma2(w[2 * pi].y, x1, ffma2(w[2 * pi + 1].x, x2, ffma2(w[2 * pi + 1].y, x3, z[i][pi]))));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t o[4];
#pragma unroll
      for (int pi = 0; pi < 4; ++pi) {
        float lo, hi;
        f32x2_unpack(z[i][pi], lo, hi);
        o[pi] = (ACT == PCC_ACT_RELU) ? pack_bf16x2_relu(lo, hi) : pack_bf16x2_pair(act2<ACT>(z[i][pi]));
      }
      *reinterpret_cast<uint4*>(bufA + act_chunk_off(rw + 32 * i, s * 64 + cg * 8)) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (tr0) trace_ev(trace, 0, tn, 50 + s);
    fence_proxy_async();
    mbar_arrive_warp(&slab_ready[s]);
    if (tr0) trace_ev(trace, 0, tn, 60 + s);
  }
}

// ------------------------------------------------------------------ forward kernel
// L = 2 (phi = Linear, final Linear) or 3 (one H x H hidden layer / ResidualBlock in between).
// TMEM: L = 3: hidden accumulator = columns [0,256), final accumulator = [256,512);
//       L = 2: the final accumulator alternates between the two halves from tile to tile.
template <int H, int ACT, int Q, bool POOLH>
__global__ void __launch_bounds__(kFwdThreads, 1) phi_pool_fwd_kernel(const PhiParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const SmemLayout lay = smem_layout(H, p.L, Q, POOLH ? 1 : 0);
  uint8_t* bufA = smem + lay.bufA;
  uint8_t* indS = smem + lay.ind;
  uint8_t* ring = smem + lay.ring;
  const ulonglong2* w0S = reinterpret_cast<const ulonglong2*>(smem + lay.w0);
  float* xS = reinterpret_cast<float*>(smem + lay.xs);
  __nv_bfloat16* onesS = reinterpret_cast<__nv_bfloat16*>(smem + lay.ones);
  __nv_bfloat16* bimgS = reinterpret_cast<__nv_bfloat16*>(smem + lay.bimg);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bars);
  uint64_t* full = bars;                        // [kRingF]
  uint64_t* empty = bars + kRingF;              // [kRingF]
  uint64_t* slab_ready = bars + 2 * kRingF;     // [4] 64-column slab of the activation image written
  uint64_t* acc_h = bars + 2 * kRingF + 4;      // hidden-layer accumulator complete
  uint64_t* acc_f = bars + 2 * kRingF + 5;      // [2] final accumulator (slot) complete
  uint64_t* pool_done = bars + 2 * kRingF + 7;  // [2] final accumulator (slot) drained by the pool warps
  uint64_t* ind_ready = bars + 2 * kRingF + 9;  // poolh: set-indicator operand of this tile built by the pool warps
  uint64_t* img_free = bars + 2 * kRingF + 10;  // poolh: every pooling MMA of the tile has finished reading the activation image
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRingF + 11);
  volatile uint32_t* nsetsS = tmem_slot + 1;    // poolh: padded number of sets (MMA N) of the current chunk

  constexpr uint32_t SLAB = w_slab_bytes(H);   // K = 64 slab of a weight image
  constexpr int HALVES = H / 128;              // M halves of the transposed final layer
  constexpr int NSLAB = H / 64;                // slabs per H x H layer
  constexpr int NCHUNK = H / 32;               // 32-column chunks of a hidden layer
  constexpr int XW = 4 * Q;                    // padded layer-0 row: {-, x_0 .. x_{4Q-2}}

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;
  // poolh (sum / mean pooling): sum_i (W h_i + b) = W (sum_i h_i) + n b, so the final Linear leaves the per-point
  // path: the kernel pools the last HIDDEN activations with one small MMA per tile,
  //   D[feature, set] += h^T[feature, point] * 1[point in set]      (A = MN-major view of the activation image,
  // B = 0/1 indicator built by the pool warps, N = sets intersecting the tile rounded up to 16), and the host
  // applies W_{L-1} to the [B, H] result.  No transposed final layer, 8x fewer accumulator columns to read.
  constexpr bool poolh = POOLH;
  const int ringn = poolh ? kRingF - 1 : kRingF;
  const int l_end = poolh ? L - 1 : L;   // layers streamed through the ring: 1 .. l_end-1

  // constant operands of the bias K step of the hidden layer: A = [128 x 16] with ones in columns 0, 1;
  // B = [H x 16] with (bf16 hi, bf16 lo) of b_1 in columns 0, 1 (hi + lo carries ~16 mantissa bits)
  for (int i = threadIdx.x; i < kTileM * kK0; i += kFwdThreads) {  // image [2][128][8]: i = (kc, row, k8)
    const int kc = i / (kTileM * 8), k8 = i % 8;
    onesS[i] = __float2bfloat16_rn((kc == 0 && k8 < 2) ? 1.f : 0.f);
  }
  for (int i = threadIdx.x; i < H * kK0; i += kFwdThreads) {       // image [2][H][8]
    const int kc = i / (H * 8), row = (i / 8) % H, k8 = i % 8;
    float v = 0.f;
    if (L == 3 && kc == 0 && k8 < 2) {
      const float b = __ldg(p.bias[1] + row);
      const float hi = bf16_round(b);
      v = (k8 == 0) ? hi : (b - hi);
    }
    bimgS[i] = __float2bfloat16_rn(v);
  }
  if (warp == kFwdMmaWarp) tmem_alloc<512>(tmem_slot);
  // everything above reads only parameters; the packed weights, the layer-0 table, the pool accumulator and the
  // tile tables below are written by fwd_prep_kernel, the predecessor in the stream (see pdl_enter)
  pdl_wait();
  for (int i = threadIdx.x; i < H * 4 * Q; i += kFwdThreads) reinterpret_cast<float*>(smem + lay.w0)[i] = __ldg(p.w0tab + i);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRingF; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 4; ++i) mbar_init(&slab_ready[i], kFwdHidWarps);
    mbar_init(acc_h, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_f[i], 1); mbar_init(&pool_done[i], 4 * HALVES); }
    mbar_init(ind_ready, kFwdPoolWarps);
    mbar_init(img_free, 1);
    fence_mbar_init();
  }
  fence_proxy_async();  // the two constant images are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kFwdProdWarp) {
    // ===================== producer: stream the weight slabs of layers 1 .. L-1 through the ring
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int l = 1; l < l_end; ++l) {
          for (int s = 0; s < NSLAB; ++s) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], SLAB);
            bulk_g2s(ring + stage * SLAB, p.wpack + p.w_off[l] + (size_t)s * SLAB, SLAB, &full[stage]);
            if (++stage == (uint32_t)ringn) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == kFwdMmaWarp) {
    // ===================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC_N = make_idesc_bf16(128, H, 0, 0);    // points x features
      constexpr uint32_t IDESC_T = make_idesc_bf16(128, 128, 0, 0);  // features(128) x points
      uint32_t stage = 0, phase = 0, sl_phase = 0, ind_phase = 0;
      uint32_t use = 0;  // poolh: running count of final-accumulator uses (one per chunk of <= 128 sets)
      int tn = 0;
      const uint32_t a_base = smem_u32(bufA), r_base = smem_u32(ring);
      int nt = 0;  // tiles done by this CTA
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++nt) {
        if (L == 3) {  // hidden layer -> accumulator A
          trace_ev(p.trace, 1, tn, 101);
          // z_1 starts as 1 * b_1 (K = 16 step on the two constant images); accumulator A is free here: its
          // previous reader, the hidden epilogue of the last tile, precedes the h_0 slabs waited for below
          // in the hidden warps' program order — so the bias step is issued after the first slab wait
          for (int s = 0; s < NSLAB; ++s) {
            mbar_wait(&slab_ready[s], sl_phase);  // h_0 slab s written by the hidden warps
            mbar_wait(&full[stage], phase);        // weight slab s landed
            tc_fence_after();
            if (s == 0) trace_ev(p.trace, 1, tn, 121);
            const uint32_t w_slab = r_base + stage * SLAB;
            if (s == 0)
              umma_bf16(tmem, make_smem_desc(smem_u32(onesS), kTileM * 16, 128), make_smem_desc(smem_u32(bimgS), H * 16, 128),
                        IDESC_N, 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16(tmem, make_smem_desc_sw128_k(a_base + s * kActSlab + ks * 32),
                        make_smem_desc_sw128_k(w_slab + ks * 32), IDESC_N, 1);
            umma_commit(&empty[stage]);
            if (++stage == (uint32_t)ringn) { stage = 0; phase ^= 1; }
          }
          sl_phase ^= 1;
          umma_commit(acc_h);
          trace_ev(p.trace, 1, tn, 141);
        }
        if (poolh) {
          // The sets that intersect the tile (interior empty ones included) are pooled in chunks of <= 128: one
          // pooling MMA group (N <= 128 columns) per chunk, each a separate use of a final-accumulator slot.
          const int nsets_t = __ldg(p.tile_last + tile) - __ldg(p.tile_first + tile) + 1;
          const int nch = (nsets_t + 127) >> 7;
          for (int s = 0; s < NSLAB; ++s) mbar_wait(&slab_ready[s], sl_phase);  // the whole activation image (K = points)
          sl_phase ^= 1;
          for (int ch = 0; ch < nch; ++ch, ++use) {
            const int slot = (L == 2) ? (int)(use & 1) : 0;
            const uint32_t k = (L == 2) ? (use >> 1) : use;
            if (k >= 1) mbar_wait(&pool_done[slot], (k - 1) & 1);
            const uint32_t accT = tmem + ((L == 2) ? slot * 256 : 256);
            if (ch == 0) trace_ev(p.trace, 1, tn, 102);
            mbar_wait(ind_ready, ind_phase);
            ind_phase ^= 1;
            tc_fence_after();
            const uint32_t N = *nsetsS;
            const uint32_t idesc = make_idesc_bf16(128, (int)N, 1, 0);
            const uint32_t i_base = smem_u32(indS);
#pragma unroll
            for (int h = 0; h < HALVES; ++h)
#pragma unroll
              for (int ks = 0; ks < kTileM / 16; ++ks)
                umma_bf16(accT + h * 128, make_smem_desc_sw128_mn(a_base + h * (2 * kActSlab) + ks * 2048, kActSlab),
                          make_smem_desc_sw128_k(i_base + (ks >> 2) * (N * 128) + (ks & 3) * 32), idesc, ks != 0);
            umma_commit(&acc_f[slot]);
          }
          umma_commit(img_free);
          trace_ev(p.trace, 1, tn, 142);
          continue;
        }
        // final layer (transposed) -> accumulator slot; the pool warps must have drained its previous use
        const int slot = (L == 2) ? (nt & 1) : 0;
        const int k = (L == 2) ? (nt >> 1) : nt;
        if (k >= 1) {
          mbar_wait(&pool_done[slot], (uint32_t)((k - 1) & 1));
          tc_fence_after();
        }
        const uint32_t accT = tmem + ((L == 2) ? slot * 256 : 256);
        trace_ev(p.trace, 1, tn, 102);
        for (int s = 0; !poolh && s < NSLAB; ++s) {
          mbar_wait(&slab_ready[s], sl_phase);
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (s == 0) trace_ev(p.trace, 1, tn, 122);
          const uint32_t w_slab = r_base + stage * SLAB;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t act_desc = make_smem_desc_sw128_k(a_base + s * kActSlab + ks * 32);
#pragma unroll
            for (int h = 0; h < HALVES; ++h)
              umma_bf16(accT + h * 128, make_smem_desc_sw128_k(w_slab + h * (128 * 128) + ks * 32), act_desc, IDESC_T,
                        (s | ks) != 0);
          }
          umma_commit(&empty[stage]);
          if (++stage == (uint32_t)ringn) { stage = 0; phase ^= 1; }
        }
        if (!poolh) {
          sl_phase ^= 1;
          umma_commit(&acc_f[slot]);
          trace_ev(p.trace, 1, tn, 142);
        }
      }
    }
  } else if (warp < kFwdHidWarps) {
    // ===================== hidden warps 0-7: layer 0 on the FP32 pipe + hidden-layer epilogue
    const int quarter = warp & 3, grp = warp >> 2;
    const int r = quarter * 32 + lane;  // tile row (hidden-layer epilogue)
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    const int d = p.d;
    // x tile -> registers -> xS (padded rows, bf16-rounded like every MMA operand: the backward recomputes
    // this layer with MMAs): slot s = tid + 256 k covers row s / XW, column s % XW - 1
    float xp[2 * Q];
    auto load_x = [&](int64_t tile) {
#pragma unroll
      for (int k = 0; k < 2 * Q; ++k) {
        const int slot = threadIdx.x + kFwdHidWarps * 32 * k;
        const int64_t row = tile * kTileM + slot / XW;
        const int j = slot % XW - 1;
        xp[k] = (j >= 0 && j < d && row < p.n && tile < p.num_tiles) ? __ldg(p.x + row * d + j) : 0.f;  // raw: no use before store_x
      }
    };
    auto store_x = [&](int buf) {
#pragma unroll
      for (int k = 0; k < 2 * Q; ++k) xS[buf * (kTileM * XW) + threadIdx.x + kFwdHidWarps * 32 * k] = bf16_round(xp[k]);
      asm volatile("bar.sync 1, 256;" ::: "memory");  // hidden warps only
    };
    load_x(blockIdx.x);
    store_x(0);
    // (poolh: the pool warps compute half of layer 0 and wait on barrier 4 for the staged inputs of their tile)
    if (POOLH && (int64_t)blockIdx.x < p.num_tiles) asm volatile("bar.arrive 4, 512;" ::: "memory");
    int tn = 0;
    uint32_t acch_phase = 0;
    const bool tr0 = (threadIdx.x == 0);
    int nt = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++nt) {
      // ---- the activation image is free once the final layer of the previous tile has completed
      if (nt >= 1) {
        if (poolh) {
          mbar_wait(img_free, (uint32_t)((nt - 1) & 1));
        } else {
          const int pslot = (L == 2) ? ((nt - 1) & 1) : 0;
          const int pk = (L == 2) ? ((nt - 1) >> 1) : (nt - 1);
          mbar_wait(&acc_f[pslot], (uint32_t)(pk & 1));
        }
      }
      if (tr0) trace_ev(p.trace, 0, tn, 0);
      // ---- layer 0: h_0 = act(W_0 x + b_0), one 64-column slab at a time, handed to the MMA warp at once
      const float* xT = xS + (nt & 1) * (kTileM * XW);
      // next tile's inputs: issued here so that they have landed before the first fence below (the proxy fence
      // drains every outstanding memory operation of the thread, global loads included)
      load_x(tile + gridDim.x);
      // (sum / mean pooling: the pool warps have almost nothing to do per tile, so they take the second half of the slabs
      //  and of the hidden epilogue below — the tile period of the yaml model was 17k cycles, 9.1k of them this layer on
      //  8 warps and 4.6k the epilogue, all serial)
      layer0_slabs<ACT, Q, NSLAB>(xT, w0S, bufA, slab_ready, 0, POOLH ? NSLAB / 2 : NSLAB, threadIdx.x, p.trace, tn, tr0);
      if (tr0) trace_ev(p.trace, 0, tn, 5);

      // ---- hidden layer: TMEM -> bias/act/residual -> bf16 image (in place); TMEM loads run one chunk
      //      ahead of the math; every finished 64-column slab is handed to the MMA warp at once
      if (L == 3) {
        mbar_wait(acc_h, acch_phase);
        acch_phase ^= 1;
        tc_fence_after();
        if (tr0) trace_ev(p.trace, 0, tn, 11);
        const bool res = (p.res_mask >> 1) & 1;
        uint32_t va[32], vb[32];
        if (POOLH) {
          // four warp groups (two of them pool warps): chunks grp, grp + 4
          tmem_ld32(lane_base + grp * 32, va);
          if (grp + 4 < NCHUNK) tmem_ld32(lane_base + (grp + 4) * 32, vb);
          tmem_wait_ld();
          epi_store_chunk<ACT>(va, bufA, r, grp, res);
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive_warp(&slab_ready[grp >> 1]);
          if (grp + 4 < NCHUNK) {
            epi_store_chunk<ACT>(vb, bufA, r, grp + 4, res);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive_warp(&slab_ready[(grp + 4) >> 1]);
          }
        } else {
        tmem_ld32(lane_base + grp * 32, va);
#pragma unroll 1
        for (int c = grp; c < NCHUNK; c += 4) {
          tmem_wait_ld();
          if (c + 2 < NCHUNK) tmem_ld32(lane_base + (c + 2) * 32, vb);
          epi_store_chunk<ACT>(va, bufA, r, c, res);
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive_warp(&slab_ready[c >> 1]);
          if (c + 2 < NCHUNK) {
            tmem_wait_ld();
            if (c + 4 < NCHUNK) tmem_ld32(lane_base + (c + 4) * 32, va);
            epi_store_chunk<ACT>(vb, bufA, r, c + 2, res);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive_warp(&slab_ready[(c + 2) >> 1]);
          }
        }
        }
        if (tr0) trace_ev(p.trace, 0, tn, 21);
      }
      store_x((nt + 1) & 1);  // buffer last read by h_0 of the previous tile; every warp is past it (barrier above)
      if (POOLH && tile + gridDim.x < p.num_tiles) asm volatile("bar.arrive 4, 512;" ::: "memory");
    }
  } else if (warp < kFwdHidWarps + kFwdPoolWarps) {
    // ===================== pool warps 8-15: thread = feature, TMEM columns = the tile's points
    const int pw = warp - kFwdHidWarps;
    const int quarter = warp & 3, h = pw >> 2;
    if (poolh) {
      // ---- poolh: build the 0/1 set-indicator operand of the tile, then read the few pooled columns
      const int pt = pw * 32 + lane;  // 0..255
      const int f = h * 128 + quarter * 32 + lane;
      const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
      const bool tr0 = (pw == 0 && lane == 0);
      float* hsum = reinterpret_cast<float*>(p.pool_acc);
      int tn = 0;
      uint32_t use = 0;  // running count of final-accumulator uses, in step with the MMA thread
      int nt = 0;
      uint32_t acch_phase = 0;
      // indicator operand of sets [b_first, b_first + nsets) over the tile's points
      auto build_indicator = [&](int64_t r0, int b_first, int nsets, uint32_t N) {
        for (uint32_t q = pt; q < N * 16; q += 32 * kFwdPoolWarps) {
          const uint32_t sidx = q >> 4, kc = q & 15;   // set slot, 8-point chunk
          int64_t lo = 0, hi = 0;
          if ((int)sidx < nsets) { lo = __ldg(p.offsets + b_first + sidx); hi = __ldg(p.offsets + b_first + sidx + 1); }
          const int64_t p0 = r0 + kc * 8;
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t b0 = (p0 + 2 * j >= lo && p0 + 2 * j < hi) ? 0x3F80u : 0u;          // bf16 1.0
            const uint32_t b1 = (p0 + 2 * j + 1 >= lo && p0 + 2 * j + 1 < hi) ? 0x3F80u : 0u;
            w[j] = b0 | (b1 << 16);
          }
          *reinterpret_cast<uint4*>(indS + (kc >> 3) * (N * 128) + sidx * 128 + (((kc & 7) ^ (sidx & 7)) << 4)) =
              make_uint4(w[0], w[1], w[2], w[3]);
        }
        if (pt == 0) *nsetsS = N;
        fence_proxy_async();
        mbar_arrive_warp(ind_ready);
      };
      // pooled columns of one chunk: final accumulator -> [B,H] sums
      auto read_pooled = [&](int b_first, int nsets, uint32_t N) {
        const int slot = (L == 2) ? (int)(use & 1) : 0;
        const uint32_t k = (L == 2) ? (use >> 1) : use;
        mbar_wait(&acc_f[slot], k & 1);
        tc_fence_after();
        if (tr0) trace_ev(p.trace, 2, tn, 30);
        if (h < HALVES) {
          const uint32_t accT = lane_base + ((L == 2) ? slot * 256 : 256) + h * 128;
          for (uint32_t c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(accT + c0, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if ((int)(c0 + j) < nsets) atomicAdd(hsum + (int64_t)(b_first + c0 + j) * H + f, __uint_as_float(v[j]));
          }
          tc_fence_before();
          mbar_arrive_warp(&pool_done[slot]);
        }
        if (tr0) trace_ev(p.trace, 2, tn, 40);
        ++use;
      };
      // chunks of <= 128 sets (a tile of 128 one-point sets plus interior empty sets holds more than 128): the first
      // chunk's indicator is built right after this group's layer-0 share, the pooled columns are read at the start of
      // the next tile
      auto finish_tile = [&](int64_t tile_f) {
        const int64_t r0 = tile_f * kTileM;
        const int b_first0 = __ldg(p.tile_first + tile_f), b_last = __ldg(p.tile_last + tile_f);
        const int nsets_t = b_last - b_first0 + 1;
        for (int c0s = 0; c0s < nsets_t; c0s += 128) {
          const int b_first = b_first0 + c0s;
          const int nsets = (nsets_t - c0s < 128) ? nsets_t - c0s : 128;
          const uint32_t N = (uint32_t)((nsets + 15) & ~15);   // MMA N: multiple of 16, >= 16, <= 128
          if (c0s > 0) build_indicator(r0, b_first, nsets, N);
          read_pooled(b_first, nsets, N);
        }
      };
      int64_t pending = -1;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++nt) {
        // ---- this group's share of the per-point work: the second half of layer 0's slabs (the hidden warps staged
        //      the tile's inputs: barrier 4; the image is free once the previous tile's pooling MMAs have read it)
        if (pending >= 0) finish_tile(pending);   // pooled columns of the previous tile (measured: reading them after
                                                  // this tile's layer-0 share instead only moves the wait to the MMA)
        if (nt >= 1) mbar_wait(img_free, (uint32_t)((nt - 1) & 1));
        asm volatile("bar.sync 4, 512;" ::: "memory");
        layer0_slabs<ACT, Q, NSLAB>(xS + (nt & 1) * (kTileM * XW), w0S, bufA, slab_ready, NSLAB / 2, NSLAB, pt, nullptr, 0, false);
        // ---- first chunk's indicator of this tile (the previous pooling MMA, which read the indicator, is complete)
        {
          const int b_first0 = __ldg(p.tile_first + tile), b_last = __ldg(p.tile_last + tile);
          const int nsets_t = b_last - b_first0 + 1;
          const int nsets = nsets_t < 128 ? nsets_t : 128;
          build_indicator(tile * kTileM, b_first0, nsets, (uint32_t)((nsets + 15) & ~15));
        }
        if (L == 3) {
          // ---- and of the hidden-layer epilogue: warp groups 2, 3 of four (chunks grp, grp + 4)
          mbar_wait(acc_h, acch_phase);
          acch_phase ^= 1;
          tc_fence_after();
          const bool res = (p.res_mask >> 1) & 1;
          const int grp = 2 + h, r = quarter * 32 + lane;
          uint32_t va[32], vb[32];
          tmem_ld32(lane_base + grp * 32, va);
          if (grp + 4 < NCHUNK) tmem_ld32(lane_base + (grp + 4) * 32, vb);
          tmem_wait_ld();
          epi_store_chunk<ACT>(va, bufA, r, grp, res);
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive_warp(&slab_ready[grp >> 1]);
          if (grp + 4 < NCHUNK) {
            epi_store_chunk<ACT>(vb, bufA, r, grp + 4, res);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive_warp(&slab_ready[(grp + 4) >> 1]);
          }
        }
        pending = tile;
      }
      if (pending >= 0) finish_tile(pending);
    } else if (h < HALVES) {
      const int r = quarter * 32 + lane;
      const int f = h * 128 + r;
      const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
      const bool is_max = (p.pooling == PCC_POOL_MAX);
      const bool tr0 = (pw == 0 && lane == 0);
      int tn = 0;
      int nt = 0;
      // segment lookups run one tile ahead (two dependent global loads) so they never sit on the pooling path
      int64_t nb = 0, nlo = 0, nhi = 0;
      auto seg_fetch = [&](int64_t tile) {
        nb = p.B; nlo = 0; nhi = 0;
        if (tile < p.num_tiles) {
          nb = __ldg(p.tile_first + tile);
          if (nb < p.B) { nlo = __ldg(p.offsets + nb); nhi = __ldg(p.offsets + nb + 1); }
        }
      };
      seg_fetch(blockIdx.x);
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++nt) {
        const int64_t r0 = tile * kTileM;
        const int64_t tile_end = (r0 + kTileM < p.n) ? r0 + kTileM : p.n;
        int64_t b = nb, seg_lo = nlo, seg_hi = nhi;
        seg_fetch(tile + gridDim.x);
        const int slot = (L == 2) ? (nt & 1) : 0;
        const int k = (L == 2) ? (nt >> 1) : nt;
        mbar_wait(&acc_f[slot], (uint32_t)(k & 1));
        tc_fence_after();
        if (tr0) trace_ev(p.trace, 2, tn, 30);
        const uint32_t accT = lane_base + ((L == 2) ? slot * 256 : 256) + h * 128;
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
        tmem_ld32(accT, va);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_wait_ld();
          uint32_t (&v)[32] = (c & 1) ? vb : va;
          if (c + 1 < 4) tmem_ld32(accT + (c + 1) * 32, (c & 1) ? va : vb);
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
        // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above): hand it back
        tc_fence_before();
        mbar_arrive_warp(&pool_done[slot]);
        if (b < p.B && seg_lo < tile_end) flush(b);  // partial of the set that continues past this tile
        if (tr0) trace_ev(p.trace, 2, tn, 40);
      }
    }
  }

  __syncthreads();
  if (warp == kFwdMmaWarp) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------ forward kernel, CTA pairs (H = 256, max pooling)
// Same three roles as above, but two CTAs (one cluster, two SMs of a TPC) work on two consecutive tiles with
// cta_group::2 MMAs, each CTA holding HALF of every weight matrix resident in shared memory:
//   hidden layer   D[256 points, 256 features]: A = each CTA's own activation image, B = W_1 split by output feature;
//   final layer    D[256 features, 256 points]: A = W_{L-1} split by output feature, B = each CTA's own image
//                  -> CTA c pools ITS 128 features over BOTH tiles.
// What this buys (profiles/notes_r1.md: the single-CTA kernel is bound by shared-memory bandwidth): no weight
// stream at all (256 KB per tile written into shared memory before), the transposed final layer becomes one
// M = 256 MMA per K step reading 64 B/clk per SM instead of two M = 128 MMAs reading 128 B/clk, and the hidden
// layer reads 64 instead of 96 B/clk.
constexpr int kPairMmaWarp = 16;
constexpr int kPairThreads = 17 * 32;

struct PairSmem {
  uint32_t bufA, w1h, w2h, ones, bimg, w0, xs, bars, total;
};
__host__ __device__ inline PairSmem pair_smem(int Q) {
  PairSmem s;
  uint32_t o = 0;
  s.bufA = o; o += kTileM * 256 * 2;           // activation image of this CTA's tile
  s.w1h = o;  o += 4 * kActSlab;               // rows [128 rank, +128) of the hidden weight image, 4 K slabs
  s.w2h = o;  o += 4 * kActSlab;               // same rows of the final weight image
  s.ones = o; o += kTileM * kK0 * 2;
  s.bimg = o; o += kTileM * kK0 * 2;           // [2][128][8]: (hi, lo) of b_1 for this CTA's 128 features
  s.w0 = o;   o += 256u * 4 * Q * 4;
  s.xs = o;   o += 2u * kTileM * 4 * Q * 4;
  s.bars = o; o += 256;
  s.total = o;
  return s;
}

template <int ACT, int Q>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1) phi_pool_fwd_pair_kernel(const PhiParams p) {
  constexpr int H = 256;
  extern __shared__ __align__(1024) uint8_t smem[];
  const PairSmem lay = pair_smem(Q);
  uint8_t* bufA = smem + lay.bufA;
  uint8_t* w1h = smem + lay.w1h;
  uint8_t* w2h = smem + lay.w2h;
  __nv_bfloat16* onesS = reinterpret_cast<__nv_bfloat16*>(smem + lay.ones);
  __nv_bfloat16* bimgS = reinterpret_cast<__nv_bfloat16*>(smem + lay.bimg);
  const ulonglong2* w0S = reinterpret_cast<const ulonglong2*>(smem + lay.w0);
  float* xS = reinterpret_cast<float*>(smem + lay.xs);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bars);
  uint64_t* slab_ready = bars;        // [4] (leader's copy is used) slab of BOTH images written: 16 warp arrivals
  uint64_t* acc_h = bars + 4;         // hidden accumulator complete (multicast commit: both CTAs)
  uint64_t* acc_f = bars + 5;         // [2] final accumulator (slot) complete (multicast)
  uint64_t* pool_done = bars + 7;     // [2] (leader's copy) final accumulator drained in BOTH CTAs: 16 warp arrivals
  uint64_t* wres = bars + 9;          // this CTA's weight halves have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  constexpr uint32_t SLAB = w_slab_bytes(H);   // K = 64 slab of a full weight image (256 rows)
  constexpr int NSLAB = H / 64, NCHUNK = H / 32;
  constexpr int XW = 4 * Q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;
  const uint32_t rank = cluster_ctarank();
  const int64_t npairs = gridDim.x >> 1, pair0 = blockIdx.x >> 1;
  const int64_t ntp = (p.num_tiles + 1) >> 1;  // tile pairs

  // ---- constants: ones image, bias halves, layer-0 table
  for (int i = threadIdx.x; i < kTileM * kK0; i += kPairThreads) {
    const int kc = i / (kTileM * 8), row = (i / 8) % kTileM, k8 = i % 8;
    onesS[i] = __float2bfloat16_rn((kc == 0 && k8 < 2) ? 1.f : 0.f);
    float v = 0.f;
    if (L == 3 && kc == 0 && k8 < 2) {
      const float b = __ldg(p.bias[1] + rank * 128 + row);
      const float hi = bf16_round(b);
      v = (k8 == 0) ? hi : (b - hi);
    }
    bimgS[i] = __float2bfloat16_rn(v);
  }
  if (warp == kPairMmaWarp) tmem_alloc_pair<512>(tmem_slot);
  // everything above reads only parameters; what follows was written by fwd_prep_kernel, the predecessor in the stream
  pdl_wait();
  for (int i = threadIdx.x; i < H * 4 * Q; i += kPairThreads) reinterpret_cast<float*>(smem + lay.w0)[i] = __ldg(p.w0tab + i);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&slab_ready[i], 2 * kFwdHidWarps);
    mbar_init(acc_h, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_f[i], 1); mbar_init(&pool_done[i], 2 * kFwdPoolWarps); }
    mbar_init(wres, 1);
    fence_mbar_init();
    // resident weight halves: rows [128 rank, +128) of every K slab = one contiguous 16 KB piece per slab
    const uint32_t nb = (uint32_t)((L == 3 ? 2 : 1) * NSLAB) * kActSlab;
    mbar_arrive_expect_tx(wres, nb);
    for (int sidx = 0; sidx < NSLAB; ++sidx) {
      if (L == 3) bulk_g2s(w1h + sidx * kActSlab, p.wpack + p.w_off[1] + (size_t)sidx * SLAB + (size_t)rank * 128 * 128, kActSlab, wres);
      bulk_g2s(w2h + sidx * kActSlab, p.wpack + p.w_off[L - 1] + (size_t)sidx * SLAB + (size_t)rank * 128 * 128, kActSlab, wres);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // barriers of both CTAs initialised before any remote arrival; both TMEM allocations done
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  mbar_wait(wres, 0);

  if (warp == kPairMmaWarp) {
    // ===================== MMA issuer: one thread of the leader CTA drives both tensor cores
    if (rank == 0 && lane == 0) {
      constexpr uint32_t IDESC = make_idesc_bf16(256, 256, 0, 0);
      uint32_t sl_phase = 0;
      const uint32_t a_base = smem_u32(bufA), w1_base = smem_u32(w1h), w2_base = smem_u32(w2h);
      int tn = 0, nt = 0;
      for (int64_t tp = pair0; tp < ntp; tp += npairs, ++nt) {
        if (L == 3) {
          trace_ev(p.trace, 1, tn, 101);
          for (int sidx = 0; sidx < NSLAB; ++sidx) {
            mbar_wait_cluster(&slab_ready[sidx], sl_phase);  // h_0 slab of both tiles
            tc_fence_after();
            if (sidx == 0) {
              trace_ev(p.trace, 1, tn, 121);
              umma_bf16_pair(tmem, make_smem_desc(smem_u32(onesS), kTileM * 16, 128), make_smem_desc(smem_u32(bimgS), kTileM * 16, 128),
                             IDESC, 0);
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16_pair(tmem, make_smem_desc_sw128_k(a_base + sidx * kActSlab + ks * 32),
                             make_smem_desc_sw128_k(w1_base + sidx * kActSlab + ks * 32), IDESC, 1);
          }
          sl_phase ^= 1;
          umma_commit_pair(acc_h);
          trace_ev(p.trace, 1, tn, 141);
        }
        const int slot = (L == 2) ? (nt & 1) : 0;
        const int k = (L == 2) ? (nt >> 1) : nt;
        if (k >= 1) {
          mbar_wait_cluster(&pool_done[slot], (uint32_t)((k - 1) & 1));
          tc_fence_after();
        }
        const uint32_t accT = tmem + ((L == 2) ? slot * 256 : 256);
        trace_ev(p.trace, 1, tn, 102);
        for (int sidx = 0; sidx < NSLAB; ++sidx) {
          mbar_wait_cluster(&slab_ready[sidx], sl_phase);
          tc_fence_after();
          if (sidx == 0) trace_ev(p.trace, 1, tn, 122);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16_pair(accT, make_smem_desc_sw128_k(w2_base + sidx * kActSlab + ks * 32),
                           make_smem_desc_sw128_k(a_base + sidx * kActSlab + ks * 32), IDESC, (sidx | ks) != 0);
        }
        sl_phase ^= 1;
        umma_commit_pair(&acc_f[slot]);
        trace_ev(p.trace, 1, tn, 142);
      }
    }
  } else if (warp < kFwdHidWarps) {
    // ===================== hidden warps: layer 0 on the FP32 pipe + hidden-layer epilogue (this CTA's tile)
    const int quarter = warp & 3, grp = warp >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    const int d = p.d;
    const int cg = threadIdx.x & 7, rw = threadIdx.x >> 3;
    float xp[2 * Q];
    auto load_x = [&](int64_t tile) {
#pragma unroll
      for (int k = 0; k < 2 * Q; ++k) {
        const int slot = threadIdx.x + kFwdHidWarps * 32 * k;
        const int64_t row = tile * kTileM + slot / XW;
        const int j = slot % XW - 1;
        xp[k] = (j >= 0 && j < d && row < p.n && tile < p.num_tiles) ? __ldg(p.x + row * d + j) : 0.f;
      }
    };
    auto store_x = [&](int buf) {
#pragma unroll
      for (int k = 0; k < 2 * Q; ++k) xS[buf * (kTileM * XW) + threadIdx.x + kFwdHidWarps * 32 * k] = bf16_round(xp[k]);
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    load_x(2 * pair0 + rank);
    store_x(0);
    int tn = 0, nt = 0;
    uint32_t acch_phase = 0;
    const bool tr0 = (threadIdx.x == 0 && rank == 0);
    for (int64_t tp = pair0; tp < ntp; tp += npairs, ++nt) {
      if (tr0) trace_ev(p.trace, 0, tn, 0);
      const float* xT = xS + (nt & 1) * (kTileM * XW);
      load_x(2 * (tp + npairs) + rank);
      float4 xrow[Q][4];  // the thread's four input rows stay in registers for all slabs
#pragma unroll
      for (int qq = 0; qq < Q; ++qq)
#pragma unroll
        for (int i = 0; i < 4; ++i) xrow[qq][i] = *reinterpret_cast<const float4*>(xT + (rw + 32 * i) * XW + 4 * qq);
#pragma unroll 1
      for (int sidx = 0; sidx < NSLAB; ++sidx) {
        uint64_t z[4][4];
#pragma unroll
        for (int qq = 0; qq < Q; ++qq) {
          ulonglong2 w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w[e] = w0S[(qq * NSLAB + sidx) * 64 + e * 8 + cg];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 xv = xrow[qq][i];
            const uint64_t x0 = f32x2(xv.x, xv.x), x1 = f32x2(xv.y, xv.y), x2 = f32x2(xv.z, xv.z), x3 = f32x2(xv.w, xv.w);
#pragma unroll
            for (int pi = 0; pi < 4; ++pi) {
              if (qq == 0) z[i][pi] = ffma2(w[2 * pi].y, x1, ffma2(w[2 * pi + 1].x, x2, ffma2(w[2 * pi + 1].y, x3, w[2 * pi].x)));
              else z[i][pi] = ffma2(w[2 * pi].x, x0, ffma2(w[2 * pi].y, x1, ffma2(w[2 * pi + 1].x, x2, ffma2(w[2 * pi + 1].y, x3, z[i][pi]))));
            }
          }
        }
        uint32_t o[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int pi = 0; pi < 4; ++pi) {
            float lo, hi;
            f32x2_unpack(z[i][pi], lo, hi);
            o[i][pi] = (ACT == PCC_ACT_RELU) ? pack_bf16x2_relu(lo, hi) : pack_bf16x2_pair(act2<ACT>(z[i][pi]));
          }
        // The image is free once the previous final layer has completed (multicast commit).  The first slab of the
        // new tile has been computed in registers by now: its math hides under the tail of that final layer.
        if (sidx == 0 && nt >= 1) {
          const int pslot = (L == 2) ? ((nt - 1) & 1) : 0;
          const int pk = (L == 2) ? ((nt - 1) >> 1) : (nt - 1);
          mbar_wait(&acc_f[pslot], (uint32_t)(pk & 1));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(bufA + act_chunk_off(rw + 32 * i, sidx * 64 + cg * 8)) =
              make_uint4(o[i][0], o[i][1], o[i][2], o[i][3]);
        fence_proxy_async();
        mbar_arrive_warp_remote(&slab_ready[sidx], 0);
      }
      if (tr0) trace_ev(p.trace, 0, tn, 5);
      if (L == 3) {
        mbar_wait(acc_h, acch_phase);
        acch_phase ^= 1;
        tc_fence_after();
        if (tr0) trace_ev(p.trace, 0, tn, 11);
        const bool res = (p.res_mask >> 1) & 1;
        uint32_t va[32], vb[32];
        tmem_ld32(lane_base + grp * 32, va);
#pragma unroll 1
        for (int c = grp; c < NCHUNK; c += 4) {
          tmem_wait_ld();
          if (c + 2 < NCHUNK) tmem_ld32(lane_base + (c + 2) * 32, vb);
          epi_store_chunk<ACT>(va, bufA, r, c, res);
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive_warp_remote(&slab_ready[c >> 1], 0);
          if (c + 2 < NCHUNK) {
            tmem_wait_ld();
            if (c + 4 < NCHUNK) tmem_ld32(lane_base + (c + 4) * 32, va);
            epi_store_chunk<ACT>(vb, bufA, r, c + 2, res);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive_warp_remote(&slab_ready[(c + 2) >> 1], 0);
          }
        }
        if (tr0) trace_ev(p.trace, 0, tn, 21);
      }
      store_x((nt + 1) & 1);
    }
  } else {
    // ===================== pool warps: thread = one of THIS CTA's 128 features, columns = points of tile 2 tp + h
    const int pw = warp - kFwdHidWarps;
    const int quarter = warp & 3, h = pw >> 2;
    const int f = (int)rank * 128 + quarter * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    const bool is_max = (p.pooling == PCC_POOL_MAX);
    const bool tr0 = (pw == 0 && lane == 0 && rank == 0);
    int tn = 0, nt = 0;
    int64_t nb = 0, nlo = 0, nhi = 0;
    auto seg_fetch = [&](int64_t tile) {
      nb = p.B; nlo = 0; nhi = 0;
      if (tile < p.num_tiles) {
        nb = __ldg(p.tile_first + tile);
        if (nb < p.B) { nlo = __ldg(p.offsets + nb); nhi = __ldg(p.offsets + nb + 1); }
      }
    };
    seg_fetch(2 * pair0 + h);
    for (int64_t tp = pair0; tp < ntp; tp += npairs, ++nt) {
      const int64_t tile = 2 * tp + h;
      const int64_t r0 = tile * kTileM;
      const int64_t tile_end = (r0 + kTileM < p.n) ? r0 + kTileM : p.n;
      int64_t b = nb, seg_lo = nlo, seg_hi = nhi;
      seg_fetch(2 * (tp + npairs) + h);
      const int slot = (L == 2) ? (nt & 1) : 0;
      const int k = (L == 2) ? (nt >> 1) : nt;
      mbar_wait(&acc_f[slot], (uint32_t)(k & 1));
      tc_fence_after();
      if (tr0) trace_ev(p.trace, 2, tn, 30);
      const uint32_t accT = lane_base + ((L == 2) ? slot * 256 : 256) + h * 128;
      float acc = is_max ? -INFINITY : 0.f;
      int arg = -1;
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
      tmem_ld32(accT, va);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_wait_ld();
        uint32_t (&v)[32] = (c & 1) ? vb : va;
        if (c + 1 < 4) tmem_ld32(accT + (c + 1) * 32, (c & 1) ? va : vb);
        const int64_t col0 = r0 + c * 32;
        while (b < p.B && seg_lo < tile_end && seg_lo < col0 + 32) {
          const int lo = (int)((seg_lo > col0 ? seg_lo : col0) - col0);
          const int hi = (int)((seg_hi < col0 + 32 ? seg_hi : col0 + 32) - col0);
          if (lo == 0 && hi == 32) {
            if (is_max) {
              const float m = fmaxf(fmaxf(max8(v), max8(v + 8)), fmaxf(max8(v + 16), max8(v + 24)));
              if (m > acc || arg < 0) {
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
          if (seg_hi > col0 + 32) break;
          flush(b);
          ++b;
          if (b < p.B) { seg_lo = seg_hi; seg_hi = __ldg(p.offsets + b + 1); }
        }
      }
      tc_fence_before();
      mbar_arrive_warp_remote(&pool_done[slot], 0);
      if (b < p.B && seg_lo < tile_end) flush(b);
      if (tr0) trace_ev(p.trace, 2, tn, 40);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be reading this CTA's operands / signalling its barriers
  if (warp == kPairMmaWarp) tmem_dealloc_pair<512>(tmem);
}

// pool accumulator -> pooled[B,H] (+ argmax): adds the final bias after pooling
// (max(z+b) = max(z)+b, mean(z+b) = mean(z)+b, sum(z+b)/sqrt(n) = sum(z)/sqrt(n) + b*sqrt(n))
__global__ void pool_finalize_kernel(const void* __restrict__ pool_acc, const int64_t* __restrict__ offsets,
                                     const float* __restrict__ bias, int64_t B, int H, int pooling,
                                     float* __restrict__ pooled, int32_t* __restrict__ argmax) {
  pdl_enter();
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
  int64_t pool_off, tile_first_off, tile_last_off, bscale_off, total;
};
static WsLayout ws_layout(const pcc_phi_desc* d, int64_t n, int64_t B) {
  WsLayout w{};
  int64_t o = 0;
  w.pool_off = o;
  o += B * d->hidden * 8;
  o = (o + 255) / 256 * 256;
  w.tile_first_off = o;
  o += cdiv(n, kTileM) * 4;
  o = (o + 255) / 256 * 256;
  w.tile_last_off = o;
  o += cdiv(n, kTileM) * 4;
  o = (o + 255) / 256 * 256;
  w.bscale_off = o;
  o += B * 4;
  w.total = (o + 255) / 256 * 256;
  return w;
}

// ONE launch for everything the forward kernel needs prepared: blockIdx.y < 2L packs weight image
// (layer y>>1, transposed if y&1); y == 2L zeroes the pool accumulator; y == 2L+1 computes tile_first;
// y == 2L+2 writes the fp32 layer-0 table {b_0[c], bf16(W_0[c][0..d-1]), 0...} of the forward kernel
__global__ void fwd_prep_kernel(PackParams pk, unsigned long long* pool, int64_t pool_count, const int64_t* offsets,
                                int64_t B, int64_t num_tiles, int32_t* tile_first, int32_t* tile_last, int64_t n) {
  pdl_enter();
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
  } else if (y == 2 * pk.L + 2) {
    // layer-0 table of the forward kernel, 16-byte entries [qq][slab][e = 2 pair + half][cg]: the four terms
    // t_0..t_3 of quad qq (qq = 0: b_0, w_0, w_1, w_2; else w_{4qq-1} .. w_{4qq+2}; bf16-rounded weights, zero
    // padded) for the column pair (c, c + 1), c = 64 slab + 8 cg + 2 pair: half 0 = {t0[c], t0[c+1], t1[c], t1[c+1]},
    // half 1 = {t2[c], t2[c+1], t3[c], t3[c+1]}
    float* tab = reinterpret_cast<float*>(pk.wpack + pk.w0tab_off);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pk.H * 4 * pk.q4; i += gridDim.x * blockDim.x) {
      const int f = i & 3, ent = i >> 2;
      const int cg = ent & 7, e = (ent >> 3) & 7, slab = (ent >> 6) % (pk.H / 64), qq = (ent >> 6) / (pk.H / 64);
      const int c = 64 * slab + 8 * cg + 2 * (e >> 1) + (f & 1);
      const int term = 2 * (e & 1) + (f >> 1);
      const int j = 4 * qq + term - 1;
      tab[i] = (j < 0) ? __ldg(pk.b0 + c) : (j < pk.d ? bf16_round(__ldg(pk.w[0] + (int64_t)c * pk.d + j)) : 0.f);
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < num_tiles; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r0 = i * kTileM;
      int64_t lo = 0, hi = B;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid + 1) <= r0) lo = mid + 1; else hi = mid;
      }
      tile_first[i] = (int32_t)lo;
      // last set intersecting the tile = set of its last valid row
      const int64_t rl = (r0 + kTileM - 1 < n - 1) ? r0 + kTileM - 1 : n - 1;
      int64_t lo2 = lo, hi2 = B;
      while (lo2 < hi2) {
        int64_t mid = (lo2 + hi2) >> 1;
        if (__ldg(offsets + mid + 1) <= rl) lo2 = mid + 1; else hi2 = mid;
      }
      tile_last[i] = (int32_t)(lo2 < B ? lo2 : B - 1);
    }
  }
}

// poolh: hsum[b,:] (sum over the set's points of the last hidden activations) -> ph[b,:] = rs_b * hsum[b,:] with
// rs = 1/sqrt(n) (sum pooling, deep_sets.py:99) or 1/n (mean, :102); bscale[b] = n * rs_b multiplies the final bias
__global__ void poolh_finalize_kernel(const float* __restrict__ hsum, const int64_t* __restrict__ offsets, int64_t B, int H,
                                      int pooling, float* __restrict__ ph, float* __restrict__ bscale) {
  pdl_enter();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int64_t b = i / H;
  const float n = (float)(offsets[b + 1] - offsets[b]);
  const float rs = (n <= 0.f) ? 0.f : (pooling == PCC_POOL_SUM ? rsqrtf(n) : 1.f / n);
  ph[i] = rs * hsum[i];
  if (i % H == 0) bscale[b] = n * rs;
}

int check_phi_desc(const pcc_phi_desc* d, const char* where) {
  if (!d) return fail(where, "null descriptor");
  // the backward chain keeps z of the last hidden layer in TMEM and recomputes only z_0, which
  // covers phi = Linear(d,H) [+ one H x H hidden layer] + final Linear(H,H)
  if (d->n_layers < 2 || d->n_layers > 3) return fail(where, "fused path needs 1 or 2 hidden phi layers (+ final Linear)");
  if (d->input_dim < 1 || d->input_dim > 7) return fail(where, "fused path needs input_dim <= 7");
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

static inline int q4_of(int d) { return d <= 3 ? 1 : 2; }

template <int ACT, int Q>
static int launch_fwd_pair(const PhiParams& p, cudaStream_t st) {
  const PairSmem lay = pair_smem(Q);
  auto kern = phi_pool_fwd_pair_kernel<ACT, Q>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total);
  if (e != cudaSuccess) return fail("pcc_deepsets_phi_pool_fwd", cudaGetErrorString(e));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntp = (p.num_tiles + 1) / 2;
  int64_t pairs = sms / 2;
  if (pairs > ntp) pairs = ntp;
  {
    ProfScope prof(0, st);
    launch_dep(kern, dim3((unsigned)(2 * pairs)), dim3(kPairThreads), lay.total, st, p);
  }
  return 0;
}

static int g_fwd_pair = -1;  // -1: read PCC_FWD_PAIR on first use; pcc_debug_set_fwd_pair overrides
static bool use_pair_kernel() {
  if (g_fwd_pair < 0) {
    const char* e = getenv("PCC_FWD_PAIR");
    g_fwd_pair = (e && e[0] == '0') ? 0 : 1;
  }
  return g_fwd_pair == 1;
}

template <int H, int ACT, int Q>
static int launch_fwd(const PhiParams& p, cudaStream_t st) {
  if (H == 256 && !p.poolh && p.num_tiles >= 2 && use_pair_kernel()) return launch_fwd_pair<ACT, Q>(p, st);
  const SmemLayout lay = smem_layout(H, p.L, Q, p.poolh);
  auto kern = p.poolh ? phi_pool_fwd_kernel<H, ACT, Q, true> : phi_pool_fwd_kernel<H, ACT, Q, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total);
  if (e != cudaSuccess) return fail("pcc_deepsets_phi_pool_fwd", cudaGetErrorString(e));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(p.num_tiles < sms ? p.num_tiles : sms);
  {
    ProfScope prof(0, st);
    launch_dep(kern, dim3(grid), dim3(kFwdThreads), lay.total, st, p);
  }
  return 0;
}

}  // namespace pcc

using namespace pcc;

extern "C" int pcc_debug_set_trace(void* device_buf_2x4096_i64) {
  pcc::g_trace_buf = device_buf_2x4096_i64;
  return 0;
}

extern "C" int pcc_debug_set_fwd_pair(int on) {
  pcc::g_fwd_pair = on ? 1 : 0;
  return 0;
}

extern "C" int pcc_phi_fused_supported(const pcc_phi_desc* d) { return check_phi_desc(d, __func__); }

extern "C" int64_t pcc_phi_fused_workspace_bytes(const pcc_phi_desc* d, int64_t n, int64_t B) {
  if (check_phi_desc(d, __func__) != 0) return -1;
  const int64_t fwd = ws_layout(d, n, B).total, bwd = phi_bwd_workspace_bytes(d, n, B);
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
  pk.b0 = d->b[0]; pk.w0tab_off = pl.w0tab_off; pk.q4 = q4_of(d->input_dim);

  PhiParams p{};
  p.x = x; p.offsets = offsets; p.n = n; p.B = B; p.num_tiles = cdiv(n, kTileM);
  p.d = d->input_dim; p.L = L; p.pooling = d->pooling; p.res_mask = d->residual_mask;
  p.wpack = (const uint8_t*)wpack;
  p.w0tab = (const float*)((const uint8_t*)wpack + pl.w0tab_off);
  for (int l = 0; l < L; ++l) { p.w_off[l] = pl.w_off[l]; p.bias[l] = d->b[l]; }
  p.pool_acc = wsb + wl.pool_off;
  p.trace = (long long*)g_trace_buf;
  p.tile_first = (const int32_t*)(wsb + wl.tile_first_off);
  p.tile_last = (const int32_t*)(wsb + wl.tile_last_off);
  // sum / mean pooling with an aux buffer: pooling commuted with the final Linear (see the kernel)
  const bool poolh = d->pooling != PCC_POOL_MAX && argmax != nullptr;
  p.poolh = poolh ? 1 : 0;
  launch_dep(fwd_prep_kernel, dim3(32, 2 * L + 3), dim3(256), 0, st, pk, (unsigned long long*)(wsb + wl.pool_off), B * H, offsets, B,
             p.num_tiles, (int32_t*)(wsb + wl.tile_first_off), (int32_t*)(wsb + wl.tile_last_off), n);
  if (p.num_tiles > 0) {
    int rc = 0;
#define PCC_DISPATCH_A(HH, QQ)                                                        \
    switch (d->act) {                                                                 \
      case PCC_ACT_RELU: rc = launch_fwd<HH, PCC_ACT_RELU, QQ>(p, st); break;         \
      case PCC_ACT_GELU: rc = launch_fwd<HH, PCC_ACT_GELU, QQ>(p, st); break;         \
      default: rc = launch_fwd<HH, PCC_ACT_SILU, QQ>(p, st); break;                   \
    }
#define PCC_DISPATCH(HH)                                                              \
    switch (pk.q4) {                                                                  \
      case 1: PCC_DISPATCH_A(HH, 1) break;                                            \
      default: PCC_DISPATCH_A(HH, 2) break;                                           \
    }
    if (H == 256) { PCC_DISPATCH(256) } else { PCC_DISPATCH(128) }
#undef PCC_DISPATCH
#undef PCC_DISPATCH_A
    if (rc != 0) return rc;
  }
  if (B * H > 0 && poolh) {
    // pooled = ph W_{L-1}^T + bscale (x) b_{L-1}  with ph = rs * hsum kept in the aux buffer for the backward
    float* ph = reinterpret_cast<float*>(argmax);
    float* bscale = reinterpret_cast<float*>(wsb + wl.bscale_off);
    launch_dep(poolh_finalize_kernel, dim3((unsigned)cdiv(B * H, 256)), dim3(256), 0, st, (const float*)(wsb + wl.pool_off), offsets, B, H,
               d->pooling, ph, bscale);
    HeadTileParams hp{};
    hp.act = PCC_ACT_RELU;
    HeadTileProb& pr = hp.prob[0];
    pr.A = HeadOperand{ph, nullptr, H, 1, 0, 0, 0};
    pr.B = HeadOperand{d->w[L - 1], nullptr, H, 1, 0, 0, 0};
    pr.I = (int)B; pr.J = H; pr.KK = H;
    pr.C = pooled; pr.ldc = H; pr.bias = d->b[L - 1]; pr.bias_scale = bscale;
    head_set_tiles(pr);
    launch_head_tiles(hp, pr.tiles, st);
  } else if (B * H > 0) {
    launch_dep(pool_finalize_kernel, dim3((unsigned)cdiv(B * H, 256)), dim3(256), 0, st, wsb + wl.pool_off, offsets, d->b[L - 1], B, H,
               d->pooling, pooled, argmax);
  }
  return check_launch(__func__);
}

