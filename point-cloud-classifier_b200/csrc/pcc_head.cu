// rho head (set encoder): [Linear + act] x (0..3) + Linear, M = batch rows.
//   reference: /root/reference/models/deep_sets.py:112 (self.rho(pooled)) with the layer stack of :59-72
//   (no LayerNorm variant) and its autograd.
// The head is tiny (0.07 GFLOP per train step at the yaml shape), so what matters is the LENGTH of the
// dependent chain, not throughput.  An earlier version ran one CTA per 4 rows through all layers in a
// single launch: every CTA streamed every weight matrix through a latency-bound loop and the weight
// gradients needed ~1M vector atomics (16 us forward, 30 us backward, warm).  Here every layer is ONE
// launch of 32x32 output tiles spread over the whole chip (a 256x256x256 layer = 64 CTAs, K loop of 4
// tiles), with all elementwise work folded into the operand loads:
//   forward  layer l : z_l = f(in) W_l^T + b_l          f = act for l > 0 (input is the saved z_{l-1})
//   backward layer l : dz_l = g_l * act'(z_l) built on load (g_L-1 = dy), and in the SAME launch
//                        dW_l = dz_l^T in_l , db_l = colsum(dz_l)      (tiles over [N_l, K_l])
//                        g_{l-1} = dz_l W_l                            (tiles over [M,   K_l])
// No atomics, no zero-fill launches, bitwise deterministic.
#include "pcc_head.cuh"

namespace pcc {

constexpr int kHeadMaxDim = 1024;
constexpr int kHeadMaxLayers = 4;
constexpr int kHT = 32;    // output tile edge
constexpr int kHK = 64;    // contraction tile
constexpr int kHThreads = 256;

// The raw loads of a K tile are issued as one unconditional batch (memory-level parallelism: the loop is
// latency bound); the elementwise transforms run afterwards, with the mode tests hoisted out of the loops.
// Math: 256 threads = 4 k-groups x 64 threads; a group covers the whole 32x32 tile with 4x4 register
// micro-tiles (two LDS.128 per 16 FMA) over a quarter of every K tile; the four partial tiles are summed
// through shared memory at the end (fixed order).  Operand staging is bank-conflict free both ways: rows
// of 36 floats, lanes laid out 8 (kk) x 4 (i) when the operand is contiguous along kk.
template <int ACT>
__global__ void __launch_bounds__(kHThreads) head_tile_kernel(const HeadTileParams hp) {
  constexpr int S = kHT + 4;
  __shared__ __align__(16) float smem_f[2 * kHK * S];
  float (*As)[S] = reinterpret_cast<float (*)[S]>(smem_f);
  float (*Bs)[S] = reinterpret_cast<float (*)[S]>(smem_f + kHK * S);
  int t = blockIdx.x;
  const bool second = t >= hp.prob[0].tiles;
  if (second) t -= hp.prob[0].tiles;
  const HeadTileProb& pr = second ? hp.prob[1] : hp.prob[0];
  const int I = pr.I, J = pr.J, KK = pr.KK;
  const int i0 = (t / pr.tiles_j) * kHT, j0 = (t % pr.tiles_j) * kHT;
  const int tid = threadIdx.x;
  const int grp = tid >> 6, tyq = (tid & 63) >> 3, txq = tid & 7;  // group: rows 4 tyq.., columns 4 txq..
  const float* __restrict__ ap = pr.A.p;
  const float* __restrict__ aq = pr.A.q;
  const float* __restrict__ bp = pr.B.p;
  const int64_t a_si = pr.A.si, a_sk = pr.A.sk, a_qi = pr.A.qi, a_qk = pr.A.qk, b_si = pr.B.si, b_sk = pr.B.sk;
  const int a_mode = pr.A.mode, b_mode = pr.B.mode;
  const bool a_kc = a_sk == 1, b_kc = b_sk == 1;  // contiguous along kk -> lanes run along kk
  constexpr int PER = kHT * kHK / kHThreads;       // 8 elements per thread per operand
  // element e (0..7) of this thread inside a 32 (i) x 64 (kk) operand tile: contiguous-along-kk operands use
  // lanes 8 (kk) x 4 (i): i = ((tid >> 3) & 3) + 4 e, kk = ((tid >> 5) << 3 | (tid & 7)); the others use lanes
  // along i: i = tid & 31, kk = (tid >> 5) + 8 e.  Global and shared addresses are base + e * step.
  const int ai_f = a_kc ? ((tid >> 3) & 3) : (tid & 31), ai_s = a_kc ? 4 : 0;
  const int ak_f = a_kc ? (((tid >> 5) << 3) | (tid & 7)) : (tid >> 5), ak_s = a_kc ? 0 : 8;
  const int bi_f = b_kc ? ((tid >> 3) & 3) : (tid & 31), bi_s = b_kc ? 4 : 0;
  const int bk_f = b_kc ? (((tid >> 5) << 3) | (tid & 7)) : (tid >> 5), bk_s = b_kc ? 0 : 8;
  const int64_t a_step_e = ai_s * a_si + ak_s * a_sk, a_step_t = kHK * a_sk;
  const int64_t q_step_e = ai_s * a_qi + ak_s * a_qk, q_step_t = kHK * a_qk;
  const int64_t b_step_e = bi_s * b_si + bk_s * b_sk, b_step_t = kHK * b_sk;
  float* const a_sts = &As[ak_f][ai_f];
  float* const b_sts = &Bs[bk_f][bi_f];
  const int a_sstep = ak_s * S + ai_s, b_sstep = bk_s * S + bi_s;

  // All loads of a 256-wide K range (4 staging tiles) are in flight at once: a dependent global round trip
  // costs more than the math of a whole tile, so the K loop must not serialise them.
  constexpr int NT = 4;
  float ra[NT][PER], rq[NT][PER], rb[NT][PER];
  auto fetch = [&](int kbase) {
    const float* pa = ap + (i0 + ai_f) * a_si + (kbase + ak_f) * a_sk;
    const float* pb = bp + (j0 + bi_f) * b_si + (kbase + bk_f) * b_sk;
    const float* pq = aq + (i0 + ai_f) * a_qi + (kbase + ak_f) * a_qk;
    if (i0 + kHT <= I && j0 + kHT <= J && kbase + NT * kHK <= KK) {  // interior: no predicates
#pragma unroll
      for (int tl = 0; tl < NT; ++tl)
#pragma unroll
        for (int e = 0; e < PER; ++e) {
          ra[tl][e] = __ldg(pa + tl * a_step_t + e * a_step_e);
          rb[tl][e] = __ldg(pb + tl * b_step_t + e * b_step_e);
        }
      if (a_mode == 2) {
#pragma unroll
        for (int tl = 0; tl < NT; ++tl)
#pragma unroll
          for (int e = 0; e < PER; ++e) rq[tl][e] = __ldg(pq + tl * q_step_t + e * q_step_e);
      }
    } else {
#pragma unroll
      for (int tl = 0; tl < NT; ++tl)
#pragma unroll
        for (int e = 0; e < PER; ++e) {
          const bool oka = (i0 + ai_f + e * ai_s < I) && (kbase + tl * kHK + ak_f + e * ak_s < KK);
          const bool okb = (j0 + bi_f + e * bi_s < J) && (kbase + tl * kHK + bk_f + e * bk_s < KK);
          ra[tl][e] = oka ? __ldg(pa + tl * a_step_t + e * a_step_e) : 0.f;
          rb[tl][e] = okb ? __ldg(pb + tl * b_step_t + e * b_step_e) : 0.f;
          rq[tl][e] = (oka && a_mode == 2) ? __ldg(pq + tl * q_step_t + e * q_step_e) : 0.f;
        }
    }
  };
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float csum = 0.f;
  const bool want_colsum = pr.colsum != nullptr && j0 == 0;

  for (int kbase = 0; kbase < KK; kbase += NT * kHK) {
    fetch(kbase);
#pragma unroll
    for (int tl = 0; tl < NT; ++tl) {
      if (kbase + tl * kHK >= KK) break;
      if (a_mode == 1) {
#pragma unroll
        for (int e = 0; e < PER; ++e) ra[tl][e] = act_fwd(ACT, ra[tl][e]);
      } else if (a_mode == 2) {
#pragma unroll
        for (int e = 0; e < PER; ++e) ra[tl][e] *= act_grad(ACT, rq[tl][e]);
      }
      if (b_mode == 1) {
#pragma unroll
        for (int e = 0; e < PER; ++e) rb[tl][e] = act_fwd(ACT, rb[tl][e]);
      }
      if (tl > 0 || kbase > 0) __syncthreads();  // the previous tile has been consumed
#pragma unroll
      for (int e = 0; e < PER; ++e) {
        a_sts[e * a_sstep] = ra[tl][e];
        b_sts[e * b_sstep] = rb[tl][e];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kHK / 4; ++kk) {
        const int k = grp * (kHK / 4) + kk;
        const float4 a = *reinterpret_cast<const float4*>(&As[k][tyq * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][txq * 4]);
        acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
        acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
        acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
        acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
        acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
        acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
        acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
        acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
      }
      if (want_colsum && tid < kHT) {
        if (pr.colsum_w) {
          const int kb = kbase + tl * kHK;
#pragma unroll 16
          for (int k = 0; k < kHK; ++k) csum += As[k][tid] * (kb + k < KK ? __ldg(pr.colsum_w + kb + k) : 0.f);
        } else {
#pragma unroll 16
          for (int k = 0; k < kHK; ++k) csum += As[k][tid];
        }
      }
    }
  }
  __syncthreads();

  // ---- sum the four k-group partials (reusing the staging memory: 4 x 32 x 33 floats) and store
  float (*red)[kHT][kHT + 1] = reinterpret_cast<float (*)[kHT][kHT + 1]>(smem_f);
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) red[grp][tyq * 4 + a][txq * 4 + b] = acc[a][b];
  __syncthreads();
  {
    const int li = tid >> 3, lj = (tid & 7) * 4;
    const int i = i0 + li;
    if (i < I) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int j = j0 + lj + b;
        if (j >= J) continue;
        float v = (red[0][li][lj + b] + red[1][li][lj + b]) + (red[2][li][lj + b] + red[3][li][lj + b]);
        if (pr.row_scale) v *= __ldg(pr.row_scale + i);
        if (pr.bias) v += (pr.bias_scale ? __ldg(pr.bias_scale + i) : 1.f) * __ldg(pr.bias + j);
        pr.C[(int64_t)i * pr.ldc + j] = v;
      }
    }
  }
  if (want_colsum && tid < kHT && i0 + tid < I) pr.colsum[i0 + tid] = csum;
}

void launch_head_tiles(const HeadTileParams& hp, int tiles, cudaStream_t st) {
  if (tiles <= 0) return;
  auto kern = head_tile_kernel<PCC_ACT_TANH>;
  switch (hp.act) {
    case PCC_ACT_RELU: kern = head_tile_kernel<PCC_ACT_RELU>; break;
    case PCC_ACT_GELU: kern = head_tile_kernel<PCC_ACT_GELU>; break;
    case PCC_ACT_SILU: kern = head_tile_kernel<PCC_ACT_SILU>; break;
    default: break;
  }
  PCC_K(kern)<<<tiles, kHThreads, 0, st>>>(hp);
}

static int check_head(const pcc_head_desc* d, const char* where) {
  if (!d) return fail(where, "null descriptor");
  if (d->n_layers < 1 || d->n_layers > kHeadMaxLayers) return fail(where, "head needs 1..4 layers");
  for (int l = 0; l <= d->n_layers; ++l)
    if (d->dims[l] < 1 || d->dims[l] > kHeadMaxDim) return fail(where, "head widths must be in [1,1024]");
  if (d->act != PCC_ACT_RELU && d->act != PCC_ACT_GELU && d->act != PCC_ACT_SILU && d->act != PCC_ACT_TANH)
    return fail(where, "head activation must be relu/gelu/silu/tanh");
  return 0;
}

struct HeadGeom {
  int zoff[kHeadMaxLayers];
  int zwidth, maxhid;
};
static HeadGeom head_geom(const pcc_head_desc* d) {
  HeadGeom g{};
  int off = 0, mh = 1;
  for (int l = 0; l < d->n_layers - 1; ++l) {
    g.zoff[l] = off;
    off += d->dims[l + 1];
    if (d->dims[l + 1] > mh) mh = d->dims[l + 1];
  }
  g.zwidth = off;
  g.maxhid = mh;
  return g;
}

void head_set_tiles(HeadTileProb& p) {
  p.tiles_j = (int)cdiv(p.J, kHT);
  p.tiles = (int)cdiv(p.I, kHT) * p.tiles_j;
}
static void set_tiles(HeadTileProb& p) { head_set_tiles(p); }

}  // namespace pcc

using namespace pcc;

extern "C" int pcc_mlp_head_supported(const pcc_head_desc* d) { return check_head(d, __func__); }

extern "C" int64_t pcc_mlp_head_workspace_bytes(const pcc_head_desc* d, int64_t M) {
  if (check_head(d, __func__) != 0) return -1;
  const HeadGeom g = head_geom(d);
  return 2 * (M > 0 ? M : 1) * (int64_t)g.maxhid * (int64_t)sizeof(float);
}

extern "C" int pcc_mlp_head_fwd(const pcc_head_desc* d, const float* x, float* y, float* zsave, int64_t M, int device,
                                void* stream) {
  PCC_ENTER(device);
  if (check_head(d, __func__) != 0) return -1;
  PCC_REQUIRE(M < (int64_t)0x7fffffff, "row count exceeds int32");
  if (M == 0) return 0;
  const HeadGeom g = head_geom(d);
  const int L = d->n_layers;
  for (int l = 0; l < L; ++l) {
    const int K = d->dims[l], N = d->dims[l + 1];
    HeadTileParams hp{};
    hp.act = d->act;
    HeadTileProb& p = hp.prob[0];
    if (l == 0) p.A = HeadOperand{x, nullptr, K, 1, 0, 0, 0};
    else p.A = HeadOperand{zsave + g.zoff[l - 1], nullptr, g.zwidth, 1, 0, 0, 1};
    p.B = HeadOperand{d->w[l], nullptr, K, 1, 0, 0, 0};
    p.I = (int)M; p.J = N; p.KK = K;
    if (l < L - 1) { p.C = zsave + g.zoff[l]; p.ldc = g.zwidth; } else { p.C = y; p.ldc = N; }
    p.bias = d->b[l];
    set_tiles(p);
    launch_head_tiles(hp, p.tiles, (cudaStream_t)stream);
  }
  return check_launch(__func__);
}

extern "C" int pcc_mlp_head_bwd(const pcc_head_desc* d, const float* x, const float* zsave, const float* dy, float* dx,
                                float* const* dw, float* const* db, void* ws, int64_t M, int device, void* stream) {
  PCC_ENTER(device);
  if (check_head(d, __func__) != 0) return -1;
  PCC_REQUIRE(M < (int64_t)0x7fffffff, "row count exceeds int32");
  const HeadGeom g = head_geom(d);
  const int L = d->n_layers;
  PCC_REQUIRE(L == 1 || ws != nullptr || M == 0, "workspace required (pcc_mlp_head_workspace_bytes)");
  float* gbuf[2] = {(float*)ws, (float*)ws + (M > 0 ? M : 1) * (int64_t)g.maxhid};
  for (int l = L - 1; l >= 0; --l) {
    const int K = d->dims[l], N = d->dims[l + 1];
    HeadTileParams hp{};
    hp.act = d->act;
    // dz_l(r, u): dy for the final layer, g_l * act'(z_l) for hidden layers
    const float* gsrc = (l == L - 1) ? dy : gbuf[l & 1];
    const int64_t gld = N;
    const float* zq = (l == L - 1) ? nullptr : zsave + g.zoff[l];
    const int dzmode = (l == L - 1) ? 0 : 2;
    // ---- dW_l[u][k] = sum_r dz(r,u) in(r,k); db_l[u] = sum_r dz(r,u)
    HeadTileProb& pw = hp.prob[0];
    pw.A = HeadOperand{gsrc, zq, 1, gld, 1, g.zwidth, dzmode};
    if (l == 0) pw.B = HeadOperand{x, nullptr, 1, K, 0, 0, 0};
    else pw.B = HeadOperand{zsave + g.zoff[l - 1], nullptr, 1, g.zwidth, 0, 0, 1};
    pw.I = N; pw.J = K; pw.KK = (int)M;
    pw.C = dw[l]; pw.ldc = K; pw.bias = nullptr; pw.colsum = db[l];
    set_tiles(pw);
    // ---- g_{l-1}[r][k] = sum_u dz(r,u) W_l[u][k]
    HeadTileProb& pd = hp.prob[1];
    float* dst = (l == 0) ? dx : gbuf[(l - 1) & 1];
    if (dst != nullptr && M > 0) {
      pd.A = HeadOperand{gsrc, zq, gld, 1, g.zwidth, 1, dzmode};
      pd.B = HeadOperand{d->w[l], nullptr, 1, K, 0, 0, 0};
      pd.I = (int)M; pd.J = K; pd.KK = N;
      pd.C = dst; pd.ldc = K;
      set_tiles(pd);
    }
    launch_head_tiles(hp, pw.tiles + pd.tiles, (cudaStream_t)stream);
  }
  return check_launch(__func__);
}
