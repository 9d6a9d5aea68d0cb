// Relative-PE graph attention layer (reference: prosim/models/layers/attention_layer.py:56-118),
// refactored so that no per-edge GEMM and no per-edge [E,128] K/V tensors exist (weights_layout.h, aw::).
//
// Kernels
//   attn_kv_kernel      source rows  : LN_src -> K', V'                 (row-tile GEMM, batched over layers)
//   attn_dstpre_kernel  dest rows    : LN_dst -> q, Qhat[8][128], S, Gx (row-tile GEMM)
//   attn_edge_kernel    one CTA / destination row: scores, segment softmax, sum_e a_e V'_j and sum_e a_e z_e
//   attn_post_kernel    dest rows    : Wvr-contraction, gate, out-proj, LN, FFN, LN (+ next layer's dstpre fused)
#pragma once
#include "common.cuh"
#include "weights_layout.h"

namespace prosim {

// ------------------------------------------------------------------------------------------------ K', V'
template <int RPT>
__global__ void __launch_bounds__(256) attn_kv_kernel(const float* __restrict__ X, int N, const float* __restrict__ Wbase,
                                                      size_t w_layer_stride, float* __restrict__ KV,
                                                      size_t kv_layer_stride) {
  constexpr int R = 2 * RPT;
  __shared__ __align__(16) float xs[R * LDS_PAD];
  const float* W = Wbase + (size_t)blockIdx.y * w_layer_stride;
  float* kv = KV + (size_t)blockIdx.y * kv_layer_stride;
  const int row0 = blockIdx.x * R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < R; r += 8) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) {
      v = *reinterpret_cast<const float4*>(X + (size_t)(row0 + r) * D + 4 * lane);
      v = ln_row(v, W + aw::LN_SRC_G, W + aw::LN_SRC_B, lane);
    }
    *reinterpret_cast<float4*>(xs + r * LDS_PAD + 4 * lane) = v;
  }
  __syncthreads();
  const int n = threadIdx.x & 127, rg = threadIdx.x >> 7;
  float acc[RPT];
  acc_init(acc, __ldg(W + aw::KB + n));
  gemm_tile_acc<RPT>(acc, xs, LDS_PAD, D, W + aw::WKT, D);
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int row = row0 + rg * RPT + r;
    if (row < N) kv[(size_t)row * 256 + n] = acc[r];
  }
  acc_init(acc, __ldg(W + aw::VB + n));
  gemm_tile_acc<RPT>(acc, xs, LDS_PAD, D, W + aw::WVT, D);
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int row = row0 + rg * RPT + r;
    if (row < N) kv[(size_t)row * 256 + 128 + n] = acc[r];
  }
}

// ------------------------------------------------------------------------------------------------ dst pre
// xd: smem tile of LN_dst-normalised rows; qs: smem scratch tile.  Writes q (pre-scaled), Qhat, S, Gx to global.
template <int RPT>
__device__ __forceinline__ void attn_dst_pre(const float* xd, float* qs, const float* __restrict__ W, int row0, int N,
                                             float* __restrict__ Qg, float* __restrict__ Qhat, float* __restrict__ Sg,
                                             float* __restrict__ Gxg) {
  const int n = threadIdx.x & 127, rg = threadIdx.x >> 7;
  float acc[RPT];
  acc_init(acc, __ldg(W + aw::BQ + n));
  gemm_tile_acc<RPT>(acc, xd, LDS_PAD, D, W + aw::WQT, D);
  acc_store_smem<RPT>(acc, qs, LDS_PAD, false);
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int row = row0 + rg * RPT + r;
    if (row < N) Qg[(size_t)row * D + n] = acc[r];
  }
  acc_init(acc, __ldg(W + aw::BS + n));
  gemm_tile_acc<RPT>(acc, xd, LDS_PAD, D, W + aw::WST, D);
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int row = row0 + rg * RPT + r;
    if (row < N) Sg[(size_t)row * D + n] = acc[r];
  }
  acc_init(acc, __ldg(W + aw::BG + n));
  gemm_tile_acc<RPT>(acc, xd, LDS_PAD, D, W + aw::WGXT, D);
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int row = row0 + rg * RPT + r;
    if (row < N) Gxg[(size_t)row * D + n] = acc[r];
  }
  __syncthreads();  // qs complete
  // Qhat[row][h][d] = sum_c q[row][h*16+c] * WKRG[h*16+c][d]   (thread column n plays d)
#pragma unroll 1
  for (int h = 0; h < H; ++h) {
    acc_init(acc, 0.f);
    gemm_tile_acc<RPT>(acc, qs + h * DH, LDS_PAD, DH, W + aw::WKRG + h * DH * D, D);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      int row = row0 + rg * RPT + r;
      if (row < N) Qhat[(size_t)row * (H * D) + h * D + n] = acc[r];
    }
  }
}

template <int RPT>
__global__ void __launch_bounds__(256) attn_dstpre_kernel(const float* __restrict__ X, int N, const float* __restrict__ W,
                                                          float* __restrict__ Qg, float* __restrict__ Qhat,
                                                          float* __restrict__ Sg, float* __restrict__ Gxg) {
  constexpr int R = 2 * RPT;
  __shared__ __align__(16) float xd[R * LDS_PAD];
  __shared__ __align__(16) float qs[R * LDS_PAD];
  const int row0 = blockIdx.x * R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < R; r += 8) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) {
      v = *reinterpret_cast<const float4*>(X + (size_t)(row0 + r) * D + 4 * lane);
      v = ln_row(v, W + aw::LN_DST_G, W + aw::LN_DST_B, lane);
    }
    *reinterpret_cast<float4*>(xd + r * LDS_PAD + 4 * lane) = v;
  }
  __syncthreads();
  attn_dst_pre<RPT>(xd, qs, W, row0, N, Qg, Qhat, Sg, Gxg);
}

// ------------------------------------------------------------------------------------------------ edges
// One CTA (8 warps) per destination row.  Neighbour list: nbr[row*stride + j], j < deg[row]; the
// normalised relative PE of that edge is Z[(row*stride + j)*128 ..].  KV rows are [K'(128) | V'(128)].
// Dynamic smem: scores [8][sstride] where sstride >= max degree (multiple of 4).
constexpr int EDGE_PART = H * D + D;  // 1152 partial outputs per warp
__global__ void __launch_bounds__(256) attn_edge_kernel(const float* __restrict__ Qg, const float* __restrict__ Qhat,
                                                        const float* __restrict__ KV, const float* __restrict__ Z,
                                                        const int* __restrict__ nbr, const int* __restrict__ deg,
                                                        int stride, int sstride, float* __restrict__ Rbar,
                                                        float* __restrict__ AggV) {
  extern __shared__ __align__(16) float smem[];
  float* sQhat = smem;                    // [8][128]
  float* sQ = sQhat + H * D;              // [128]
  float* sPart = sQ + D;                  // [8 warps][1152]
  float* sS = sPart + 8 * EDGE_PART;      // [8][sstride]
  const int row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_e = min(deg[row], stride);
  const size_t ebase = (size_t)row * stride;

  for (int i = threadIdx.x; i < H * D / 4; i += 256)
    reinterpret_cast<float4*>(sQhat)[i] = reinterpret_cast<const float4*>(Qhat + (size_t)row * H * D)[i];
  if (threadIdx.x < D / 4)
    reinterpret_cast<float4*>(sQ)[threadIdx.x] = reinterpret_cast<const float4*>(Qg + (size_t)row * D)[threadIdx.x];
  __syncthreads();

  // pass A: lane = edge, all 8 head scores in registers
  for (int e = warp * 32 + lane; e < n_e; e += 256) {
    const int j = __ldg(nbr + ebase + e);
    const float4* kp = reinterpret_cast<const float4*>(KV + (size_t)j * 256);
    const float4* zp = reinterpret_cast<const float4*>(Z + (ebase + e) * D);
    float s[H];
#pragma unroll
    for (int h = 0; h < H; ++h) s[h] = 0.f;
#pragma unroll
    for (int h = 0; h < H; ++h) {
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        float4 k4 = __ldg(kp + h * 4 + c4);
        float4 q4 = reinterpret_cast<const float4*>(sQ)[h * 4 + c4];
        s[h] = fmaf(q4.x, k4.x, s[h]);
        s[h] = fmaf(q4.y, k4.y, s[h]);
        s[h] = fmaf(q4.z, k4.z, s[h]);
        s[h] = fmaf(q4.w, k4.w, s[h]);
      }
    }
#pragma unroll 4
    for (int d4 = 0; d4 < D / 4; ++d4) {
      float4 z4 = __ldg(zp + d4);
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float4 q4 = reinterpret_cast<const float4*>(sQhat + h * D)[d4];
        s[h] = fmaf(q4.x, z4.x, s[h]);
        s[h] = fmaf(q4.y, z4.y, s[h]);
        s[h] = fmaf(q4.z, z4.z, s[h]);
        s[h] = fmaf(q4.w, z4.w, s[h]);
      }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) sS[h * sstride + e] = s[h];
  }
  __syncthreads();

  // pass M: warp h normalises head h  (torch_geometric.utils.softmax: exp(s - max) / (sum + 1e-16))
  {
    float* sh = sS + warp * sstride;
    float m = -INFINITY;
    for (int e = lane; e < n_e; e += 32) m = fmaxf(m, sh[e]);
    m = warp_max(m);
    float sum = 0.f;
    for (int e = lane; e < n_e; e += 32) {
      float p = expf(sh[e] - m);
      sh[e] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / (sum + 1e-16f);
    for (int e = lane; e < n_e; e += 32) sh[e] *= inv;
  }
  __syncthreads();

  // pass B: warp takes edges e = warp (mod 8); lane owns 4 columns of every head's Rbar and of AggV
  float4 rb[H];
#pragma unroll
  for (int h = 0; h < H; ++h) rb[h] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
  const int own_h = lane >> 2;
  for (int e = warp; e < n_e; e += 8) {
    const int j = __ldg(nbr + ebase + e);
    const float4 z4 = __ldg(reinterpret_cast<const float4*>(Z + (ebase + e) * D) + lane);
    const float4 v4 = __ldg(reinterpret_cast<const float4*>(KV + (size_t)j * 256 + 128) + lane);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float a = sS[h * sstride + e];
      rb[h].x = fmaf(a, z4.x, rb[h].x);
      rb[h].y = fmaf(a, z4.y, rb[h].y);
      rb[h].z = fmaf(a, z4.z, rb[h].z);
      rb[h].w = fmaf(a, z4.w, rb[h].w);
    }
    const float ao = sS[own_h * sstride + e];
    av.x = fmaf(ao, v4.x, av.x);
    av.y = fmaf(ao, v4.y, av.y);
    av.z = fmaf(ao, v4.z, av.z);
    av.w = fmaf(ao, v4.w, av.w);
  }
  float* part = sPart + warp * EDGE_PART;
#pragma unroll
  for (int h = 0; h < H; ++h) *reinterpret_cast<float4*>(part + h * D + 4 * lane) = rb[h];
  *reinterpret_cast<float4*>(part + H * D + 4 * lane) = av;
  __syncthreads();
  for (int o = threadIdx.x; o < EDGE_PART; o += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sPart[w * EDGE_PART + o];
    if (o < H * D) Rbar[(size_t)row * H * D + o] = t;
    else AggV[(size_t)row * D + (o - H * D)] = t;
  }
}

inline size_t attn_edge_smem_bytes(int sstride) { return sizeof(float) * (size_t)(H * D + D + 8 * EDGE_PART + 8 * sstride); }

// ------------------------------------------------------------------------------------------------ post
// Per tile of R = 2*RPT destination rows:
//   agg = AggV + Wvr' Rbar ; g = sigmoid(Wga agg + Gx) ; u = agg + g (S - agg) ; o = Wo u + bo
//   x1 = x + LN_post(o) ; y = W2 relu(W1 LN_ffpre(x1) + b1) + b2 ; out = x1 + LN_ffpost(y)
// and, when Wn != nullptr, the NEXT layer's destination-side pre (LN_dst, q, Qhat, S, Gx) on `out`.
template <int RPT>
struct PostSmem {
  static constexpr int R = 2 * RPT;
  static constexpr int LDR = H * D + 4;   // Rbar tile row stride
  static constexpr int LDH = 4 * D + 4;   // FFN hidden tile row stride
  static constexpr size_t floats = (size_t)R * LDR + 2 * (size_t)R * LDS_PAD;
  static constexpr size_t bytes = floats * sizeof(float);
};

template <int RPT>
__global__ void __launch_bounds__(256) attn_post_kernel(const float* __restrict__ Xdst, int N,
                                                        const float* __restrict__ Rbar, const float* __restrict__ AggV,
                                                        const float* __restrict__ Sg, const float* __restrict__ Gxg,
                                                        const float* __restrict__ W, float* __restrict__ Out,
                                                        const float* __restrict__ Wn, float* __restrict__ Qg_n,
                                                        float* __restrict__ Qhat_n, float* __restrict__ Sg_n,
                                                        float* __restrict__ Gxg_n) {
  using SM = PostSmem<RPT>;
  constexpr int R = SM::R;
  extern __shared__ __align__(16) float smem[];
  float* sR = smem;                       // [R][LDR]  (later aliased as FFN hidden [R][LDH])
  float* sA = sR + R * SM::LDR;           // [R][LDS_PAD]
  float* sB = sA + R * LDS_PAD;           // [R][LDS_PAD]
  const int row0 = blockIdx.x * R;
  const int n = threadIdx.x & 127, rg = threadIdx.x >> 7;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < R * (H * D / 4); i += 256) {
    int r = i / (H * D / 4), c = (i % (H * D / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) v = *reinterpret_cast<const float4*>(Rbar + (size_t)(row0 + r) * H * D + c);
    *reinterpret_cast<float4*>(sR + r * SM::LDR + c) = v;
  }
  __syncthreads();

  float acc[RPT];
  // 1. agg[r][c] = AggV[r][c] + sum_d WVRGT[d][c] * Rbar[r][c/16][d]
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int row = row0 + rg * RPT + r;
    acc[r] = row < N ? AggV[(size_t)row * D + n] : 0.f;
  }
  gemm_tile_acc<RPT>(acc, sR + (n >> 4) * D, SM::LDR, D, W + aw::WVRGT, D);
  acc_store_smem<RPT>(acc, sA, LDS_PAD, false);
  float agg[RPT];
#pragma unroll
  for (int r = 0; r < RPT; ++r) agg[r] = acc[r];
  __syncthreads();

  // 2. gate and update
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int row = row0 + rg * RPT + r;
    acc[r] = row < N ? Gxg[(size_t)row * D + n] : 0.f;
  }
  gemm_tile_acc<RPT>(acc, sA, LDS_PAD, D, W + aw::WGAT, D);
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int row = row0 + rg * RPT + r;
    float s = row < N ? Sg[(size_t)row * D + n] : 0.f;
    float g = 1.0f / (1.0f + expf(-acc[r]));
    acc[r] = agg[r] + g * (s - agg[r]);
  }
  acc_store_smem<RPT>(acc, sB, LDS_PAD, false);
  __syncthreads();

  // 3. out projection, post-norm, residual
  acc_init(acc, __ldg(W + aw::BO + n));
  gemm_tile_acc<RPT>(acc, sB, LDS_PAD, D, W + aw::WOT, D);
  acc_store_smem<RPT>(acc, sA, LDS_PAD, false);
  __syncthreads();
  for (int r = warp; r < R; r += 8) {
    float4 o = *reinterpret_cast<const float4*>(sA + r * LDS_PAD + 4 * lane);
    o = ln_row(o, W + aw::LN_POST_G, W + aw::LN_POST_B, lane);
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) x = *reinterpret_cast<const float4*>(Xdst + (size_t)(row0 + r) * D + 4 * lane);
    float4 x1 = make_float4(x.x + o.x, x.y + o.y, x.z + o.z, x.w + o.w);
    *reinterpret_cast<float4*>(sB + r * LDS_PAD + 4 * lane) = x1;
    *reinterpret_cast<float4*>(sA + r * LDS_PAD + 4 * lane) = ln_row(x1, W + aw::LN_FFPRE_G, W + aw::LN_FFPRE_B, lane);
  }
  __syncthreads();

  // 4. FFN up (128 -> 512), ReLU, into the (now free) Rbar tile
  float* sH = sR;
#pragma unroll 1
  for (int nb = 0; nb < 4; ++nb) {
    acc_init(acc, __ldg(W + aw::B1 + nb * D + n));
    gemm_tile_acc<RPT>(acc, sA, LDS_PAD, D, W + aw::W1T + nb * D, 4 * D);
    acc_store_smem<RPT>(acc, sH + nb * D, SM::LDH, true);
  }
  __syncthreads();

  // 5. FFN down (512 -> 128), post-norm, residual
  acc_init(acc, __ldg(W + aw::B2 + n));
  gemm_tile_acc<RPT>(acc, sH, SM::LDH, 4 * D, W + aw::W2T, D);
  acc_store_smem<RPT>(acc, sA, LDS_PAD, false);
  __syncthreads();
  for (int r = warp; r < R; r += 8) {
    float4 y = *reinterpret_cast<const float4*>(sA + r * LDS_PAD + 4 * lane);
    y = ln_row(y, W + aw::LN_FFPOST_G, W + aw::LN_FFPOST_B, lane);
    float4 x1 = *reinterpret_cast<const float4*>(sB + r * LDS_PAD + 4 * lane);
    float4 o = make_float4(x1.x + y.x, x1.y + y.y, x1.z + y.z, x1.w + y.w);
    if (row0 + r < N) *reinterpret_cast<float4*>(Out + (size_t)(row0 + r) * D + 4 * lane) = o;
    if (Wn != nullptr)
      *reinterpret_cast<float4*>(sA + r * LDS_PAD + 4 * lane) = ln_row(o, Wn + aw::LN_DST_G, Wn + aw::LN_DST_B, lane);
  }
  if (Wn == nullptr) return;
  __syncthreads();
  // 6. next layer's destination-side pre on the fresh rows (saves a launch and a re-read per layer)
  attn_dst_pre<RPT>(sA, sB, Wn, row0, N, Qg_n, Qhat_n, Sg_n, Gxg_n);
}

}  // namespace prosim
