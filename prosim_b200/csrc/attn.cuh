// Relative-PE graph attention layer (reference: prosim/models/layers/attention_layer.py:56-118),
// refactored so that no per-edge GEMM and no per-edge [E,128] K/V tensors exist (weights_layout.h, aw::).
//
// Kernels
//   attn_kv_kernel      source rows  : LN_src -> K', V'                 (row-tile GEMM, batched over layers)
//   attn_dstpre_kernel  dest rows    : LN_dst -> q, Qhat[8][128], S, Gx (row-tile GEMM)
//   (edge phase: edge2.cuh)
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

// ------------------------------------------------------------------------------------------------ post
// Per tile of R = 2*RPT destination rows:
//   agg = AggV + Wvr' Rbar ; g = sigmoid(Wga agg + Gx) ; u = agg + g (S - agg) ; o = Wo u + bo
//   x1 = x + LN_post(o) ; y = W2 relu(W1 LN_ffpre(x1) + b1) + b2 ; out = x1 + LN_ffpost(y)
// and, when Wn != nullptr, the NEXT layer's destination-side pre (LN_dst, q, Qhat, S, Gx) on `out`.
template <int RPT>
struct PostSmem {
  static constexpr int R = 2 * RPT;
  static constexpr int LDR = H * D + 4;   // Rbar tile row stride (largest case, zd = 128)
  static constexpr int LDH = 4 * D + 4;   // FFN hidden tile row stride
  static constexpr size_t floats = (size_t)R * LDR + 2 * (size_t)R * LDS_PAD;
  static constexpr size_t bytes = floats * sizeof(float);
};

template <int RPT>
__global__ void __launch_bounds__(256) attn_post_kernel(const float* __restrict__ Xdst, int N, int zd,
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

  const int rw = H * zd;            // Rbar row: [8 heads][zd]
  const int ldr = rw + 4;
  for (int i = threadIdx.x; i < R * (rw / 4); i += 256) {
    int r = i / (rw / 4), c = (i % (rw / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) v = *reinterpret_cast<const float4*>(Rbar + (size_t)(row0 + r) * rw + c);
    *reinterpret_cast<float4*>(sR + r * ldr + c) = v;
  }
  __syncthreads();

  float acc[RPT];
  // 1. agg[r][c] = AggV[r][c] + sum_d WVRGT[d][c] * Rbar[r][c/16][d]
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int row = row0 + rg * RPT + r;
    acc[r] = row < N ? AggV[(size_t)row * D + n] : 0.f;
  }
  gemm_tile_acc<RPT>(acc, sR + (n >> 4) * zd, ldr, zd, W + (zd == 96 ? aw::WVRG96T : aw::WVRGT), D);
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
