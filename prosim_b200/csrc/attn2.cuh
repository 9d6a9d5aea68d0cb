// Throughput variants of the node-level attention kernels (attn.cuh) built on gemm_tile.cuh: 16*RT rows per
// CTA, weights streamed once per CTA through a cp.async ring.  Same math, same summation order per output
// (ascending k), so results are bit-identical to the v1 kernels; api.cu picks v2 for large row counts and v1
// (2..16 rows per CTA, more CTAs) for small latency-bound launches.
#pragma once
#include "attn.cuh"
#include "gemm_tile.cuh"

namespace prosim {

// ------------------------------------------------------------------------------------------------ K', V'
template <int TR, int NW>
struct Kv2Smem {
  static constexpr int M = NW * TR;
  static constexpr size_t bytes = WPIPE_BYTES + (size_t)M * LDS_PAD * sizeof(float);
};

template <int TR, int NW>
__global__ void __launch_bounds__(NW * 32) attn_kv2_kernel(const float* __restrict__ X, int N, const float* __restrict__ Wbase,
                                                       size_t w_layer_stride, float* __restrict__ KV,
                                                       size_t kv_layer_stride) {
  constexpr int M = NW * TR;
  extern __shared__ __align__(16) float smem[];
  const float* W = Wbase + (size_t)blockIdx.y * w_layer_stride;
  WPipe p = wpipe_init<NW>(smem, [&](WSeg* sg) {
    sg[0] = WSeg{W + aw::WKT, D, D};
    sg[1] = WSeg{W + aw::WVT, D, D};
    return 2;
  });
  float* xs = smem + WPIPE_BYTES / sizeof(float);
  float* kv = KV + (size_t)blockIdx.y * kv_layer_stride;
  const int row0 = blockIdx.x * M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < M; r += NW) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) {
      v = *reinterpret_cast<const float4*>(X + (size_t)(row0 + r) * D + 4 * lane);
      v = ln_row(v, W + aw::LN_SRC_G, W + aw::LN_SRC_B, lane);
    }
    *reinterpret_cast<float4*>(xs + r * LDS_PAD + 4 * lane) = v;
  }
  float acc[TR][4];
  acc2_init_bias<TR>(acc, W + aw::KB);
  gemm2<TR, NW>(acc, xs, LDS_PAD, p);
  acc2_store_global<TR>(acc, kv, 256, 0, row0, N);
  acc2_init_bias<TR>(acc, W + aw::VB);
  gemm2<TR, NW>(acc, xs, LDS_PAD, p);
  acc2_store_global<TR>(acc, kv, 256, 128, row0, N);
}

// ------------------------------------------------------------------------------------------------ dst pre
// Weight segments of the destination-side projections, in the order attn_dst_pre2 consumes them.
__device__ __forceinline__ int dst_pre_segments(WSeg* sg, const float* W) {
  sg[0] = WSeg{W + aw::WQT, D, D};
  sg[1] = WSeg{W + aw::WST, D, D};
  sg[2] = WSeg{W + aw::WGXT, D, D};
  for (int h = 0; h < H; ++h) sg[3 + h] = WSeg{W + aw::WKRG + h * DH * D, D, DH};
  return 3 + H;
}

// xd: LN_dst-normalised tile in smem; sq: scratch tile.  The pipe's next segments must be dst_pre_segments(W).
template <int TR, int NW>
__device__ __forceinline__ void attn_dst_pre2(const float* xd, float* sq, const float* __restrict__ W, int row0, int N,
                                              float* __restrict__ Qg, float* __restrict__ Qhat, float* __restrict__ Sg,
                                              float* __restrict__ Gxg, WPipe& p) {
  float acc[TR][4];
  acc2_init_bias<TR>(acc, W + aw::BQ);
  gemm2<TR, NW>(acc, xd, LDS_PAD, p);
  acc2_store_smem<TR>(acc, sq, LDS_PAD, false);
  acc2_store_global<TR>(acc, Qg, D, 0, row0, N);
  acc2_init_bias<TR>(acc, W + aw::BS);
  gemm2<TR, NW>(acc, xd, LDS_PAD, p);
  acc2_store_global<TR>(acc, Sg, D, 0, row0, N);
  acc2_init_bias<TR>(acc, W + aw::BG);
  gemm2<TR, NW>(acc, xd, LDS_PAD, p);
  acc2_store_global<TR>(acc, Gxg, D, 0, row0, N);
#pragma unroll 1
  for (int h = 0; h < H; ++h) {
    acc2_init<TR>(acc, 0.f);
    gemm2<TR, NW>(acc, sq + h * DH, LDS_PAD, p);
    acc2_store_global<TR>(acc, Qhat, H * D, h * D, row0, N);
  }
}

template <int TR, int NW>
struct Pre2Smem {
  static constexpr int M = NW * TR;
  static constexpr size_t bytes = WPIPE_BYTES + 2 * (size_t)M * LDS_PAD * sizeof(float);
};

template <int TR, int NW>
__global__ void __launch_bounds__(NW * 32) attn_dstpre2_kernel(const float* __restrict__ X, int N, const float* __restrict__ W,
                                                           float* __restrict__ Qg, float* __restrict__ Qhat,
                                                           float* __restrict__ Sg, float* __restrict__ Gxg) {
  constexpr int M = NW * TR;
  extern __shared__ __align__(16) float smem[];
  WPipe p = wpipe_init<NW>(smem, [&](WSeg* sg) { return dst_pre_segments(sg, W); });
  float* xd = smem + WPIPE_BYTES / sizeof(float);
  float* sq = xd + M * LDS_PAD;
  const int row0 = blockIdx.x * M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < M; r += NW) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) {
      v = *reinterpret_cast<const float4*>(X + (size_t)(row0 + r) * D + 4 * lane);
      v = ln_row(v, W + aw::LN_DST_G, W + aw::LN_DST_B, lane);
    }
    *reinterpret_cast<float4*>(xd + r * LDS_PAD + 4 * lane) = v;
  }
  attn_dst_pre2<TR, NW>(xd, sq, W, row0, N, Qg, Qhat, Sg, Gxg, p);
}

// ------------------------------------------------------------------------------------------------ post
template <int TR, int NW>
struct Post2Smem {
  static constexpr int M = NW * TR;
  static constexpr int LDR = H * 96 + 4;      // the v2 kernel serves 96-wide z only (zd = 128 graphs are small: v1)
  static constexpr int LDH = 4 * D + 4;
  static constexpr size_t bytes = WPIPE_BYTES + ((size_t)M * LDR + 2 * (size_t)M * LDS_PAD) * sizeof(float);
};

template <int TR, int NW>
__global__ void __launch_bounds__(NW * 32, 1) attn_post2_kernel(const float* __restrict__ Xdst, int N, int zd,
                                                            const float* __restrict__ Rbar, const float* __restrict__ AggV,
                                                            const float* __restrict__ Sg, const float* __restrict__ Gxg,
                                                            const float* __restrict__ W, float* __restrict__ Out,
                                                            const float* __restrict__ Wn, float* __restrict__ Qg_n,
                                                            float* __restrict__ Qhat_n, float* __restrict__ Sg_n,
                                                            float* __restrict__ Gxg_n) {
  using SM = Post2Smem<TR, NW>;
  constexpr int M = SM::M;
  extern __shared__ __align__(16) float smem[];
  const float* Wvr = W + (zd == 96 ? aw::WVRG96T : aw::WVRGT);
  WPipe p = wpipe_init<NW>(smem, [&](WSeg* sg) {
    int n = 0;
    sg[n++] = WSeg{Wvr, D, zd};
    sg[n++] = WSeg{W + aw::WGAT, D, D};
    sg[n++] = WSeg{W + aw::WOT, D, D};
    for (int nb = 0; nb < 4; ++nb) sg[n++] = WSeg{W + aw::W1T + nb * D, 4 * D, D};
    sg[n++] = WSeg{W + aw::W2T, D, 4 * D};
    if (Wn != nullptr) n += dst_pre_segments(sg + n, Wn);
    return n;
  });
  float* sR = smem + WPIPE_BYTES / sizeof(float);   // [M][LDR], later the FFN hidden tile [M][LDH]
  float* sA = sR + M * SM::LDR;
  float* sB = sA + M * LDS_PAD;
  const int row0 = blockIdx.x * M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // the first weight chunks fly while the Rbar tile is loaded
  const int rw = H * zd, ldr = rw + 4;            // Rbar row: [8 heads][zd]
  for (int i = threadIdx.x; i < M * (rw / 4); i += NW * 32) {
    const int r = i / (rw / 4), c = (i % (rw / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) v = __ldg(reinterpret_cast<const float4*>(Rbar + (size_t)(row0 + r) * rw + c));
    *reinterpret_cast<float4*>(sR + r * ldr + c) = v;
  }

  float acc[TR][4], agg[TR][4];
  // 1. agg = AggV + Wvr' Rbar (block diagonal: the A row of an output column is the Rbar row of its head)
  acc2_load_global<TR>(acc, AggV, D, row0, N);
  gemm2<TR, NW>(acc, sR + (tile_coord<TR>().col >> 4) * zd, ldr, p);
  acc2_store_smem<TR>(acc, sA, LDS_PAD, false);
#pragma unroll
  for (int r = 0; r < TR; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) agg[r][c] = acc[r][c];

  // 2. gate: g = sigmoid(Wga agg + Gx) ; u = agg + g (S - agg)
  acc2_load_global<TR>(acc, Gxg, D, row0, N);
  gemm2<TR, NW>(acc, sA, LDS_PAD, p);
  {
    float s[TR][4];
    acc2_load_global<TR>(s, Sg, D, row0, N);
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float g = 1.0f / (1.0f + expf(-acc[r][c]));
        acc[r][c] = agg[r][c] + g * (s[r][c] - agg[r][c]);
      }
  }
  acc2_store_smem<TR>(acc, sB, LDS_PAD, false);

  // 3. out projection, post-norm, residual, FFN pre-norm
  acc2_init_bias<TR>(acc, W + aw::BO);
  gemm2<TR, NW>(acc, sB, LDS_PAD, p);
  acc2_store_smem<TR>(acc, sA, LDS_PAD, false);
  __syncthreads();
  for (int r = warp; r < M; r += NW) {
    float4 o = *reinterpret_cast<const float4*>(sA + r * LDS_PAD + 4 * lane);
    o = ln_row(o, W + aw::LN_POST_G, W + aw::LN_POST_B, lane);
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) x = *reinterpret_cast<const float4*>(Xdst + (size_t)(row0 + r) * D + 4 * lane);
    const float4 x1 = make_float4(x.x + o.x, x.y + o.y, x.z + o.z, x.w + o.w);
    *reinterpret_cast<float4*>(sB + r * LDS_PAD + 4 * lane) = x1;
    *reinterpret_cast<float4*>(sA + r * LDS_PAD + 4 * lane) = ln_row(x1, W + aw::LN_FFPRE_G, W + aw::LN_FFPRE_B, lane);
  }

  // 4. FFN up, ReLU, into the (free) Rbar tile
  float* sH = sR;
#pragma unroll 1
  for (int nb = 0; nb < 4; ++nb) {
    acc2_init_bias<TR>(acc, W + aw::B1 + nb * D);
    gemm2<TR, NW>(acc, sA, LDS_PAD, p);
    acc2_store_smem<TR>(acc, sH + nb * D, SM::LDH, true);
  }

  // 5. FFN down, post-norm, residual
  acc2_init_bias<TR>(acc, W + aw::B2);
  gemm2<TR, NW>(acc, sH, SM::LDH, p);
  acc2_store_smem<TR>(acc, sA, LDS_PAD, false);
  __syncthreads();
  for (int r = warp; r < M; r += NW) {
    float4 y = *reinterpret_cast<const float4*>(sA + r * LDS_PAD + 4 * lane);
    y = ln_row(y, W + aw::LN_FFPOST_G, W + aw::LN_FFPOST_B, lane);
    const float4 x1 = *reinterpret_cast<const float4*>(sB + r * LDS_PAD + 4 * lane);
    const float4 o = make_float4(x1.x + y.x, x1.y + y.y, x1.z + y.z, x1.w + y.w);
    if (row0 + r < N) *reinterpret_cast<float4*>(Out + (size_t)(row0 + r) * D + 4 * lane) = o;
    if (Wn != nullptr)
      *reinterpret_cast<float4*>(sA + r * LDS_PAD + 4 * lane) = ln_row(o, Wn + aw::LN_DST_G, Wn + aw::LN_DST_B, lane);
  }
  if (Wn == nullptr) return;
  // 6. next layer's destination-side projections on the fresh rows (their first weight chunks are already in flight)
  attn_dst_pre2<TR, NW>(sA, sB, Wn, row0, N, Qg_n, Qhat_n, Sg_n, Gxg_n, p);
}

}  // namespace prosim
