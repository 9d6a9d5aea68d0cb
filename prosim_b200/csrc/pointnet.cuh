// PointNet polyline encoder (reference: prosim/models/scene_encoder/pointnet_encoder.py:24-62 with the
// obs / map configurations of obs_encoder.py:75-86 and map_encoder.py:67-88, and the drag-point condition encoder
// condition_transformer/condition_encoders.py:147-191).
//   pre_mlps (per point) -> max-pool over the polyline -> [point | pooled] -> mlps -> max-pool -> out_mlps
// Masked points contribute zeros to both max-pools (the reference scatters valid rows into a zero buffer);
// since every pooled tensor is a ReLU output this equals a max over valid points clamped at 0.
// The 256-wide first `mlps` layer is split into its point half and its pooled half so the pooled half is
// evaluated once per polyline instead of once per point.
// One CTA = G polylines (G*P <= 64 point rows), all activations stay in shared memory.
#pragma once
#include "common.cuh"
#include "gemm_tile.cuh"
#include "weights_layout.h"

namespace prosim {

template <int P>
struct PointNetCfg {
  static constexpr int G = 64 / P;
  static constexpr int ROWS = 64;
  static constexpr size_t smem_floats = 2 * (size_t)ROWS * LDS_PAD + 2 * 8 * LDS_PAD + ROWS;
  static constexpr size_t smem_bytes = smem_floats * sizeof(float);
};

// X: [n_all][P][IN] raw inputs, mask: uint8 [n_all][P][mask_inner] (point valid = all mask_inner bytes != 0);
// mask == nullptr: validity is derived from the data (no NaN feature) -- the drag-point condition encoder
// rows: [n_poly] indices into n_all (compaction list of valid polylines); Out: [n_poly][128]
template <int IN, int IN_PAD, int NPRE, int P>
__global__ void __launch_bounds__(256) pointnet_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask,
                                                       int mask_inner, const int* __restrict__ rows, int n_poly,
                                                       const float* __restrict__ W, float* __restrict__ Out) {
  using C = PointNetCfg<P>;
  constexpr int G = C::G, ROWS = C::ROWS, RPT = 16;   // gemm_tile_acc2: 4 row groups x 16 rows, 64 column pairs
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem;                         // [64][132]
  float* bufB = bufA + ROWS * LDS_PAD;        // [64][132]
  float* sPool = bufB + ROWS * LDS_PAD;       // [8][132]
  float* sPP = sPool + 8 * LDS_PAD;           // [8][132]
  int* sValid = reinterpret_cast<int*>(sPP + 8 * LDS_PAD);  // [64]
  const int poly0 = blockIdx.x * G;
  const int n = threadIdx.x & 127, rg = threadIdx.x >> 7;   // small pooled-row GEMMs (gemm_tile_acc<4>)
  const int cp = threadIdx.x & 63, rg2 = threadIdx.x >> 6;  // per-point GEMMs (gemm_tile_acc2<16>)

  for (int r = threadIdx.x; r < ROWS; r += 256) {
    int g = r / P, p = r % P, v = 0;
    if (g < G && poly0 + g < n_poly) {
      v = 1;
      if (mask != nullptr) {
        const uint8_t* m = mask + ((size_t)rows[poly0 + g] * P + p) * mask_inner;
        for (int i = 0; i < mask_inner; ++i) v &= (m[i] != 0);
      } else {   // no mask tensor: a point is valid when none of its features is NaN (condition_encoders.py:178)
        const float* xp = X + ((size_t)rows[poly0 + g] * P + p) * IN;
        for (int i = 0; i < IN; ++i) v &= !isnan(xp[i]);
      }
    }
    sValid[r] = v;
  }
  for (int i = threadIdx.x; i < 8 * LDS_PAD; i += 256) sPool[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < ROWS * IN_PAD; i += 256) {
    int r = i / IN_PAD, c = i % IN_PAD;
    float v = 0.f;
    if (sValid[r] && c < IN) v = X[((size_t)rows[poly0 + r / P] * P + (r % P)) * IN + c];
    bufA[r * LDS_PAD + c] = v;
  }
  __syncthreads();

  float2 acc[RPT];
  auto init_bias = [&](const float* b) {
    const float2 bb = __ldg(reinterpret_cast<const float2*>(b) + cp);
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r] = bb;
  };
  auto store = [&](float* dst, bool relu, bool masked) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int row = rg2 * RPT + r;
      float2 v = acc[r];
      if (relu) v = make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
      if (masked && !sValid[row]) v = make_float2(0.f, 0.f);
      *reinterpret_cast<float2*>(dst + row * LDS_PAD + 2 * cp) = v;
    }
  };
  auto store_masked = [&](float* dst, bool relu) { store(dst, relu, true); };

  // ---- pre_mlps
  init_bias(W + pw::PRE0_B);
  gemm_tile_acc2<RPT>(acc, bufA, LDS_PAD, IN_PAD, W + pw::PRE0_W, D);
  if (NPRE == 1) {
    store_masked(bufB, true);
  } else {
    store(bufB, false, false);
    __syncthreads();
    ln_tile_inplace<D>(bufB, LDS_PAD, ROWS, W + pw::PRE0_G, W + pw::PRE0_BB, true);
    __syncthreads();
    init_bias(W + pw::PRE1_B);
    gemm_tile_acc2<RPT>(acc, bufB, LDS_PAD, D, W + pw::PRE1_W, D);
    store(bufA, false, false);
    __syncthreads();
    ln_tile_inplace<D>(bufA, LDS_PAD, ROWS, W + pw::PRE1_G, W + pw::PRE1_BB, true);
    __syncthreads();
    init_bias(W + pw::PRE2_B);
    gemm_tile_acc2<RPT>(acc, bufA, LDS_PAD, D, W + pw::PRE2_W, D);
    store_masked(bufB, true);
  }
  __syncthreads();

  // ---- max-pool per polyline (zeros of masked points included) and the pooled half of mlps.0
  for (int i = threadIdx.x; i < G * D; i += 256) {
    int g = i >> 7, c = i & 127;
    float m = bufB[(g * P) * LDS_PAD + c];
    for (int p = 1; p < P; ++p) m = fmaxf(m, bufB[(g * P + p) * LDS_PAD + c]);
    sPool[g * LDS_PAD + c] = m;
  }
  __syncthreads();
  {
    float a4[4];
    acc_init(a4, 0.f);
    gemm_tile_acc<4>(a4, sPool, LDS_PAD, D, W + pw::MLP0_WB, D);
    acc_store_smem<4>(a4, sPP, LDS_PAD, false);
  }
  __syncthreads();

  // ---- mlps.0 (point half + pooled half), LN, ReLU ; mlps.1, ReLU
  {
    const float2 b = __ldg(reinterpret_cast<const float2*>(W + pw::MLP0_B) + cp);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int g = (rg2 * RPT + r) / P;
      const float2 pp = g < G ? *reinterpret_cast<const float2*>(sPP + g * LDS_PAD + 2 * cp) : make_float2(0.f, 0.f);
      acc[r] = make_float2(b.x + pp.x, b.y + pp.y);
    }
  }
  gemm_tile_acc2<RPT>(acc, bufB, LDS_PAD, D, W + pw::MLP0_WA, D);
  store(bufA, false, false);
  __syncthreads();
  ln_tile_inplace<D>(bufA, LDS_PAD, ROWS, W + pw::MLP0_G, W + pw::MLP0_BB, true);
  __syncthreads();
  init_bias(W + pw::MLP1_B);
  gemm_tile_acc2<RPT>(acc, bufA, LDS_PAD, D, W + pw::MLP1_W, D);
  store_masked(bufB, true);
  __syncthreads();
  for (int i = threadIdx.x; i < G * D; i += 256) {
    int g = i >> 7, c = i & 127;
    float m = bufB[(g * P) * LDS_PAD + c];
    for (int p = 1; p < P; ++p) m = fmaxf(m, bufB[(g * P + p) * LDS_PAD + c]);
    sPool[g * LDS_PAD + c] = m;
  }
  __syncthreads();

  // ---- out_mlps on the G pooled rows
  {
    float a4[4];
    acc_init(a4, __ldg(W + pw::OUT0_B + n));
    gemm_tile_acc<4>(a4, sPool, LDS_PAD, D, W + pw::OUT0_W, D);
    acc_store_smem<4>(a4, sPP, LDS_PAD, true);
    __syncthreads();
    acc_init(a4, __ldg(W + pw::OUT1_B + n));
    gemm_tile_acc<4>(a4, sPP, LDS_PAD, D, W + pw::OUT1_W, D);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int g = rg * 4 + r;
      if (g < G && poly0 + g < n_poly) {
        int any = 0;
        for (int p = 0; p < P; ++p) any |= sValid[g * P + p];
        Out[(size_t)(poly0 + g) * D + n] = any ? a4[r] : 0.f;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Version 2: same math and the same summation order as pointnet_kernel (bit-identical results), but every GEMM of the
// chain goes through the shared-memory weight stream of gemm_tile.cuh (gemm2 / WPipe): a weight element is fetched once
// per CTA by cp.async, 2 chunks ahead of the math and across GEMM boundaries, instead of by every thread with LDG.
// (v1 ran 3 CTAs of 64 rows per SM at 23 % of the fp32 FMA rate, waiting on its own weight loads.)
// One CTA = 8 warps, M = 8*TR point rows = G polylines of P points (TR = 16: 11 agent histories / 6 map polylines).
// The per-polyline GEMMs (pooled half of mlps.0, out_mlps) run on 8*TRS >= G pooled rows with the same routine.
// ---------------------------------------------------------------------------------------------------------------
template <int P, int TR>
struct PointNet2Cfg {
  static constexpr int NW = 8;
  static constexpr int ROWS = NW * TR;
  static constexpr int G = ROWS / P;
  static constexpr int TRS = (G + NW - 1) / NW;         // rows per thread of the pooled-row GEMMs
  static constexpr int PROWS = NW * TRS;                // pooled rows held (>= G)
  static constexpr size_t smem_floats = WPIPE_BYTES / sizeof(float) + 2 * (size_t)ROWS * LDS_PAD + 2 * (size_t)PROWS * LDS_PAD + ROWS;
  static constexpr size_t smem_bytes = smem_floats * sizeof(float);
};

template <int IN, int IN_PAD, int NPRE, int P, int TR>
__global__ void __launch_bounds__(256, 1) pointnet2_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask,
                                                           int mask_inner, const int* __restrict__ rows, int n_poly,
                                                           const float* __restrict__ W, float* __restrict__ Out) {
  using C = PointNet2Cfg<P, TR>;
  constexpr int NW = C::NW, ROWS = C::ROWS, G = C::G, TRS = C::TRS, PROWS = C::PROWS;
  extern __shared__ __align__(16) float smem[];
  WPipe pipe = wpipe_init<NW>(smem, [&](WSeg* sg) {
    int n = 0;
    sg[n++] = WSeg{W + pw::PRE0_W, D, IN_PAD};
    if (NPRE == 3) {
      sg[n++] = WSeg{W + pw::PRE1_W, D, D};
      sg[n++] = WSeg{W + pw::PRE2_W, D, D};
    }
    sg[n++] = WSeg{W + pw::MLP0_WB, D, D};
    sg[n++] = WSeg{W + pw::MLP0_WA, D, D};
    sg[n++] = WSeg{W + pw::MLP1_W, D, D};
    sg[n++] = WSeg{W + pw::OUT0_W, D, D};
    sg[n++] = WSeg{W + pw::OUT1_W, D, D};
    return n;
  });
  float* bufA = smem + WPIPE_BYTES / sizeof(float);   // [ROWS][132]
  float* bufB = bufA + ROWS * LDS_PAD;                // [ROWS][132]
  float* sPool = bufB + ROWS * LDS_PAD;               // [PROWS][132]
  float* sPP = sPool + PROWS * LDS_PAD;               // [PROWS][132]
  int* sValid = reinterpret_cast<int*>(sPP + PROWS * LDS_PAD);   // [ROWS]
  const int poly0 = blockIdx.x * G;

  for (int r = threadIdx.x; r < ROWS; r += 256) {
    const int g = r / P, p = r % P;
    int v = 0;
    if (g < G && poly0 + g < n_poly) {
      v = 1;
      if (mask != nullptr) {
        const uint8_t* m = mask + ((size_t)rows[poly0 + g] * P + p) * mask_inner;
        for (int i = 0; i < mask_inner; ++i) v &= (m[i] != 0);
      } else {
        const float* xp = X + ((size_t)rows[poly0 + g] * P + p) * IN;
        for (int i = 0; i < IN; ++i) v &= !isnan(xp[i]);
      }
    }
    sValid[r] = v;
  }
  for (int i = threadIdx.x; i < PROWS * LDS_PAD; i += 256) sPool[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < ROWS * IN_PAD; i += 256) {
    const int r = i / IN_PAD, c = i % IN_PAD;
    float v = 0.f;
    if (sValid[r] && c < IN) v = X[((size_t)rows[poly0 + r / P] * P + (r % P)) * IN + c];
    bufA[r * LDS_PAD + c] = v;
  }
  __syncthreads();

  const TileCoord tc = tile_coord<TR>();
  float acc[TR][4];
  auto store = [&](float* dst, bool relu, bool masked) {
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int row = tc.row + r;
      float4 v = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
      if (masked && !sValid[row]) v = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(dst + row * LDS_PAD + tc.col) = v;
    }
  };
  auto pool = [&](const float* src) {   // max over the P points of each polyline (zeros of masked points included)
    for (int i = threadIdx.x; i < G * D; i += 256) {
      const int g = i >> 7, c = i & 127;
      float m = src[(g * P) * LDS_PAD + c];
      for (int p = 1; p < P; ++p) m = fmaxf(m, src[(g * P + p) * LDS_PAD + c]);
      sPool[g * LDS_PAD + c] = m;
    }
  };

  // ---- pre_mlps
  acc2_init_bias<TR>(acc, W + pw::PRE0_B);
  gemm2<TR, NW>(acc, bufA, LDS_PAD, pipe);
  if (NPRE == 1) {
    store(bufB, true, true);
  } else {
    store(bufB, false, false);
    __syncthreads();
    ln_tile_inplace<D>(bufB, LDS_PAD, ROWS, W + pw::PRE0_G, W + pw::PRE0_BB, true);
    __syncthreads();
    acc2_init_bias<TR>(acc, W + pw::PRE1_B);
    gemm2<TR, NW>(acc, bufB, LDS_PAD, pipe);
    __syncthreads();                                   // every warp is done reading bufA's predecessor tile
    store(bufA, false, false);
    __syncthreads();
    ln_tile_inplace<D>(bufA, LDS_PAD, ROWS, W + pw::PRE1_G, W + pw::PRE1_BB, true);
    __syncthreads();
    acc2_init_bias<TR>(acc, W + pw::PRE2_B);
    gemm2<TR, NW>(acc, bufA, LDS_PAD, pipe);
    __syncthreads();                                   // all reads of bufB (PRE1's input) are long over; keeps the pattern uniform
    store(bufB, true, true);
  }
  __syncthreads();

  // ---- max-pool per polyline and the pooled half of mlps.0 (per polyline, not per point)
  pool(bufB);
  __syncthreads();
  {
    const TileCoord ts = tile_coord<TRS>();
    float a4[TRS][4];
    acc2_init<TRS>(a4, 0.f);
    gemm2<TRS, NW>(a4, sPool, LDS_PAD, pipe);
#pragma unroll
    for (int r = 0; r < TRS; ++r)
      *reinterpret_cast<float4*>(sPP + (ts.row + r) * LDS_PAD + ts.col) = make_float4(a4[r][0], a4[r][1], a4[r][2], a4[r][3]);
  }
  __syncthreads();

  // ---- mlps.0 (point half + pooled half), LN, ReLU ; mlps.1, ReLU
  {
    const float4 b = __ldg(reinterpret_cast<const float4*>(W + pw::MLP0_B + tc.col));
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int g = (tc.row + r) / P;
      float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < G) pp = *reinterpret_cast<const float4*>(sPP + g * LDS_PAD + tc.col);
      acc[r][0] = b.x + pp.x; acc[r][1] = b.y + pp.y; acc[r][2] = b.z + pp.z; acc[r][3] = b.w + pp.w;
    }
  }
  gemm2<TR, NW>(acc, bufB, LDS_PAD, pipe);
  store(bufA, false, false);
  __syncthreads();
  ln_tile_inplace<D>(bufA, LDS_PAD, ROWS, W + pw::MLP0_G, W + pw::MLP0_BB, true);
  __syncthreads();
  acc2_init_bias<TR>(acc, W + pw::MLP1_B);
  gemm2<TR, NW>(acc, bufA, LDS_PAD, pipe);
  __syncthreads();                                     // the pooled-half / point-half GEMMs have finished reading bufB
  store(bufB, true, true);
  __syncthreads();
  pool(bufB);
  __syncthreads();

  // ---- out_mlps on the G pooled rows
  {
    const TileCoord ts = tile_coord<TRS>();
    float a4[TRS][4];
    acc2_init_bias<TRS>(a4, W + pw::OUT0_B);
    gemm2<TRS, NW>(a4, sPool, LDS_PAD, pipe);
#pragma unroll
    for (int r = 0; r < TRS; ++r)
      *reinterpret_cast<float4*>(sPP + (ts.row + r) * LDS_PAD + ts.col) =
          make_float4(fmaxf(a4[r][0], 0.f), fmaxf(a4[r][1], 0.f), fmaxf(a4[r][2], 0.f), fmaxf(a4[r][3], 0.f));
    __syncthreads();
    acc2_init_bias<TRS>(a4, W + pw::OUT1_B);
    gemm2<TRS, NW>(a4, sPP, LDS_PAD, pipe);
#pragma unroll
    for (int r = 0; r < TRS; ++r) {
      const int g = ts.row + r;
      if (g < G && poly0 + g < n_poly) {
        int any = 0;
        for (int p = 0; p < P; ++p) any |= sValid[g * P + p];
        *reinterpret_cast<float4*>(Out + (size_t)(poly0 + g) * D + ts.col) =
            any ? make_float4(a4[r][0], a4[r][1], a4[r][2], a4[r][3]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

}  // namespace prosim
