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

// Tried and dropped (round 1): the same chain with every GEMM on gemm_tile.cuh's shared-memory weight stream (gemm2 /
// WPipe, 128-row CTAs, 8 warps x 16 rows or 16 warps x 8 rows; bit-identical results).  ncu: 174 us against 187 us for the
// agent histories but 1.67 ms against 1.40 ms for the map polylines (one CTA per SM at 218 KB, issue slots 34 % busy
// against 48 %), 3.06 ms against 2.91 ms per forward in total.  This kernel already runs at 45 % (map) / 27 % (agent
// histories) of the fp32 FMA peak; the next step for it is the tensor pipe, not another FFMA tiling.

}  // namespace prosim
