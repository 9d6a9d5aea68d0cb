// Edge phase of the relative-PE attention layer, version 2: ONE pass over the per-edge data.
//
// For destination row i and head h (weights_layout.h, aw::):
//   s_e   = q_i,h . K'_j,h + Qhat_i,h . z_e                 (both pre-scaled by 1/sqrt(16))
//   a_e   = exp(s_e - max) / (sum_e exp(s_e - max) + 1e-16)  (torch_geometric.utils.softmax)
//   Rbar_i,h = sum_e a_e z_e        AggV_i,h = sum_e a_e V'_j,h
// z rows are ZD floats wide: 128 in general, 96 for pure relative-PE edges whose last two 32-feature groups
// are identical (the reference feeds [d, dtheta, phi, phi] to the embedding) -- Qhat's two halves are summed
// when the row is staged and the Wvr' contraction uses the matching folded weight (aw::WVRG96T).
//
// Mapping: a CTA of EDGE_NW = 6 warps (two CTAs per SM) serves 6/WPR destination rows; each row's edges are cut into 32-edge tiles dealt
// round-robin to its WPR warps.  A warp stages its tile's z rows in shared memory with cp.async (one coalesced
// 32*ZD*4-byte stream), scores it with lane = edge (all 8 heads in registers, Qhat broadcast from smem), keeps
// flash-style running max / sum / accumulators, then aggregates the same staged tile with lane = feature
// column.  z is read from HBM exactly once per layer; K'|V' rows come from L2.  Partials of the WPR warps are
// merged in a fixed order (deterministic, batch invariant).
#pragma once
#include "common.cuh"
#include "gemm_tile.cuh"   // cp_async16

namespace prosim {

constexpr int EDGE_NW = 6;   // warps per CTA: 6 tile buffers (12.8 KB each at ZD = 96) keep two CTAs resident per SM

template <int ZD>
struct Edge2Cfg {
  static constexpr int ZP = ZD + 4;                 // padded smem row: (ZP/4) odd -> conflict-free LDS.128 per lane-row
  static constexpr int NC = ZD / 32;                // feature columns per lane in the aggregation pass
  static constexpr int PART = 16 + H * ZD + D;      // per-warp partial: m[8], l[8], Rbar[8][ZD], AggV[128]
  static constexpr int WARP_Z = 32 * ZP;            // floats of one warp's z tile
  static_assert(PART <= WARP_Z, "partial must fit in the warp's tile buffer");
  static constexpr size_t smem_bytes(int rows_per_cta) {
    return sizeof(float) * (size_t)(EDGE_NW * WARP_Z + EDGE_NW * 32 * 8 + rows_per_cta * (H * ZD + D) + EDGE_NW * 16);
  }
};

template <int ZD, int WPR>
__global__ void __launch_bounds__(EDGE_NW * 32, 2) attn_edge2_kernel(const float* __restrict__ Qg, const float* __restrict__ Qhat,
                                                            const float* __restrict__ KV, const float* __restrict__ Z,
                                                            const int* __restrict__ nbr, const int* __restrict__ deg,
                                                            int stride, int n_dst, float* __restrict__ Rbar,
                                                            float* __restrict__ AggV) {
  using C = Edge2Cfg<ZD>;
  constexpr int ZP = C::ZP, NC = C::NC, RPC = EDGE_NW / WPR, NT = EDGE_NW * 32;
  static_assert(EDGE_NW % WPR == 0, "warps per row must divide the CTA");
  extern __shared__ __align__(16) float smem[];
  float* sZ = smem;                               // [NW warps][32][ZP]   (reused for the partials at the end)
  float* sP = sZ + EDGE_NW * C::WARP_Z;           // [NW warps][32 edges][8 heads]
  float* sQh = sP + EDGE_NW * 32 * 8;             // [RPC][8][ZD]
  float* sQ = sQh + RPC * H * ZD;                 // [RPC][128]
  float* sScale = sQ + RPC * D;                   // [NW warps][16]: per-head merge scale, 1/(L+eps) folded in

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lrow = warp / WPR;                    // row slot of this warp inside the CTA
  const int wir = warp % WPR;                     // warp index inside the row
  const int row = blockIdx.x * RPC + lrow;
  const bool row_ok = row < n_dst;
  const int n_e = row_ok ? min(deg[row], stride) : 0;
  const size_t ebase = (size_t)(row_ok ? row : 0) * stride;

  // stage q and (folded) Qhat of the CTA's rows
  for (int i = threadIdx.x; i < RPC * H * ZD; i += NT) {
    const int r = i / (H * ZD), h = (i / ZD) % H, d = i % ZD;
    const int grow = blockIdx.x * RPC + r;
    float v = 0.f;
    if (grow < n_dst) {
      const float* qh = Qhat + (size_t)grow * H * D + h * D;
      v = qh[d];
      if (ZD == 96 && d >= 64) v += qh[d + 32];
    }
    sQh[i] = v;
  }
  for (int i = threadIdx.x; i < RPC * D; i += NT) {
    const int grow = blockIdx.x * RPC + i / D;
    sQ[i] = grow < n_dst ? Qg[(size_t)grow * D + (i % D)] : 0.f;
  }
  __syncthreads();

  float* zt = sZ + warp * C::WARP_Z;
  float* pt = sP + warp * 32 * 8;
  const float* qh = sQh + lrow * H * ZD;
  const float* q = sQ + lrow * D;

  float m[H], lsum[H];           // running max (warp uniform) and this lane's share of the running sum
  float racc[H][NC];             // Rbar[h][c*32 + lane]
  float vacc[4];                 // AggV[c*32 + lane], head = 2c + (lane >> 4)
#pragma unroll
  for (int h = 0; h < H; ++h) {
    m[h] = -INFINITY;
    lsum[h] = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) racc[h][c] = 0.f;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) vacc[c] = 0.f;

  for (int t0 = wir * 32; t0 < n_e; t0 += WPR * 32) {
    const int nt = min(32, n_e - t0);
    // ---- stage the tile: rows t0..t0+nt-1 of Z are one contiguous stream of nt*ZD floats
    {
      const float* src = Z + (ebase + t0) * ZD;
      const int chunks = nt * (ZD / 4);
      for (int ch = lane; ch < chunks; ch += 32) {
        const int r = ch / (ZD / 4), c4 = ch % (ZD / 4);
        cp_async16(zt + r * ZP + c4 * 4, src + (size_t)ch * 4);
      }
      cp_async_commit();
    }
    // ---- scores, lane = edge: K' part straight from L2 while the tile lands
    const bool valid = lane < nt;
    const int j = valid ? __ldg(nbr + ebase + t0 + lane) : 0;
    float s[H];
#pragma unroll
    for (int h = 0; h < H; ++h) s[h] = 0.f;
    if (valid) {
      const float4* kp = reinterpret_cast<const float4*>(KV + (size_t)j * 256);
#pragma unroll
      for (int h = 0; h < H; ++h) {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 k4 = __ldg(kp + h * 4 + c4);
          const float4 q4 = reinterpret_cast<const float4*>(q)[h * 4 + c4];
          s[h] = fmaf(q4.x, k4.x, s[h]);
          s[h] = fmaf(q4.y, k4.y, s[h]);
          s[h] = fmaf(q4.z, k4.z, s[h]);
          s[h] = fmaf(q4.w, k4.w, s[h]);
        }
      }
    }
    cp_async_wait<0>();
    __syncwarp();
    if (valid) {
      const float4* zr = reinterpret_cast<const float4*>(zt + lane * ZP);
#pragma unroll 4
      for (int d4 = 0; d4 < ZD / 4; ++d4) {
        const float4 z4 = zr[d4];
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float4 q4 = reinterpret_cast<const float4*>(qh + h * ZD)[d4];
          s[h] = fmaf(q4.x, z4.x, s[h]);
          s[h] = fmaf(q4.y, z4.y, s[h]);
          s[h] = fmaf(q4.z, z4.z, s[h]);
          s[h] = fmaf(q4.w, z4.w, s[h]);
        }
      }
    }
    // ---- online softmax bookkeeping (per head; max is warp uniform)
    float p[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float mt = warp_max(valid ? s[h] : -INFINITY);
      const float mn = fmaxf(m[h], mt);                       // nt >= 1 => finite
      const float corr = expf(m[h] - mn);                     // exp(-inf) = 0 on the first tile
      p[h] = valid ? expf(s[h] - mn) : 0.f;
      lsum[h] = lsum[h] * corr + p[h];
      m[h] = mn;
#pragma unroll
      for (int c = 0; c < NC; ++c) racc[h][c] *= corr;
      s[h] = corr;                                            // keep for the AggV rescale below
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) vacc[c] *= (lane >> 4) ? s[2 * c + 1] : s[2 * c];
    *reinterpret_cast<float4*>(pt + lane * 8) = make_float4(p[0], p[1], p[2], p[3]);
    *reinterpret_cast<float4*>(pt + lane * 8 + 4) = make_float4(p[4], p[5], p[6], p[7]);
    __syncwarp();
    // ---- aggregation, lane = feature column, edges of the tile in ascending order
    // V' rows come from L2 (~600 cycles): fetch them EB edges ahead of their use so the loads of a whole group
    // are in flight together; edges past the tile end are clamped and get weight 0.
    constexpr int EB = 4;
    for (int e0 = 0; e0 < nt; e0 += EB) {
      float vv[EB][4];
#pragma unroll
      for (int u = 0; u < EB; ++u) {
        const int je = __shfl_sync(0xffffffffu, j, min(e0 + u, nt - 1));
        const float* vp = KV + (size_t)je * 256 + 128 + lane;
#pragma unroll
        for (int c = 0; c < 4; ++c) vv[u][c] = __ldg(vp + c * 32);
      }
#pragma unroll
      for (int u = 0; u < EB; ++u) {
        const int e = e0 + u;
        if (e < nt) {
          const float4 pa = *reinterpret_cast<const float4*>(pt + e * 8);
          const float4 pb = *reinterpret_cast<const float4*>(pt + e * 8 + 4);
          const float pe[H] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
          const float* zr = zt + e * ZP + lane;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float zv = zr[c * 32];
#pragma unroll
            for (int h = 0; h < H; ++h) racc[h][c] = fmaf(pe[h], zv, racc[h][c]);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float pv = (lane >> 4) ? pe[2 * c + 1] : pe[2 * c];
            vacc[c] = fmaf(pv, vv[u][c], vacc[c]);
          }
        }
      }
    }
    __syncwarp();   // the tile buffer and pt are rewritten by the next iteration
  }

  // ---- publish this warp's partial in its own tile buffer
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const float lt = warp_sum(lsum[h]);
    if (lane == 0) {
      zt[h] = m[h];
      zt[8 + h] = lt;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) zt[16 + h * ZD + c * 32 + lane] = racc[h][c];
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) zt[16 + H * ZD + c * 32 + lane] = vacc[c];
  __syncthreads();
  // merge scales: scale[w][h] = exp(m_w - M) / (sum_w exp(m_w - M) l_w + 1e-16); first warp of each row computes them
  if (wir == 0 && lane < H) {
    const float* base = sZ + (lrow * WPR) * C::WARP_Z;
    float M = -INFINITY;
    for (int w = 0; w < WPR; ++w) M = fmaxf(M, base[w * C::WARP_Z + lane]);
    float L = 0.f, sc[WPR];
    for (int w = 0; w < WPR; ++w) {
      const float mw = base[w * C::WARP_Z + lane];
      sc[w] = mw == -INFINITY ? 0.f : expf(mw - M);
      L += sc[w] * base[w * C::WARP_Z + 8 + lane];
    }
    const float inv = 1.0f / (L + 1e-16f);
    for (int w = 0; w < WPR; ++w) sScale[(lrow * WPR + w) * 16 + lane] = sc[w] * inv;
  }
  __syncthreads();
  if (row_ok) {
    const float* base = sZ + (lrow * WPR) * C::WARP_Z + 16;
    const float* scl = sScale + (lrow * WPR) * 16;
    for (int o = wir * 32 + lane; o < H * ZD + D; o += WPR * 32) {
      const int h = o < H * ZD ? o / ZD : (o - H * ZD) >> 4;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < WPR; ++w) t = fmaf(scl[w * 16 + h], base[w * C::WARP_Z + o], t);
      if (o < H * ZD) Rbar[(size_t)row * H * ZD + o] = t;
      else AggV[(size_t)row * D + (o - H * ZD)] = t;
    }
  }
}

}  // namespace prosim
