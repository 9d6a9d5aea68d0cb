// Edge phase of the relative-PE attention layer: three lean kernels per layer.
//
// For destination row i and head h (weights_layout.h, aw::):
//   s_e   = q_i,h . K'_j,h + Qhat_i,h . z_e                 (both pre-scaled by 1/sqrt(16))
//   a_e   = exp(s_e - max) / (sum_e exp(s_e - max) + 1e-16)  (torch_geometric.utils.softmax)
//   Rbar_i,h = sum_e a_e z_e        AggV_i,h = sum_e a_e V'_j,h
// z rows are ZD floats wide: 128 in general, 96 for pure relative-PE edges whose last two 32-feature groups
// are identical (the reference feeds [d, dtheta, phi, phi] to the embedding) -- Qhat's two halves are summed
// when the row is staged and the Wvr' contraction uses the matching folded weight (aw::WVRG96T).
//
//   edge_qk_kernel     Sk[e][h] = q . K'_j      gather of K' rows from L2: a warp per row, one coalesced 512 B
//                      row load per edge, 8 loads in flight per lane, ~40 registers -> full occupancy
//   attn_edge3_kernel  streams z ONCE from HBM: 32-edge tiles staged in smem by cp.async, scores with
//                      lane = edge (all 8 heads in registers, Qhat broadcast from smem), flash-style running
//                      max / sum, aggregation of Rbar with lane = feature column; writes the final
//                      attention weights a_e[8] back for the third kernel
//   edge_av_kernel     AggV = sum_e a_e V'_j     gather of V' rows, same shape as edge_qk_kernel
// The first version fused all three; ncu showed it latency bound on the K'/V' gathers at 12 warps/SM
// (profiles/r1_edge2_ncu_summary.txt).  Partials of the warps of a row are merged in a fixed order
// (deterministic, batch invariant).
#pragma once
#include "common.cuh"
#include "gemm_tile.cuh"   // cp_async16

namespace prosim {

constexpr int EDGE_NW = 6;   // warps per CTA: 6 tile buffers (12.8 KB each at ZD = 96) keep two CTAs resident per SM

template <int ZD>
struct Edge2Cfg {
  static constexpr int ZP = ZD + 4;                 // padded smem row: (ZP/4) odd -> conflict-free LDS.128 per lane-row
  static constexpr int NC = ZD / 32;                // feature columns per lane in the aggregation pass
  static constexpr int PART = 16 + H * ZD;          // per-warp partial: m[8], l[8], Rbar[8][ZD]
  static constexpr int WARP_Z = 32 * ZP;            // floats of one warp's z tile
  static_assert(PART <= WARP_Z, "partial must fit in the warp's tile buffer");
  static constexpr int MAXT = 8;                    // tiles per warp whose running max is remembered for the final rescale
  static constexpr size_t smem_bytes(int rows_per_cta) {
    return sizeof(float) * (size_t)(EDGE_NW * WARP_Z + EDGE_NW * 32 * 8 + rows_per_cta * H * ZD + EDGE_NW * 16 +
                                    EDGE_NW * MAXT * 8);
  }
};

template <int ZD, int WPR>
__global__ void __launch_bounds__(EDGE_NW * 32, 2) attn_edge3_kernel(const float* __restrict__ Qhat,
                                                                     const float* __restrict__ Sk,
                                                                     const float* __restrict__ Z,
                                                                     const int* __restrict__ deg, int stride, int n_dst,
                                                                     float* __restrict__ Rbar, float* __restrict__ Pw) {
  using C = Edge2Cfg<ZD>;
  constexpr int ZP = C::ZP, NC = C::NC, RPC = EDGE_NW / WPR, NT = EDGE_NW * 32;
  static_assert(EDGE_NW % WPR == 0, "warps per row must divide the CTA");
  extern __shared__ __align__(16) float smem[];
  float* sZ = smem;                               // [NW warps][32][ZP]   (reused for the partials at the end)
  float* sP = sZ + EDGE_NW * C::WARP_Z;           // [NW warps][32 edges][8 heads]
  float* sQh = sP + EDGE_NW * 32 * 8;             // [RPC][ZD][8 heads]: head-interleaved so FFMA2 gets natural head pairs
  float* sScale = sQh + RPC * H * ZD;             // [NW warps][16]: per-head merge scale, 1/(L+eps) folded in
  float* sMt = sScale + EDGE_NW * 16;             // [NW warps][MAXT][8]: running max after each of the warp's tiles

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lrow = warp / WPR;                    // row slot of this warp inside the CTA
  const int wir = warp % WPR;                     // warp index inside the row
  const int row = blockIdx.x * RPC + lrow;
  const bool row_ok = row < n_dst;
  const int n_e = row_ok ? min(deg[row], stride) : 0;
  const size_t ebase = (size_t)(row_ok ? row : 0) * stride;

  // the warp's first z tile starts flying before anything else: its HBM latency then overlaps the Qhat staging
  float* zt = sZ + warp * C::WARP_Z;
  auto stage_tile = [&](int t0) {
    const int nt = min(32, n_e - t0);
    const float* src = Z + (ebase + t0) * ZD;      // rows t0..t0+nt-1 of Z are one contiguous stream of nt*ZD floats
    const int chunks = nt * (ZD / 4);
    for (int ch = lane; ch < chunks; ch += 32) {
      const int r = ch / (ZD / 4), c4 = ch % (ZD / 4);
      cp_async16(zt + r * ZP + c4 * 4, src + (size_t)ch * 4);
    }
    cp_async_commit();
  };
  if (wir * 32 < n_e) stage_tile(wir * 32);

  // stage (folded) Qhat of the CTA's rows: all global loads of a thread are issued before the first use (ncu showed
  // 20 % of the kernel's stall samples on the dependent load->add->store chain of the rolled loop)
  {
    constexpr int ITER = (RPC * H * ZD + NT - 1) / NT;
    float va[ITER], vb[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int i = threadIdx.x + it * NT;
      const int r = i / (H * ZD), h = (i / ZD) % H, d = i % ZD;
      const int grow = blockIdx.x * RPC + r;
      va[it] = 0.f;
      vb[it] = 0.f;
      if (i < RPC * H * ZD && grow < n_dst) {
        const float* qh = Qhat + (size_t)grow * H * D + h * D;
        va[it] = __ldg(qh + d);
        if (ZD == 96 && d >= 64) vb[it] = __ldg(qh + d + 32);
      }
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int i = threadIdx.x + it * NT;
      const int r = i / (H * ZD), h = (i / ZD) % H, d = i % ZD;
      if (i < RPC * H * ZD) sQh[r * H * ZD + d * H + h] = va[it] + vb[it];
    }
  }
  __syncthreads();

  float* pt = sP + warp * 32 * 8;
  const float* qh = sQh + lrow * H * ZD;
  float* mt_w = sMt + warp * C::MAXT * 8;

  float m[H], lsum[H];           // running max (warp uniform) and this lane's share of the running sum
  float racc[H][NC];             // Rbar[h][c*32 + lane]
#pragma unroll
  for (int h = 0; h < H; ++h) {
    m[h] = -INFINITY;
    lsum[h] = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) racc[h][c] = 0.f;
  }
  int tile = 0;
  for (int t0 = wir * 32; t0 < n_e; t0 += WPR * 32, ++tile) {
    const int nt = min(32, n_e - t0);
    if (tile > 0) stage_tile(t0);          // (tile 0 was issued at kernel entry)
    // ---- scores, lane = edge: the q.K' part was precomputed by edge_qk_kernel (32 B per edge, coalesced)
    const bool valid = lane < nt;
    float s[H];
    {
      float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
      if (valid) {
        const float4* sp = reinterpret_cast<const float4*>(Sk + (ebase + t0 + lane) * 8);
        s0 = __ldg(sp);
        s1 = __ldg(sp + 1);
      }
      s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w;
      s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
    }
    cp_async_wait<0>();
    __syncwarp();
    if (valid) {
      const float4* zr = reinterpret_cast<const float4*>(zt + lane * ZP);
#pragma unroll 4
      for (int d4 = 0; d4 < ZD / 4; ++d4) {
        const float4 z4 = zr[d4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float zv = i == 0 ? z4.x : i == 1 ? z4.y : i == 2 ? z4.z : z4.w;
          const float4 qa = reinterpret_cast<const float4*>(qh + (d4 * 4 + i) * H)[0];   // heads 0..3 of column d
          const float4 qb = reinterpret_cast<const float4*>(qh + (d4 * 4 + i) * H)[1];   // heads 4..7
          const float2 zz = make_float2(zv, zv);
          const float2 s01 = __ffma2_rn(zz, make_float2(qa.x, qa.y), make_float2(s[0], s[1]));
          const float2 s23 = __ffma2_rn(zz, make_float2(qa.z, qa.w), make_float2(s[2], s[3]));
          const float2 s45 = __ffma2_rn(zz, make_float2(qb.x, qb.y), make_float2(s[4], s[5]));
          const float2 s67 = __ffma2_rn(zz, make_float2(qb.z, qb.w), make_float2(s[6], s[7]));
          s[0] = s01.x; s[1] = s01.y; s[2] = s23.x; s[3] = s23.y;
          s[4] = s45.x; s[5] = s45.y; s[6] = s67.x; s[7] = s67.y;
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
    }
    if (tile < C::MAXT) {
#pragma unroll
      for (int h = 0; h < H; ++h)
        if (lane == h) mt_w[tile * 8 + h] = m[h];
    }
    *reinterpret_cast<float4*>(pt + lane * 8) = make_float4(p[0], p[1], p[2], p[3]);
    *reinterpret_cast<float4*>(pt + lane * 8 + 4) = make_float4(p[4], p[5], p[6], p[7]);
    __syncwarp();
    // ---- aggregation, lane = feature column, edges of the tile in ascending order
    for (int e = 0; e < nt; ++e) {
      const float4 pa = *reinterpret_cast<const float4*>(pt + e * 8);
      const float4 pb = *reinterpret_cast<const float4*>(pt + e * 8 + 4);
      const float* zr = zt + e * ZP + lane;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float zv = zr[c * 32];
        const float2 zz = make_float2(zv, zv);
        const float2 r01 = __ffma2_rn(zz, make_float2(pa.x, pa.y), make_float2(racc[0][c], racc[1][c]));
        const float2 r23 = __ffma2_rn(zz, make_float2(pa.z, pa.w), make_float2(racc[2][c], racc[3][c]));
        const float2 r45 = __ffma2_rn(zz, make_float2(pb.x, pb.y), make_float2(racc[4][c], racc[5][c]));
        const float2 r67 = __ffma2_rn(zz, make_float2(pb.z, pb.w), make_float2(racc[6][c], racc[7][c]));
        racc[0][c] = r01.x; racc[1][c] = r01.y; racc[2][c] = r23.x; racc[3][c] = r23.y;
        racc[4][c] = r45.x; racc[5][c] = r45.y; racc[6][c] = r67.x; racc[7][c] = r67.y;
      }
    }
    // unnormalised weights of this tile (relative to the running max stored in mt_w), rescaled at the end
    if (valid) {
      float4* pw = reinterpret_cast<float4*>(Pw + (ebase + t0 + lane) * 8);
      pw[0] = make_float4(p[0], p[1], p[2], p[3]);
      pw[1] = make_float4(p[4], p[5], p[6], p[7]);
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
    for (int o = wir * 32 + lane; o < H * ZD; o += WPR * 32) {
      const int h = o / ZD;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < WPR; ++w) t = fmaf(scl[w * 16 + h], base[w * C::WARP_Z + o], t);
      Rbar[(size_t)row * H * ZD + o] = t;
    }
    // final attention weights: a_e = p_e * exp(m_tile - M) / (L + 1e-16) = p_e * exp(m_tile - m_warp) * scale_warp
    const float* myscale = sScale + warp * 16;
    tile = 0;
    for (int t0 = wir * 32; t0 < n_e; t0 += WPR * 32, ++tile) {
      if (t0 + lane < n_e) {
        float4* pw = reinterpret_cast<float4*>(Pw + (ebase + t0 + lane) * 8);
        float4 a = pw[0], b = pw[1];
        const float* mtt = mt_w + (tile < C::MAXT ? tile : C::MAXT - 1) * 8;
        float f[H];
#pragma unroll
        for (int h = 0; h < H; ++h) f[h] = expf(mtt[h] - m[h]) * myscale[h];
        a.x *= f[0]; a.y *= f[1]; a.z *= f[2]; a.w *= f[3];
        b.x *= f[4]; b.y *= f[5]; b.z *= f[6]; b.w *= f[7];
        pw[0] = a;
        pw[1] = b;
      }
    }
  }
}

// ---- gather kernels: a warp per destination row, lane = 4 of the 128 columns, EB edges in flight per lane
constexpr int GATHER_EB = 8;

// Sk[row*stride + e][h] = sum_c q[row][h*16+c] * K'[nbr[e]][h*16+c]
__global__ void __launch_bounds__(256) edge_qk_kernel(const float* __restrict__ Qg, const float* __restrict__ KV,
                                                      const int* __restrict__ nbr, const int* __restrict__ deg, int stride,
                                                      int n_dst, float* __restrict__ Sk, int* __restrict__ row_counter) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *row_counter = 0;   // dynamic row queue of the attn_edge4 launch that follows
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_dst) return;
  const int n_e = min(deg[row], stride);
  const size_t ebase = (size_t)row * stride;
  const float4 q4 = __ldg(reinterpret_cast<const float4*>(Qg + (size_t)row * D) + lane);
  for (int e0 = 0; e0 < n_e; e0 += 32) {
    const int jl = e0 + lane < n_e ? __ldg(nbr + ebase + e0 + lane) : 0;
    const int nt = min(32, n_e - e0);
    for (int g = 0; g < nt; g += GATHER_EB) {
      float4 k4[GATHER_EB];
#pragma unroll
      for (int u = 0; u < GATHER_EB; ++u) {
        const int j = __shfl_sync(0xffffffffu, jl, min(g + u, nt - 1));
        k4[u] = __ldg(reinterpret_cast<const float4*>(KV + (size_t)j * 256) + lane);
      }
#pragma unroll
      for (int u = 0; u < GATHER_EB; ++u) {
        float d = fmaf(q4.w, k4[u].w, fmaf(q4.z, k4[u].z, fmaf(q4.y, k4[u].y, q4.x * k4[u].x)));
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        if ((lane & 3) == 0 && g + u < nt) Sk[(ebase + e0 + g + u) * 8 + (lane >> 2)] = d;
      }
    }
  }
}

// AggV[row][c] = sum_e a[e][c/16] * V'[nbr[e]][c]   (edges in ascending order)
// Ft (nullable): per (row, 32-edge tile, head) factor that turns the unnormalised weights of attn_edge4_kernel into
// attention weights; NULL = Pw already holds them (attn_edge3_kernel).
__global__ void __launch_bounds__(256) edge_av_kernel(const float* __restrict__ Pw, const float* __restrict__ Ft, int ft_tiles,
                                                      const float* __restrict__ KV, const int* __restrict__ nbr,
                                                      const int* __restrict__ deg, int stride, int n_dst,
                                                      float* __restrict__ AggV) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_dst) return;
  const int n_e = min(deg[row], stride);
  const size_t ebase = (size_t)row * stride;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e0 = 0; e0 < n_e; e0 += 32) {
    const int jl = e0 + lane < n_e ? __ldg(nbr + ebase + e0 + lane) : 0;
    const int nt = min(32, n_e - e0);
    const float f = Ft != nullptr ? __ldg(Ft + ((size_t)row * ft_tiles + (e0 >> 5)) * 8 + (lane >> 2)) : 1.0f;
    for (int g = 0; g < nt; g += GATHER_EB) {
      float4 v4[GATHER_EB];
      float a[GATHER_EB];
#pragma unroll
      for (int u = 0; u < GATHER_EB; ++u) {
        const int eu = min(g + u, nt - 1);
        const int j = __shfl_sync(0xffffffffu, jl, eu);
        v4[u] = __ldg(reinterpret_cast<const float4*>(KV + (size_t)j * 256 + 128) + lane);
        a[u] = g + u < nt ? __ldg(Pw + (ebase + e0 + eu) * 8 + (lane >> 2)) * f : 0.f;
      }
#pragma unroll
      for (int u = 0; u < GATHER_EB; ++u) {
        acc.x = fmaf(a[u], v4[u].x, acc.x);
        acc.y = fmaf(a[u], v4[u].y, acc.y);
        acc.z = fmaf(a[u], v4[u].z, acc.z);
        acc.w = fmaf(a[u], v4[u].w, acc.w);
      }
    }
  }
  *reinterpret_cast<float4*>(AggV + (size_t)row * D + 4 * lane) = acc;
}

}  // namespace prosim
