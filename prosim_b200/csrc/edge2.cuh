// Edge phase of the relative-PE attention layer: three lean kernels per layer.
//
// For destination row i and head h (weights_layout.h, aw::):
//   s_e   = q_i,h . K'_j,h + Qhat_i,h . z_e                 (both pre-scaled by 1/sqrt(16))
//   a_e   = exp(s_e - max) / (sum_e exp(s_e - max) + 1e-16)  (torch_geometric.utils.softmax)
//   Rbar_i,h = sum_e a_e z_e        AggV_i,h = sum_e a_e V'_j,h
//
//   edge_qk_kernel     Sk[e][h] = q . K'_j      gather of K' rows from L2: a warp per row, one coalesced 512 B
//                      row load per edge, 8 loads in flight per lane, 48 registers -> full occupancy      (this file)
//   attn_edge4_kernel  streams z ONCE from HBM (TMA tiles), scores + flash-style softmax + Rbar            (edge4.cuh)
//   edge_av_kernel     AggV = sum_e a_e V'_j     gather of V' rows, same shape as edge_qk_kernel           (this file)
// The first version fused all three; ncu showed it latency bound on the K'/V' gathers at 12 warps/SM
// (profiles/r1_edge2_ncu_summary.txt).  A variant of the gather kernels that staged a scene's K' / V' block in
// shared memory for the dense graphs (policy agent->policy: 64 KB once per 16 rows instead of 93 x 512 B per row) was
// measured slower (29 / 35 us against 26 / 25 us): at 64 resident warps per SM the L2 gathers already run near the L2
// bandwidth, and the 80 KB tile halves the occupancy that hides the remaining latencies.
#pragma once
#include "common.cuh"
#include "gemm_tile.cuh"   // cp_async16

namespace prosim {

// ---- gather kernels: a warp per destination row, lane = 4 of the 128 columns, EB edges in flight per lane
constexpr int GATHER_EB = 8;

// Sk[row*stride + e][h] = sum_c q[row][h*16+c] * K'[nbr[e]][h*16+c]   -- one destination row, one warp
// COHERENT: the queries were written by a kernel this one may overlap with under programmatic dependent launch (edge_row.cuh):
// read them through L2, not through the non-coherent path
template <int EB = GATHER_EB, bool COHERENT = false>
__device__ __forceinline__ void edge_qk_row(const float* __restrict__ Qg, const float* __restrict__ KV,
                                            const int* __restrict__ nbr, const int* __restrict__ deg, int stride, int row,
                                            int lane, float* __restrict__ Sk, int tile_first = 0, int tile_step = 1) {
  const int n_e = min(deg[row], stride);
  const size_t ebase = (size_t)row * stride;
  const float4* qp = reinterpret_cast<const float4*>(Qg + (size_t)row * D) + lane;
  const float4 q4 = COHERENT ? __ldcg(qp) : __ldg(qp);
  for (int e0 = 32 * tile_first; e0 < n_e; e0 += 32 * tile_step) {   // (tile_first, tile_step): the caller's share of the row
    const int jl = e0 + lane < n_e ? __ldg(nbr + ebase + e0 + lane) : 0;
    const int nt = min(32, n_e - e0);
    for (int g = 0; g < nt; g += EB) {
      float4 k4[EB];
#pragma unroll
      for (int u = 0; u < EB; ++u) {
        const int j = __shfl_sync(0xffffffffu, jl, min(g + u, nt - 1));
        k4[u] = __ldg(reinterpret_cast<const float4*>(KV + (size_t)j * 256) + lane);
      }
#pragma unroll
      for (int u = 0; u < EB; ++u) {
        float d = fmaf(q4.w, k4[u].w, fmaf(q4.z, k4[u].z, fmaf(q4.y, k4[u].y, q4.x * k4[u].x)));
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        if ((lane & 3) == 0 && g + u < nt) Sk[(ebase + e0 + g + u) * 8 + (lane >> 2)] = d;
      }
    }
  }
}

__global__ void __launch_bounds__(256) edge_qk_kernel(const float* __restrict__ Qg, const float* __restrict__ KV,
                                                      const int* __restrict__ nbr, const int* __restrict__ deg, int stride,
                                                      int n_dst, float* __restrict__ Sk, int* __restrict__ row_counter) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *row_counter = 0;   // dynamic row queue of the attn_edge4 launch that follows
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_dst) return;
  edge_qk_row(Qg, KV, nbr, deg, stride, row, lane, Sk);
}

// AggV[row][c] = sum_e a[e][c/16] * V'[nbr[e]][c]   (edges in ascending order)
// Ft (nullable): per (row, 32-edge tile, head) factor that turns the unnormalised weights of attn_edge4_kernel into
// attention weights; NULL = Pw already holds them (attn_edge3_kernel).
template <int EB = GATHER_EB>
__device__ __forceinline__ void edge_av_row(const float* __restrict__ Pw, const float* __restrict__ Ft, int ft_tiles,
                                            const float* __restrict__ KV, const int* __restrict__ nbr,
                                            const int* __restrict__ deg, int stride, int row, int lane,
                                            float* __restrict__ AggV) {
  const int n_e = min(deg[row], stride);
  const size_t ebase = (size_t)row * stride;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e0 = 0; e0 < n_e; e0 += 32) {
    const int jl = e0 + lane < n_e ? __ldg(nbr + ebase + e0 + lane) : 0;
    const int nt = min(32, n_e - e0);
    const float* fp = Ft + ((size_t)row * ft_tiles + (e0 >> 5)) * 8 + (lane >> 2);
    const float f = Ft != nullptr ? __ldg(fp) : 1.0f;
    for (int g = 0; g < nt; g += EB) {
      float4 v4[EB];
      float a[EB];
#pragma unroll
      for (int u = 0; u < EB; ++u) {
        const int eu = min(g + u, nt - 1);
        const int j = __shfl_sync(0xffffffffu, jl, eu);
        v4[u] = __ldg(reinterpret_cast<const float4*>(KV + (size_t)j * 256 + 128) + lane);
        const float* pp = Pw + (ebase + e0 + eu) * 8 + (lane >> 2);
        a[u] = g + u < nt ? __ldg(pp) * f : 0.f;
      }
#pragma unroll
      for (int u = 0; u < EB; ++u) {
        acc.x = fmaf(a[u], v4[u].x, acc.x);
        acc.y = fmaf(a[u], v4[u].y, acc.y);
        acc.z = fmaf(a[u], v4[u].z, acc.z);
        acc.w = fmaf(a[u], v4[u].w, acc.w);
      }
    }
  }
  *reinterpret_cast<float4*>(AggV + (size_t)row * D + 4 * lane) = acc;
}

__global__ void __launch_bounds__(256) edge_av_kernel(const float* __restrict__ Pw, const float* __restrict__ Ft, int ft_tiles,
                                                      const float* __restrict__ KV, const int* __restrict__ nbr,
                                                      const int* __restrict__ deg, int stride, int n_dst,
                                                      float* __restrict__ AggV) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_dst) return;
  edge_av_row(Pw, Ft, ft_tiles, KV, nbr, deg, stride, row, lane, AggV);
}

}  // namespace prosim
