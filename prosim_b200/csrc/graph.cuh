// Neighbour search and per-edge relative positional encoding.
//
// Neighbour lists replace torch_cluster.{radius,radius_graph,knn_graph} (un-vendored wheel; call sites
// prosim/models/scene_encoder/attn_fusion.py:107,109, decoder/sym_coord.py:86,94,
// policy/act_decoder.py:250,259).  Semantics are those restated in oracle/graph.py and must match it
// bit for bit: squared distance dx*dx + dy*dy with every operation rounded separately (no FMA),
// strict `< r^2`, first `cap` hits in ascending source index, kNN ties to the lower index.
// Layout: fixed-stride rows  nbr[q*stride + j], j < deg[q]  (ascending source index inside a row), so no
// scan / host round trip is needed and every later kernel can be launched with a grid over rows.
#pragma once
#include "common.cuh"

namespace prosim {

__device__ __forceinline__ float sqdist_rn(float2 a, float2 b) {
  float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// seg[b] = {start0, len0, start1, len1}: the (up to two) contiguous source ranges of scene b.
__global__ void __launch_bounds__(256) radius_kernel(const float2* __restrict__ qpos, const int* __restrict__ qscene,
                                                     int Nq, const float2* __restrict__ spos,
                                                     const int4* __restrict__ seg, float r2, int cap, int drop_self,
                                                     int* __restrict__ nbr, int* __restrict__ deg, int stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + warp;
  if (q >= Nq) return;
  const float2 p = qpos[q];
  const int4 sg = seg[qscene[q]];
  const int limit = cap + (drop_self ? 1 : 0);
  const unsigned lt = (1u << lane) - 1u;
  int count = 0, self_seen = 0;
  int* out = nbr + (size_t)q * stride;
  for (int s = 0; s < 2 && count < limit; ++s) {
    const int start = s == 0 ? sg.x : sg.z, len = s == 0 ? sg.y : sg.w;
    for (int base = 0; base < len && count < limit; base += 32) {
      const int c = base + lane;
      const int idx = start + c;
      bool in = false;
      if (c < len) in = sqdist_rn(p, spos[idx]) < r2;
      const unsigned bal = __ballot_sync(0xffffffffu, in);
      const int pos = count + __popc(bal & lt);
      const bool take = in && pos < limit;
      const bool is_self = drop_self && idx == q;
      const unsigned sbal = __ballot_sync(0xffffffffu, take && is_self);
      const int opos = pos - self_seen - __popc(sbal & lt);
      if (take && !is_self && opos < stride) out[opos] = idx;
      count += __popc(bal);
      self_seen |= (sbal != 0u);
    }
  }
  if (lane == 0) deg[q] = min(min(count, limit) - self_seen, stride);
}

// k nearest sources (self included when it is among the sources): a warp per query, 8 queries per CTA.
// Selection rule (oracle/graph.py): rank by (distance, index); take rank < k.  Instead of ranking every candidate
// against every other (n^2 compares per query -- 0.8 ms per launch for the 640-token scenes) the warp finds the k-th
// smallest key with a 32-step radix select over the float bit patterns (distances are >= 0, so the unsigned integer
// order is the float order), then takes every candidate below the threshold plus the lowest-index ties: same set, same
// ascending-index output order, n * 34 compares per query.
// dynamic smem: 8 warps x nmax keys
__global__ void __launch_bounds__(256) knn_kernel(const float2* __restrict__ qpos, const int* __restrict__ qscene, int Nq,
                                                  const float2* __restrict__ spos, const int4* __restrict__ seg, int k,
                                                  int nmax, int* __restrict__ nbr, int* __restrict__ deg, int stride) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + warp;
  if (q >= Nq) return;
  unsigned* key = reinterpret_cast<unsigned*>(smem) + (size_t)warp * nmax;
  const float2 p = qpos[q];
  const int4 sg = seg[qscene[q]];
  const int n = min(sg.y + sg.w, nmax);
  for (int c = lane; c < n; c += 32) {
    const int idx = c < sg.y ? sg.x + c : sg.z + (c - sg.y);
    key[c] = __float_as_uint(sqdist_rn(p, spos[idx]));
  }
  __syncwarp();
  const int keff = min(min(k, n), stride);
  int* out = nbr + (size_t)q * stride;
  if (keff <= 0) {
    if (lane == 0) deg[q] = 0;
    return;
  }
  // radix select: T = keff-th smallest key; `remaining` ends as the number of keys equal to T that belong to the set
  unsigned prefix = 0u, decided = 0u;
  int remaining = keff;
  for (int bit = 31; bit >= 0; --bit) {
    const unsigned b = 1u << bit;
    int cnt = 0;
    for (int c = lane; c < n; c += 32) {
      const unsigned v = key[c];
      cnt += ((v & decided) == prefix) && ((v & b) == 0u);
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (remaining > cnt) {
      prefix |= b;
      remaining -= cnt;
    }
    decided |= b;
  }
  const unsigned lt = (1u << lane) - 1u;
  int count = 0, ties = 0;
  for (int base = 0; base < n; base += 32) {
    const int c = base + lane;
    const unsigned v = c < n ? key[c] : 0xffffffffu;
    const bool less = c < n && v < prefix;
    const bool tie = c < n && v == prefix;
    const unsigned tb = __ballot_sync(0xffffffffu, tie);
    const bool take = less || (tie && ties + __popc(tb & lt) < remaining);   // ties go to the lower index
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    if (take) out[count + __popc(bal & lt)] = c < sg.y ? sg.x + c : sg.z + (c - sg.y);
    count += __popc(bal);
    ties += __popc(tb);
  }
  if (lane == 0) deg[q] = count;
}

// Per-edge relative PE, LayerNorm-normalised WITHOUT affine (layer independent; the affine is folded into
// the packed weights).  Reference: policy/act_decoder.py:203-221 + layers/fourier_embedding.py:56-79.
//   input = [|dp|, wrap(theta_src - theta_dst), phi, phi],  phi = atan2(c x dp, c . dp), c = (cos, sin)(theta_dst)
//   feature[i*32 + 2m + {0,1}] = {sin, cos}( input_i * 2pi / dim_t[m] ),  dim_t[m] = 10000^(m/16)
// One CTA (4 warps) per destination row; 8 lanes per edge (4 edges per warp and pass), lane sub-index s owns the
// ZD/8 consecutive features [s*ZD/8, (s+1)*ZD/8) = ZD/16 (sin, cos) pairs: the per-edge geometry (atan2, sqrt, wrap)
// and the LayerNorm reductions are paid once per 4 edges instead of once per edge (the first version ran a warp per
// edge with 4 features per lane and a quarter of the lanes idle at ZD = 96).
// extra (optional): per-edge [128] vector added to the PE before the normalisation (condition edges,
// condition_transformer/condition_attns.py:211-216), indexed like Z.
// ZD = 128 stores all features; ZD = 96 (only without `extra`) drops features 96..127, which are bit-identical
// copies of 64..95 (the embedding gets phi twice); the statistics still run over all 128 (those features count twice).
template <int ZD>
__global__ void __launch_bounds__(128) edge_pe_kernel(const float2* __restrict__ dpos, const float* __restrict__ dori,
                                                      const float2* __restrict__ spos, const float* __restrict__ sori,
                                                      const int* __restrict__ nbr, const int* __restrict__ deg, int stride,
                                                      const float* __restrict__ dim_t, const float* __restrict__ extra,
                                                      float* __restrict__ Z) {
  constexpr int NF = ZD / 8;          // features per lane: 12 or 16
  constexpr int NP = NF / 2;          // (sin, cos) pairs per lane
  __shared__ float s_dt[16];
  if (threadIdx.x < 16) s_dt[threadIdx.x] = dim_t[threadIdx.x];
  __syncthreads();
  const int row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 3, s = lane & 7;
  const int n_e = min(deg[row], stride);
  const float2 pd = dpos[row];
  const float od = dori[row];
  const float cx = cosf(od), cy = sinf(od);
  const float TWO_PI_F = 6.28318530717958647692f;
  // Zero padding: the edge kernel (edge4.cuh) aggregates whole groups of 8 edges with weight 0 on the entries beyond the
  // list, so up to 7 entries past a list's end must hold finite values -- its own row's [n_e, round8(n_e)) and, when a
  // list ends at the stride, the first 7 entries of the NEXT row (hence at least 8 initialised entries per row).
  {
    const int n8 = (n_e + 7) & ~7;
    const int pad_end = min(stride, n8 > 8 ? n8 : 8);
    for (int i = n_e * (ZD / 4) + threadIdx.x; i < pad_end * (ZD / 4); i += 128)
      reinterpret_cast<float4*>(Z + (size_t)row * stride * ZD)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int e0 = 0; e0 < n_e; e0 += 16) {
    const int e = e0 + warp * 4 + sub;
    const bool ok = e < n_e;                               // uniform over the 8 lanes of an edge
    const size_t ei = (size_t)row * stride + (ok ? e : 0);
    const int j = nbr[ei];
    const float2 ps = spos[j];
    const float rx = ps.x - pd.x, ry = ps.y - pd.y;
    // torch.norm on the CPU reference reduces with acc = fma(v, v, acc): sqrt(fma(ry, ry, rx*rx)) bit for bit;
    // the dot product is a torch sum over two rounded products starting from +0 (so -0 + -0 becomes +0 and the
    // zero-offset self edge gets phi = atan2(+-0, +0) = 0, not pi).
    const float in0 = sqrtf(fmaf(ry, ry, __fmul_rn(rx, rx))) * TWO_PI_F;
    const float in1 = wrap_angle(sori[j] - od) * TWO_PI_F;
    const float in2 = atan2f(__fsub_rn(__fmul_rn(cx, ry), __fmul_rn(cy, rx)),
                             __fadd_rn(__fadd_rn(0.0f, __fmul_rn(cx, rx)), __fmul_rn(cy, ry))) * TWO_PI_F;
    float f[NF];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int p = s * NP + k;                            // pair index: scalar p / 16, frequency p % 16
      const int i = p >> 4;
      const float v = i == 0 ? in0 : i == 1 ? in1 : in2;
      sincosf(v / s_dt[p & 15], &f[2 * k], &f[2 * k + 1]);   // accurate path, one shared range reduction per argument
    }
    if (extra != nullptr) {
#pragma unroll
      for (int k = 0; k < NF; k += 4) {
        const float4 x = *reinterpret_cast<const float4*>(extra + ei * D + s * NF + k);
        f[k] += x.x; f[k + 1] += x.y; f[k + 2] += x.z; f[k + 3] += x.w;
      }
    }
    // LayerNorm over the 128 features (no affine); with ZD = 96 features 64..95 stand for 96..127 as well
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NF; ++k) {
      const float w = (ZD == 96 && s * NF + k >= 64) ? 2.0f : 1.0f;
      sum = fmaf(w, f[k], sum);
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    const float mean = sum * (1.0f / 128.0f);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NF; ++k) {
      const float w = (ZD == 96 && s * NF + k >= 64) ? 2.0f : 1.0f;
      const float d = f[k] - mean;
      q = fmaf(w * d, d, q);
    }
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    q += __shfl_xor_sync(0xffffffffu, q, 4);
    const float rstd = 1.0f / sqrtf(q * (1.0f / 128.0f) + LN_EPS);
    if (ok) {
      float4* out = reinterpret_cast<float4*>(Z + ei * ZD + s * NF);
#pragma unroll
      for (int k = 0; k < NF; k += 4)
        out[k >> 2] = make_float4((f[k] - mean) * rstd, (f[k + 1] - mean) * rstd, (f[k + 2] - mean) * rstd,
                                  (f[k + 3] - mean) * rstd);
    }
  }
}

}  // namespace prosim
