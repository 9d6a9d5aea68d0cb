// Edge phase, z-streaming kernel v4: one warp per destination row, persistent over rows, z tiles fetched by TMA.
//
// Same math as edge2.cuh's header (scores s_e = q.K' + Qhat.z_e, segment softmax with the +1e-16 of
// torch_geometric.utils.softmax, Rbar = sum_e a_e z_e) but restructured after the ncu capture of attn_edge3
// (profiles/r1_layer_v3_ncu_summary.txt, r1_edge3_sass_mix.txt): that kernel ran 12 warps/SM at 32 % issue
// utilisation -- a quarter of its stall samples sat on the per-CTA start-up chain (deg -> branch, Qhat staging,
// __syncthreads), 24 LDGSTS + address arithmetic per lane per tile, and a cross-warp merge per row.  Here
//   * a warp owns a whole row (tiles of 32 edges in ascending order, flash-style running max/sum): no CTA
//     barrier, no cross-warp merge, results independent of the launch shape (batch invariant by construction);
//   * warps loop over rows (grid = one CTA per SM), so nothing is paid per CTA;
//   * a z tile is 3 (ZD=96) or 4 TMA boxes of [8 edges x 32 floats] per 8 edges, written 128B-swizzled into
//     shared memory: lane = edge reads its row with conflict-free LDS.128, lane = feature column reads a
//     column conflict-free, and the load costs one instruction per 1 KB instead of 64 LDGSTS;
//   * the row's raw Qhat [8][128] arrives with one bulk copy on the same mbarrier as the row's first tile;
//   * the attention weights leave the kernel unnormalised (relative to the running max of their tile) together
//     with one factor per (row, tile, head); edge_av_kernel applies it -- no second pass over the weights.
// Per edge: 384 B of z (HBM, read once) + 32 B of q.K' in + 32 B of weights out; per row 4 KB Qhat in, 3 KB Rbar out.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "edge2.cuh"

namespace prosim {
namespace e4 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// L2 policy for data that is read exactly once (the z stream): do not let it push the layer's reusable buffers
// (K'|V', Rbar, Qhat, the packed weights) out of the 126 MB L2
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
// one [8 rows x 32 floats] box of the z tensor (128B swizzle) -> 1 KB of shared memory
__device__ __forceinline__ void tma_box(uint32_t dst, const CUtensorMap* tm, int col, int row, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
      "[%4], %5;\n" ::"r"(dst),
      "l"(tm), "r"(col), "r"(row), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

}  // namespace e4

template <int ZD>
struct Edge4Cfg {
  static constexpr int NSEG = ZD / 32;              // 32-float (128 B) column segments of a z row
  static constexpr int ZBYTES = NSEG * 4096;        // [NSEG][32 edges][128 B]
  static constexpr int QBYTES = H * D * 4;          // raw Qhat row [8][128]
  static constexpr int PBYTES = 32 * H * 4;         // tile weights [32 edges][8 heads]
  static constexpr int MT_TILES = 24;               // running max per tile: stride <= 768
  static constexpr int MBYTES = MT_TILES * H * 4;
  static constexpr int WARP_BYTES = ZBYTES + QBYTES + PBYTES + MBYTES;
  static constexpr size_t smem_bytes(int nw) { return 1024 + (size_t)nw * WARP_BYTES + (size_t)nw * 8; }
};

template <int ZD, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
    attn_edge4_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmZ32,
                      const float* __restrict__ Qhat, const float* __restrict__ Sk,
                      const int* __restrict__ deg, int stride, int n_dst, float* __restrict__ Rbar, float* __restrict__ Pw,
                      float* __restrict__ Ft, int ft_tiles, int* __restrict__ row_counter) {
  using C = Edge4Cfg<ZD>;
  constexpr int NSEG = C::NSEG;
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* gbase = smem_raw + ((1024u - (e4::smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzled boxes need 1 KB alignment
  uint8_t* zb = gbase + warp * C::ZBYTES;
  float* qb = reinterpret_cast<float*>(gbase + NW * C::ZBYTES + warp * C::QBYTES);
  float* pb = reinterpret_cast<float*>(gbase + NW * (C::ZBYTES + C::QBYTES) + warp * C::PBYTES);
  float* mt = reinterpret_cast<float*>(gbase + NW * (C::ZBYTES + C::QBYTES + C::PBYTES) + warp * C::MBYTES);
  const uint32_t bar = e4::smem_u32(gbase + NW * C::WARP_BYTES + warp * 8);
  const uint32_t zb_s = e4::smem_u32(zb), qb_s = e4::smem_u32(qb);
  if (lane == 0) e4::mbar_init(bar, 1);
  __syncwarp();

  uint32_t phase = 0;
  const uint64_t z_policy = e4::policy_evict_first();
  // swizzled byte offsets inside a tile: zoff[j] = 16-byte chunk j of this lane's row (score pass, lane = edge);
  // coff[u] = this lane's feature column in edge u of a group of 8 (aggregation pass, lane = column)
  uint32_t zoff[8], coff[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    zoff[j] = lane * 128 + ((j ^ (lane & 7)) << 4);
    coff[j] = j * 128 + (((lane >> 2) ^ j) << 4) + (lane & 3) * 4;
  }
  // rows are handed out dynamically (row_counter is zeroed by edge_qk_kernel, which always runs first): a warp gets
  // only 2-5 rows of 1-16 tiles each, so a static split leaves a quarter of the warps idle at the tail
  int row = blockIdx.x * NW + warp;
  int row_next = 0;
  for (; row < n_dst; row = __shfl_sync(0xffffffffu, row_next, 0)) {
    if (lane == 0) row_next = gridDim.x * NW + atomicAdd(row_counter, 1);   // consumed at the end of this row
    const int n_e = min(__ldg(deg + row), stride);
    const size_t ebase = (size_t)row * stride;
    float* rb = Rbar + (size_t)row * H * ZD;
    if (n_e <= 0) {   // no in-edges: the aggregate is zero (the gate / FFN update still runs on the row)
#pragma unroll
      for (int i = 0; i < H * NSEG; ++i) rb[i * 32 + lane] = 0.f;
      continue;
    }
    const int ntiles = (n_e + 31) >> 5;
    float m[H], lsum[H];           // running max (warp uniform), this lane's share of the running sum
    float2 r01[NSEG], r23[NSEG], r45[NSEG], r67[NSEG];   // Rbar[h][seg*32 + lane] as head pairs
#pragma unroll
    for (int h = 0; h < H; ++h) {
      m[h] = -INFINITY;
      lsum[h] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < NSEG; ++c) r01[c] = r23[c] = r45[c] = r67[c] = make_float2(0.f, 0.f);

    for (int t = 0; t < ntiles; ++t) {
      const int t0 = t << 5, nt = min(32, n_e - t0);
      // ---- fetch: z boxes of this tile (+ the row's Qhat with its first tile), all on one mbarrier phase
      // a full tile is NSEG boxes of [32 edges x 32 floats] (tmZ32); a partial one is fetched in 8-edge boxes so that at
      // most 7 rows beyond the list are read
      const bool full_tile = nt == 32;
      const int nbox = full_tile ? NSEG : ((nt + 7) >> 3) * NSEG;
      __syncwarp();                                   // every lane is done with the previous tile's buffers
      if (lane == 0) e4::mbar_expect_tx(bar, (full_tile ? NSEG * 4096 : nbox * 1024) + (t == 0 ? C::QBYTES : 0));
      __syncwarp();
      if (lane < nbox) {
        e4::fence_proxy_async();
        if (full_tile) {
          e4::tma_box(zb_s + lane * 4096, &tmZ32, lane * 32, (int)(ebase + t0), bar, z_policy);
        } else {
          const int g = lane / NSEG, sg = lane % NSEG;
          e4::tma_box(zb_s + sg * 4096 + g * 1024, &tmZ, sg * 32, (int)(ebase + t0 + g * 8), bar, z_policy);
        }
      } else if (t == 0 && lane == 31) {
        e4::fence_proxy_async();
        e4::bulk_copy(qb_s, Qhat + (size_t)row * H * D, C::QBYTES, bar);
      }
      // ---- the q.K' part of the scores (edge_qk_kernel), lane = edge: its latency hides behind the tile's
      const bool valid = lane < nt;
      float2 acc[H];
      {
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        if (valid) {
          const float4* sp = reinterpret_cast<const float4*>(Sk + (ebase + t0 + lane) * 8);
          s0 = __ldg(sp);
          s1 = __ldg(sp + 1);
        }
        acc[0] = make_float2(s0.x, 0.f); acc[1] = make_float2(s0.y, 0.f);
        acc[2] = make_float2(s0.z, 0.f); acc[3] = make_float2(s0.w, 0.f);
        acc[4] = make_float2(s1.x, 0.f); acc[5] = make_float2(s1.y, 0.f);
        acc[6] = make_float2(s1.z, 0.f); acc[7] = make_float2(s1.w, 0.f);
      }
      e4::mbar_wait(bar, phase);
      phase ^= 1;
      if (ZD == 96 && t == 0) {   // features 96..127 of the embedding duplicate 64..95: fold Qhat once per row
#pragma unroll
        for (int h = 0; h < H; ++h) qb[h * D + 64 + lane] += qb[h * D + 96 + lane];
        __syncwarp();
      }
      // ---- scores: s_h += sum_d z[d] Qhat[h][d]; even / odd d accumulate in the two halves of an FFMA2
      if (valid) {
#pragma unroll
        for (int sg = 0; sg < NSEG; ++sg) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 z4 = *reinterpret_cast<const float4*>(zb + sg * 4096 + zoff[j]);
            const float2 zlo = make_float2(z4.x, z4.y), zhi = make_float2(z4.z, z4.w);
#pragma unroll
            for (int h = 0; h < H; ++h) {
              const float4 q4 = *reinterpret_cast<const float4*>(qb + h * D + sg * 32 + j * 4);
              acc[h] = __ffma2_rn(zlo, make_float2(q4.x, q4.y), acc[h]);
              acc[h] = __ffma2_rn(zhi, make_float2(q4.z, q4.w), acc[h]);
            }
          }
        }
      }
      // ---- online softmax bookkeeping (per head; max is warp uniform)
      float p[H], corr[H];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float s = acc[h].x + acc[h].y;
        const float mx = warp_max(valid ? s : -INFINITY);
        const float mn = fmaxf(m[h], mx);                       // nt >= 1 => finite
        // ex2.approx(x log2 e): arguments are <= 0, 2 ulp on the weights -- the accurate expf costs ~7 instructions, 16
        // of them per tile were a quarter of the softmax bookkeeping
        corr[h] = __expf(m[h] - mn);                            // exp(-inf) = 0 on the first tile
        p[h] = valid ? __expf(s - mn) : 0.f;
        lsum[h] = lsum[h] * corr[h] + p[h];
        m[h] = mn;
        if (lane == h) mt[t * H + h] = mn;
      }
#pragma unroll
      for (int c = 0; c < NSEG; ++c) {                          // head pairs: one packed multiply per accumulator
        r01[c] = __fmul2_rn(r01[c], make_float2(corr[0], corr[1]));
        r23[c] = __fmul2_rn(r23[c], make_float2(corr[2], corr[3]));
        r45[c] = __fmul2_rn(r45[c], make_float2(corr[4], corr[5]));
        r67[c] = __fmul2_rn(r67[c], make_float2(corr[6], corr[7]));
      }
      *reinterpret_cast<float4*>(pb + lane * 8) = make_float4(p[0], p[1], p[2], p[3]);
      *reinterpret_cast<float4*>(pb + lane * 8 + 4) = make_float4(p[4], p[5], p[6], p[7]);
      if (valid) {   // unnormalised weights of this tile; edge_av_kernel applies Ft[row][tile][head]
        float4* pw = reinterpret_cast<float4*>(Pw + (ebase + t0 + lane) * 8);
        pw[0] = make_float4(p[0], p[1], p[2], p[3]);
        pw[1] = make_float4(p[4], p[5], p[6], p[7]);
      }
      __syncwarp();
      // ---- aggregation, lane = feature column of each segment, edges of the tile in ascending order
      auto agg_edge = [&](int e, int u) {
        const float4 pa = *reinterpret_cast<const float4*>(pb + e * 8);
        const float4 pq = *reinterpret_cast<const float4*>(pb + e * 8 + 4);
        const uint8_t* ze = zb + (e - u) * 128 + coff[u];
#pragma unroll
        for (int c = 0; c < NSEG; ++c) {
          const float zv = *reinterpret_cast<const float*>(ze + c * 4096);
          const float2 zz = make_float2(zv, zv);
          r01[c] = __ffma2_rn(zz, make_float2(pa.x, pa.y), r01[c]);
          r23[c] = __ffma2_rn(zz, make_float2(pa.z, pa.w), r23[c]);
          r45[c] = __ffma2_rn(zz, make_float2(pq.x, pq.y), r45[c]);
          r67[c] = __ffma2_rn(zz, make_float2(pq.z, pq.w), r67[c]);
        }
      };
      // whole groups of 8 edges: the weights of the lanes beyond the list are exactly 0 and the rows they multiply were
      // fetched with the tile's last 8-edge box (finite z of other rows, or TMA zero fill), so they add +-0 -- and the
      // loop has no remainder blocks (seven predicated copies of the body cost 24 register moves per group)
      // (groups run from the last to the first through a fall-through switch: straight-line code with one entry per
      // group count -- as a counted loop the compiler paid 24 register moves per group at the back edge, 9.5 % of the
      // kernel's instructions)
#define PROSIM_AGG_GROUP(E0)                       \
  _Pragma("unroll") for (int u = 0; u < 8; ++u) agg_edge((E0) + u, u);
      switch ((nt + 7) >> 3) {
        case 4: PROSIM_AGG_GROUP(24)
        case 3: PROSIM_AGG_GROUP(16)
        case 2: PROSIM_AGG_GROUP(8)
        default: PROSIM_AGG_GROUP(0)
      }
#undef PROSIM_AGG_GROUP
    }

    // ---- row epilogue: 1 / (sum + 1e-16), Rbar, and the per-tile factors exp(m_tile - m_final) / (sum + 1e-16)
    float inv[H];
#pragma unroll
    for (int h = 0; h < H; ++h) inv[h] = 1.0f / (warp_sum(lsum[h]) + 1e-16f);
#pragma unroll
    for (int c = 0; c < NSEG; ++c) {
      rb[0 * ZD + c * 32 + lane] = r01[c].x * inv[0];
      rb[1 * ZD + c * 32 + lane] = r01[c].y * inv[1];
      rb[2 * ZD + c * 32 + lane] = r23[c].x * inv[2];
      rb[3 * ZD + c * 32 + lane] = r23[c].y * inv[3];
      rb[4 * ZD + c * 32 + lane] = r45[c].x * inv[4];
      rb[5 * ZD + c * 32 + lane] = r45[c].y * inv[5];
      rb[6 * ZD + c * 32 + lane] = r67[c].x * inv[6];
      rb[7 * ZD + c * 32 + lane] = r67[c].y * inv[7];
    }
    if (lane == 0) {
      *reinterpret_cast<float4*>(pb) = make_float4(inv[0], inv[1], inv[2], inv[3]);
      *reinterpret_cast<float4*>(pb + 4) = make_float4(inv[4], inv[5], inv[6], inv[7]);
    }
    __syncwarp();   // mt[] (lanes 0..7) and the normalisers are read by all lanes
    {
      const float mh = mt[(ntiles - 1) * H + (lane & 7)];   // running max after the last tile = the row's max
      const float ih = pb[lane & 7];
      float* ft = Ft + (size_t)row * ft_tiles * H;
      for (int i = lane; i < ntiles * H; i += 32) ft[i] = expf(mt[i] - mh) * ih;
    }
  }
}

}  // namespace prosim
