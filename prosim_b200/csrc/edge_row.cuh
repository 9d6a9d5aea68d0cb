// Edge phase for SMALL launches (a single scene: 128 destination rows): one CTA of four warps per destination row, one launch
// for what edge_qk_kernel -> attn_edge4_kernel -> edge_av_kernel do in three.
//
// At this size nothing is throughput bound: a launch is as long as its longest row, and in the persistent kernel a row is one
// warp walking its tiles alone (1400 instructions per 32-edge tile; 157 us for the generator's 512-edge rows).  Here the four
// warps of a CTA split the row BY HEAD PAIR -- warp w owns heads 2w, 2w+1 in the score pass, the softmax and the Rbar
// aggregation, and columns 32w .. 32w+31 (the same two heads) of the V' aggregation -- and by tile in the q.K' gathers.  Every
// output element is still produced by exactly the instruction sequence of the three-kernel path (heads never interact; the
// gathers and the per-column FMA chains run in the same edge order), so the results are BIT-IDENTICAL to it and a scene's
// rollout does not depend on which path its launches took (tests/test_gpu_rollout.py::test_batch_invariance_*).
// The z tiles of a row go through a 4-stage ring filled by warp 0 before the gathers start: a 128-edge agent row has all of
// its z in flight from the first instruction.
#pragma once
#include "edge4.cuh"

namespace prosim {

template <int ZD>
struct EdgeRowCfg {
  static constexpr int NSTAGE = 4;
  static constexpr int NSEG = ZD / 32;
  static constexpr int ZBYTES = NSEG * 4096;        // [NSEG][32 edges][128 B], 128B-swizzled TMA boxes (edge4.cuh)
  static constexpr int QBYTES = H * D * 4;          // raw Qhat row [8][128]
  static constexpr int PBYTES = 32 * H * 4;         // tile weights [32 edges][8 heads]
  static constexpr int MT_TILES = Edge4Cfg<ZD>::MT_TILES;
  static constexpr int MBYTES = MT_TILES * H * 4;
  static constexpr size_t smem_bytes() { return 1024 + NSTAGE * ZBYTES + QBYTES + PBYTES + MBYTES + NSTAGE * 8; }
};

constexpr int EDGE_ROW_EB = 32;   // gathers in flight per lane

template <int ZD>
__global__ void __launch_bounds__(128)
    attn_edge_row_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmZ32,
                         const float* Qhat_, const float* Qg_, const float* KV_, const int* nbr_, const int* deg_, int stride,
                         int n_dst, float* Sk_, float* Rbar_, float* Pw_, float* Ft_, int ft_tiles, float* AggV_) {
  using C = EdgeRowCfg<ZD>;
  constexpr int NSEG = C::NSEG, NSTAGE = C::NSTAGE;
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x;
  const int h0 = 2 * warp;                               // this warp's head pair
  uint8_t* gbase = smem_raw + ((1024u - (e4::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* zb0 = gbase;
  float* qb = reinterpret_cast<float*>(gbase + NSTAGE * C::ZBYTES);
  float* pb = reinterpret_cast<float*>(gbase + NSTAGE * C::ZBYTES + C::QBYTES);
  float* mt = reinterpret_cast<float*>(gbase + NSTAGE * C::ZBYTES + C::QBYTES + C::PBYTES);
  const uint32_t bar0 = e4::smem_u32(gbase + NSTAGE * C::ZBYTES + C::QBYTES + C::PBYTES + C::MBYTES);
  const uint32_t zb0_s = e4::smem_u32(zb0), qb_s = e4::smem_u32(qb);
  pdl_launch_dependents();
  if (threadIdx.x < NSTAGE) e4::mbar_init(bar0 + threadIdx.x * 8, 1);
  __syncthreads();
  pdl_wait();   // every input of this kernel is produced by its predecessors (common.cuh: programmatic dependent launch)
  // Qg (like Sk / Pw / Ft below) is read with ld.global.cg only: under programmatic dependent launch this CTA may be resident
  // while the ping-pong q buffer is still being rewritten, which the non-coherent path's read-only contract excludes (post_sw.cuh)
  const float* __restrict__ Qhat = pdl_acquire(Qhat_);
  const float* Qg = pdl_acquire(Qg_);
  const float* __restrict__ KV = pdl_acquire(KV_);
  const int* __restrict__ nbr = pdl_acquire(nbr_);
  const int* __restrict__ deg = pdl_acquire(deg_);
  float* Sk = pdl_acquire(Sk_);
  float* Pw = pdl_acquire(Pw_);
  float* Ft = pdl_acquire(Ft_);
  float* __restrict__ Rbar = pdl_acquire(Rbar_);
  float* __restrict__ AggV = pdl_acquire(AggV_);

  const int n_e = min(__ldg(deg + row), stride);
  const int ntiles = (n_e + 31) >> 5;
  const size_t ebase = (size_t)row * stride;
  const uint64_t z_policy = e4::policy_evict_first();
  // warp 0: tile t -> buffer t % NSTAGE (+ the row's Qhat with tile 0); same box shapes as attn_edge4_kernel
  auto fetch_tile = [&](int t) {
    const int t0 = t << 5, nt = min(32, n_e - t0);
    const bool full_tile = nt == 32;
    const int nbox = full_tile ? NSEG : ((nt + 7) >> 3) * NSEG;
    const uint32_t bar = bar0 + (t % NSTAGE) * 8, zb_s = zb0_s + (t % NSTAGE) * C::ZBYTES;
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
  };
  if (warp == 0) {
#pragma unroll
    for (int t = 0; t < NSTAGE; ++t)
      if (t < ntiles) fetch_tile(t);
  }
  // ---- q.K' scores of the row: tile t by warp t % 4 (edge_qk_kernel's arithmetic, edge by edge)
  edge_qk_row<EDGE_ROW_EB, true>(Qg, KV, nbr, deg, stride, row, lane, Sk, warp, 4);
  __syncthreads();                                        // Sk is read back below by other warps (through L2)

  float* rb = Rbar + (size_t)row * H * ZD;
  if (n_e > 0) {
    uint32_t zoff[8], coff[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      zoff[j] = lane * 128 + ((j ^ (lane & 7)) << 4);
      coff[j] = j * 128 + (((lane >> 2) ^ j) << 4) + (lane & 3) * 4;
    }
    float m[2] = {-INFINITY, -INFINITY}, lsum[2] = {0.f, 0.f};
    float2 r[NSEG];                                       // Rbar[h0 | h0 + 1][seg * 32 + lane]
#pragma unroll
    for (int c = 0; c < NSEG; ++c) r[c] = make_float2(0.f, 0.f);

    for (int t = 0; t < ntiles; ++t) {
      const int t0 = t << 5, nt = min(32, n_e - t0);
      const int stage = t % NSTAGE;
      const uint8_t* zb = zb0 + stage * C::ZBYTES;
      const bool valid = lane < nt;
      float2 acc[2];
      {
        float2 s2 = make_float2(0.f, 0.f);
        if (valid) s2 = __ldcg(reinterpret_cast<const float2*>(Sk + (ebase + t0 + lane) * 8 + h0));
        acc[0] = make_float2(s2.x, 0.f);
        acc[1] = make_float2(s2.y, 0.f);
      }
      e4::mbar_wait(bar0 + stage * 8, (t / NSTAGE) & 1);
      if (ZD == 96 && t == 0) {   // features 96..127 of the embedding duplicate 64..95: fold this warp's two Qhat rows
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) qb[(h0 + hh) * D + 64 + lane] += qb[(h0 + hh) * D + 96 + lane];
        __syncwarp();
      }
      if (valid) {
#pragma unroll
        for (int sg = 0; sg < NSEG; ++sg) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 z4 = *reinterpret_cast<const float4*>(zb + sg * 4096 + zoff[j]);
            const float2 zlo = make_float2(z4.x, z4.y), zhi = make_float2(z4.z, z4.w);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const float4 q4 = *reinterpret_cast<const float4*>(qb + (h0 + hh) * D + sg * 32 + j * 4);
              acc[hh] = __ffma2_rn(zlo, make_float2(q4.x, q4.y), acc[hh]);
              acc[hh] = __ffma2_rn(zhi, make_float2(q4.z, q4.w), acc[hh]);
            }
          }
        }
      }
      float p[2], corr[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float s = acc[hh].x + acc[hh].y;
        const float mx = warp_max(valid ? s : -INFINITY);
        const float mn = fmaxf(m[hh], mx);
        corr[hh] = __expf(m[hh] - mn);
        p[hh] = valid ? __expf(s - mn) : 0.f;
        lsum[hh] = lsum[hh] * corr[hh] + p[hh];
        m[hh] = mn;
        if (lane == hh) mt[t * H + h0 + hh] = mn;
      }
#pragma unroll
      for (int c = 0; c < NSEG; ++c) r[c] = __fmul2_rn(r[c], make_float2(corr[0], corr[1]));
      *reinterpret_cast<float2*>(pb + lane * 8 + h0) = make_float2(p[0], p[1]);
      if (valid) *reinterpret_cast<float2*>(Pw + (ebase + t0 + lane) * 8 + h0) = make_float2(p[0], p[1]);
      __syncwarp();
      auto agg_edge = [&](int e, int u) {
        const float2 pa = *reinterpret_cast<const float2*>(pb + e * 8 + h0);
        const uint8_t* ze = zb + (e - u) * 128 + coff[u];
#pragma unroll
        for (int c = 0; c < NSEG; ++c) {
          const float zv = *reinterpret_cast<const float*>(ze + c * 4096);
          r[c] = __ffma2_rn(make_float2(zv, zv), pa, r[c]);
        }
      };
      // same group order as attn_edge4_kernel: 8-edge groups from the last to the first, ascending inside a group
#define PROSIM_AGG_GROUP(E0) \
  _Pragma("unroll") for (int u = 0; u < 8; ++u) agg_edge((E0) + u, u);
      switch ((nt + 7) >> 3) {
        case 4: PROSIM_AGG_GROUP(24)
        case 3: PROSIM_AGG_GROUP(16)
        case 2: PROSIM_AGG_GROUP(8)
        default: PROSIM_AGG_GROUP(0)
      }
#undef PROSIM_AGG_GROUP
      if (t + NSTAGE < ntiles) {                          // uniform over the CTA
        __syncthreads();                                  // all four warps are done with this buffer
        if (warp == 0) fetch_tile(t + NSTAGE);
      }
    }
    // ---- row epilogue for this warp's heads
    float inv[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) inv[hh] = 1.0f / (warp_sum(lsum[hh]) + 1e-16f);
#pragma unroll
    for (int c = 0; c < NSEG; ++c) {
      rb[(h0 + 0) * ZD + c * 32 + lane] = r[c].x * inv[0];
      rb[(h0 + 1) * ZD + c * 32 + lane] = r[c].y * inv[1];
    }
    __syncwarp();                                         // mt[] of this warp's heads was written by lanes 0 and 1
    {
      const int hh = lane & 1;
      const float mh = mt[(ntiles - 1) * H + h0 + hh];
      const float ih = hh ? inv[1] : inv[0];
      float* ft = Ft + (size_t)row * ft_tiles * H;
      for (int tt = lane >> 1; tt < ntiles; tt += 16) ft[tt * H + h0 + hh] = expf(mt[tt * H + h0 + hh] - mh) * ih;
    }
  } else {
#pragma unroll
    for (int c = 0; c < NSEG; ++c) {
      rb[(h0 + 0) * ZD + c * 32 + lane] = 0.f;
      rb[(h0 + 1) * ZD + c * 32 + lane] = 0.f;
    }
  }
  __syncthreads();                                        // Pw / Ft of the row are complete (read back through L2)

  // ---- AggV[row][col] = sum_e a_e V'[nbr_e][col], col = 32 warp + lane (edge_av_kernel's chain for that column)
  {
    const int col = 32 * warp + lane, head = col >> 4;
    float acc = 0.f;
    for (int e0 = 0; e0 < n_e; e0 += 32) {
      const int jl = e0 + lane < n_e ? __ldg(nbr + ebase + e0 + lane) : 0;
      const int nt = min(32, n_e - e0);
      const float f = __ldcg(Ft + ((size_t)row * ft_tiles + (e0 >> 5)) * 8 + head);
      float v[EDGE_ROW_EB], a[EDGE_ROW_EB];
#pragma unroll
      for (int u = 0; u < EDGE_ROW_EB; ++u) {
        const int eu = min(u, nt - 1);
        const int j = __shfl_sync(0xffffffffu, jl, eu);
        v[u] = __ldg(KV + (size_t)j * 256 + 128 + col);
        a[u] = u < nt ? __ldcg(Pw + (ebase + e0 + eu) * 8 + head) * f : 0.f;
      }
#pragma unroll
      for (int u = 0; u < EDGE_ROW_EB; ++u) acc = fmaf(a[u], v[u], acc);
    }
    AggV[(size_t)row * D + col] = acc;
  }
}

}  // namespace prosim
