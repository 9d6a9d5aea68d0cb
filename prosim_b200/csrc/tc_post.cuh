// Node side of the AttentionLayer on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-equivalent via 3xTF32.
//
// Same math as attn_post2_kernel (attn2.cuh; reference prosim/models/layers/attention_layer.py:102-118, :38-43 and the
// next layer's destination-side projections :56-66), one CTA per 128 destination rows:
//   agg = AggV + Rbar_h . Wvr'_h        8 x [128 x zd] . [zd x 16]
//   g   = sigmoid(agg . Wga + Gx) ;  u = agg + g (S - agg)
//   x1  = x + LN(u . Wo + bo) ;  y = W2 relu(W1 LN(x1) + b1) + b2 ;  out = x1 + LN(y)
//   next layer:  s | gx | q = LN_dst'(out) . {Ws, Wgx, Wq} ;  Qhat_h = q_h . (diag(gamma_r) Wkr_h)
// Every product a*w is evaluated as a_lo*w_hi + a_hi*w_lo + a_hi*w_hi with hi = tf32(v), lo = tf32(v - hi)
// (relative error ~2^-21 per product, fp32 accumulate in tensor memory) -- the per-tick parity gate of 1e-5 needs
// fp32-class arithmetic, a single TF32 pass (2^-11) fails it (tests/test_gpu_kernels.py::test_tc_gemm).
//
// Structure (224 threads, 1 CTA/SM, all 512 TMEM columns):
//   warps 0-3  epilogue: thread = row = TMEM lane.  tcgen05.ld the accumulator, apply bias / gate / LayerNorm /
//              residual in registers (a LayerNorm is thread local: no shuffles), split the result into hi/lo and
//              tcgen05.st it back as the A operand of the next GEMM -- activations never exist as smem operands
//   warp 4     weight producer: one thread streams the layer's pre-split B operands (aw::TC_*, weights.py) from L2
//              through a 3 x 32 KB shared-memory ring with cp.async.bulk on full/empty mbarriers
//   warp 5     MMA issuer: one thread issues tcgen05.mma kind::tf32 with A from TMEM, B from the ring; tcgen05.commit
//              releases ring stages and publishes accumulators
//   warp 6     input producer: one thread prefetches the rows' inputs (Rbar, AggV, Gx, S, x, and the x1 it parked in
//              the output buffer) as TMA boxes of [128 rows x 32 floats], 128B-swizzled, through a 4 x 16 KB ring
//   outputs leave through two 16 KB staging tiles and TMA stores.
// (v1 of this kernel let every thread load / store its own row straight from global memory: 700 warp instructions
//  with 32 distinct lines each per CTA plus ~70 exposed L2 latencies -- 184 us per launch, slower than the FFMA kernel.)
// TMEM columns: [0,128) ACC0, [128,256) ACC1, [256,384) A hi, [384,512) A lo.  During the FFN the hidden layer is
// produced and consumed in 32-column chunks (ACC0 is re-used as 2 x 32 up-accumulators + 32 hi + 32 lo hidden
// columns) so that the 512-wide hidden activation never has to exist anywhere; while the eight Rbar_h operands are
// staged (zd = 96) columns [128,512) hold two alternating [hi 96 | lo 96] buffers.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "edge4.cuh"     // e4:: mbarrier / bulk-copy / TMA helpers
#include "tc_gemm.cuh"   // tc:: tcgen05 helpers
#include "weights_layout.h"

namespace prosim {
namespace tcp {

constexpr int NSTAGE = 3;            // weight ring stages
constexpr int STAGE_BYTES = 32768;
constexpr int NIN_G = 2;             // input tile ring stages PER epilogue group (a slot is only ever consumed by one group,
                                     // so that every waiter sees every phase of its barriers: parity waits alias otherwise)
constexpr int NIN = 2 * NIN_G;
constexpr int TILE_BYTES = 16384;    // [128 rows][32 floats], 128B-swizzled
constexpr int THREADS = 352;           // 8 epilogue warps (two column groups) + weight producer + MMA issuer + input producer
constexpr uint32_t ACC0 = 0, ACC1 = 128, AHI = 256, ALO = 384;
constexpr uint32_t UP0 = 0, HHI = 64, HLO = 96;   // inside ACC0 during the FFN: up accumulators at 0 / 32

// per-column vectors staged once in shared memory (float offsets)
constexpr int V_BO = 0, V_LNPOST_G = 128, V_LNPOST_B = 256, V_LNFFPRE_G = 384, V_LNFFPRE_B = 512, V_B1 = 640, V_B2 = 1152,
              V_LNFFPOST_G = 1280, V_LNFFPOST_B = 1408, V_LNDST_G = 1536, V_LNDST_B = 1664, V_BQ = 1792, V_BS = 1920,
              V_BG = 2048, V_SIZE = 2176;

struct Maps {   // 2-D tensor maps [rows][cols] fp32, box 128 rows x 32 floats, 128B swizzle
  CUtensorMap rbar, aggv, s, gx, x, out, q_n, s_n, gx_n, qhat_n;
};

struct Smem {
  uint8_t ring[NSTAGE][STAGE_BYTES];
  uint8_t in[NIN][TILE_BYTES];
  uint8_t stage[2][TILE_BYTES];
  float vec[V_SIZE];
  uint64_t full[NSTAGE], empty[NSTAGE], in_full[NIN], in_empty[NIN];
  uint64_t a_ready, rb_ready[2], a_free, acc_done[2], acc_free[2], up_done[2], up_free[2], h_ready, h_free[2], x1_stored;
  float xchg[2][2][128];           // LayerNorm partial sums of the two column groups, double buffered
  uint32_t tmem_base;
};
constexpr size_t SMEM_BYTES = sizeof(Smem) + 1024;

// one lane of the (converged) warp; the compiler sees a warp-uniform predicate and issues tcgen05 ops without fix-up loops
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* barp, uint32_t parity) {
  const uint32_t bar = e4::smem_u32(barp);
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
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
// TMA: one [128 rows x 32 floats] box, global <-> 16 KB swizzled tile
__device__ __forceinline__ void tma_load_tile(uint32_t dst, const CUtensorMap* tm, int col, int row, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
      "l"(tm), "r"(col), "r"(row), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* tm, int col, int row, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(tm), "r"(col), "r"(row),
               "r"(src)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void group_barrier(int grp) {   // the 4 warps of one epilogue column group
  asm volatile("bar.sync %0, 128;\n" ::"r"(1 + grp) : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 3, 256;\n" ::: "memory"); }   // all 8 epilogue warps

// round-to-nearest (ties away) to tf32 precision for finite values: what cvt.rna.tf32.f32 computes, in two integer
// instructions -- the cvt expands to ~6 (inf / nan handling) and the split ran 3840 of them per row, a third of the
// epilogue warps' instruction stream (profiles/r1_tcpost_sass_mix.txt)
__device__ __forceinline__ float tf32_rna_finite(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// v (32 columns of this thread's row) -> tf32 hi / lo halves -> TMEM columns hi_col.. / lo_col.. of the thread's lane
__device__ __forceinline__ void split_store(uint32_t lane_base, uint32_t hi_col, uint32_t lo_col, const float (&v)[32]) {
  float hi[32], lo[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    hi[i] = tf32_rna_finite(v[i]);
    lo[i] = tf32_rna_finite(v[i] - hi[i]);
  }
  tmem_st32(lane_base + hi_col, hi);
  tmem_st32(lane_base + lo_col, lo);
}
// this thread's 32-float row segment of a swizzled tile: 16-byte chunk i of row r sits at chunk i ^ (r & 7)
__device__ __forceinline__ void tile_get32(const uint8_t* tile, int r, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(tile + r * 128 + ((i ^ (r & 7)) << 4));
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void tile_put32(uint8_t* tile, int r, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(tile + r * 128 + ((i ^ (r & 7)) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// 32 floats of a per-column vector (shared memory, same address for every thread: broadcast)
__device__ __forceinline__ void vec_get32(const float* p, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
// LayerNorm statistics of (acc + bias) over the 128 accumulator columns at `acc_col` of this thread's lane (two pass);
// bias may be NULL
__device__ __forceinline__ void ln_stats_tmem(uint32_t lane_base, uint32_t acc_col, const float* bias, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32], b[32];
    tc::tmem_ld32(lane_base + acc_col + c0, v);
    if (bias != nullptr) {
      vec_get32(bias + c0, b);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += b[i];
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
  }
  mean = s * (1.0f / 128.0f);
  float q = 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32], b[32];
    tc::tmem_ld32(lane_base + acc_col + c0, v);
    if (bias != nullptr) {
      vec_get32(bias + c0, b);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += b[i];
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float d = v[i] - mean;
      q = fmaf(d, d, q);
    }
  }
  rstd = 1.0f / sqrtf(q * (1.0f / 128.0f) + LN_EPS);
}

// phase timestamps of CTA 0 (SM clock): [0..15] epilogue thread 0, [16..31] MMA thread; read by prosim_tc_debug_read
__device__ long long g_tcp_dbg[32];
#define TCP_MARK(slot)                                           \
  do {                                                           \
    if (blockIdx.x == 0 && has_next) g_tcp_dbg[slot] = clock64(); \
  } while (0)

__global__ void __launch_bounds__(THREADS, 1)
    attn_post_tc_kernel(const __grid_constant__ Maps maps, int N, int zd, const float* __restrict__ W,
                        const float* __restrict__ Wn) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (e4::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * 128;
  const bool has_next = Wn != nullptr;

  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 512);
  if (tid == 32) {
    for (int i = 0; i < NSTAGE; ++i) {
      e4::mbar_init(e4::smem_u32(&sm.full[i]), 1);
      e4::mbar_init(e4::smem_u32(&sm.empty[i]), 1);
    }
    for (int i = 0; i < NIN; ++i) {
      e4::mbar_init(e4::smem_u32(&sm.in_full[i]), 1);
      e4::mbar_init(e4::smem_u32(&sm.in_empty[i]), 128);
    }
    e4::mbar_init(e4::smem_u32(&sm.a_ready), 256);
    e4::mbar_init(e4::smem_u32(&sm.rb_ready[0]), 256);
    e4::mbar_init(e4::smem_u32(&sm.rb_ready[1]), 256);
    e4::mbar_init(e4::smem_u32(&sm.a_free), 1);
    e4::mbar_init(e4::smem_u32(&sm.h_ready), 128);
    e4::mbar_init(e4::smem_u32(&sm.x1_stored), 2);
    for (int i = 0; i < 2; ++i) {
      e4::mbar_init(e4::smem_u32(&sm.h_free[i]), 1);
      e4::mbar_init(e4::smem_u32(&sm.acc_done[i]), 1);
      e4::mbar_init(e4::smem_u32(&sm.acc_free[i]), 256);
      e4::mbar_init(e4::smem_u32(&sm.up_done[i]), 1);
      e4::mbar_init(e4::smem_u32(&sm.up_free[i]), 128);
    }
  }
  // per-column vectors -> shared memory (one coalesced pass)
  for (int i = tid; i < V_SIZE; i += THREADS) {
    float v = 0.f;
    if (i < V_B1) v = W[aw::BO + i];                              // BO, LN_POST_G/B, LN_FFPRE_G/B are contiguous
    else if (i < V_B2) v = W[aw::B1 + (i - V_B1)];
    else if (i < V_LNDST_G) v = W[aw::B2 + (i - V_B2)];           // B2, LN_FFPOST_G/B are contiguous
    else if (has_next) {
      if (i < V_BQ) v = Wn[aw::LN_DST_G + (i - V_LNDST_G)];       // LN_DST_G/B are contiguous
      else if (i < V_BS) v = Wn[aw::BQ + (i - V_BQ)];
      else if (i < V_BG) v = Wn[aw::BS + (i - V_BS)];
      else v = Wn[aw::BG + (i - V_BG)];
    }
    sm.vec[i] = v;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  const int nz = zd >> 5;                            // 32-column tiles per Rbar head
  const int nbuf = zd == 96 ? 2 : 1;                 // alternating Rbar_h operand buffers
  const uint32_t rb_base = zd == 96 ? 128u : AHI;    // buffer b: hi at rb_base + 2 zd b, lo at + zd
  const int vr_floats = 16 * zd * 2;                 // one head's Wvr' chunk (hi | lo)
  const int n_chunks = has_next ? 68 : 48;

  if (warp == 8) {
    // ================================================================== weight producer
    if (lane == 0) {
      for (int i = 0; i < n_chunks; ++i) {
        const float* src;
        uint32_t bytes;
        if (i < 8) {
          src = W + (zd == 96 ? aw::TC_VR96 : aw::TC_VR128) + i * vr_floats;
          bytes = vr_floats * 4;
        } else if (i < 48) {
          src = W + aw::TC_GA + (i - 8) * 8192;    // Wga (4), Wo (4), FFN (32): contiguous in consumption order
          bytes = 32768;
        } else if (i < 56) {
          src = Wn + aw::TC_S + (i - 48) * 8192;   // next layer's Ws, Wgx (contiguous) ...
          bytes = 32768;
        } else if (i < 60) {
          src = Wn + aw::TC_Q + (i - 56) * 8192;   // ... then Wq
          bytes = 32768;
        } else {
          src = Wn + aw::TC_KRG + (i - 60) * 4096;
          bytes = 16384;
        }
        const int s = i % NSTAGE;
        if (i >= NSTAGE) mbar_wait(&sm.empty[s], ((i / NSTAGE) - 1) & 1);
        const uint32_t fb = e4::smem_u32(&sm.full[s]);
        e4::mbar_expect_tx(fb, bytes);
        e4::bulk_copy(e4::smem_u32(sm.ring[s]), src, bytes, fb);
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // ================================================================== input tile producer (consumption order!)
    if (lane == 0) {
      const int n_rbar = 8 * nz;
      const int n_tiles = n_rbar + 4 + 12 + 4 + 4;
      for (int i = 0; i < n_tiles; ++i) {
        const CUtensorMap* tm;
        int col;
        int k = i;
        if (k < n_rbar) {
          tm = &maps.rbar;
          col = (k / nz) * zd + (k % nz) * 32;
        } else if ((k -= n_rbar) < 4) {
          tm = &maps.aggv;
          col = k * 32;
        } else if ((k -= 4) < 12) {   // gate: AggV, Gx, S of chunks (0, 1) then (2, 3), pairwise so that tile parity = chunk parity
          const int which = (k % 6) >> 1;
          tm = which == 0 ? &maps.aggv : which == 1 ? &maps.gx : &maps.s;
          col = ((k / 6) * 2 + (k & 1)) * 32;
        } else if ((k -= 12) < 4) {
          tm = &maps.x;
          col = k * 32;
        } else {
          k -= 4;
          if (k == 0) mbar_wait(&sm.x1_stored, 0);   // x1 was parked in the output buffer by the epilogue warps
          tm = &maps.out;
          col = k * 32;
        }
        const int kg = i >> 1;                              // tile i is the kg-th tile of group i & 1
        const int s = (i & 1) * NIN_G + kg % NIN_G;
        if (kg >= NIN_G) mbar_wait(&sm.in_empty[s], ((kg / NIN_G) - 1) & 1);
        const uint32_t fb = e4::smem_u32(&sm.in_full[s]);
        e4::mbar_expect_tx(fb, TILE_BYTES);
        tma_load_tile(e4::smem_u32(sm.in[s]), tm, col, row0, fb);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ================================================================== MMA issuer (whole warp runs the control flow,
    // one elected lane issues the tcgen05 instructions)
    const bool leader = elect_one();
    int ci = 0;
    uint32_t ph_a = 0, ph_rb0 = 0, ph_rb1 = 0, ph_af0 = 0, ph_af1 = 0, ph_uf0 = 0, ph_uf1 = 0, ph_h = 0;
    long long w_full = 0, w_other = 0;   // cycles the issuer spent waiting for weights / for the epilogue warps
    auto wait_bar = [&](uint64_t* bar, uint32_t& ph) {
      const long long t0 = clock64();
      mbar_wait(bar, ph);
      w_other += clock64() - t0;
      ph ^= 1;
      tc::fence_after_sync();
    };
    auto commit = [&](uint64_t* bar) {
      if (leader) tc::mma_commit(bar);
    };
    // one ring chunk = B operand [N x KC] (hi | lo); A = TMEM columns a_hi.. / a_lo.. ; D (+)= A . B^T
    auto gemm_chunk = [&](auto n_tag, auto kc_tag, uint32_t d_col, uint32_t a_hi, uint32_t a_lo, bool accumulate) {
      constexpr int N_ = decltype(n_tag)::value, KC_ = decltype(kc_tag)::value;
      const int s = ci % NSTAGE;
      {
        const long long t0 = clock64();
        mbar_wait(&sm.full[s], (ci / NSTAGE) & 1);
        w_full += clock64() - t0;
      }
      tc::fence_after_sync();
      const uint32_t sb = e4::smem_u32(sm.ring[s]);
      constexpr uint32_t idesc = tc::make_idesc_tf32(128, N_);
      const uint64_t bh0 = tc::make_smem_desc(sb, 128, KC_ * 32);
      const uint64_t bl0 = tc::make_smem_desc(sb + N_ * KC_ * 4, 128, KC_ * 32);
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < KC_ / 8; ++ks) {
          const uint64_t bh = bh0 + (uint64_t)(ks * 16), bl = bl0 + (uint64_t)(ks * 16);   // +256 B in 16-byte units
          mma_ts(tmem + d_col, tmem + a_lo + ks * 8, bh, idesc, accumulate || ks > 0);
          mma_ts(tmem + d_col, tmem + a_hi + ks * 8, bl, idesc, true);
          mma_ts(tmem + d_col, tmem + a_hi + ks * 8, bh, idesc, true);
        }
        tc::mma_commit(&sm.empty[s]);   // the stage is free once these MMAs have read it
      }
      __syncwarp();
      ++ci;
    };
    using I16 = std::integral_constant<int, 16>;
    using I32 = std::integral_constant<int, 32>;
    using I96 = std::integral_constant<int, 96>;
    using I128 = std::integral_constant<int, 128>;
    auto gemm128 = [&](uint32_t d_col) {   // full 128 -> 128 projection of the A operand at AHI / ALO
#pragma unroll 1
      for (int c = 0; c < 4; ++c) gemm_chunk(I128{}, I32{}, d_col, AHI + 32 * c, ALO + 32 * c, c > 0);
    };
    TCP_MARK(16);
    // 1. agg: eight heads, A = Rbar_h staged by the epilogue warps in alternating buffers
#pragma unroll 1
    for (int h = 0; h < H; ++h) {
      const uint32_t base = rb_base + (uint32_t)((h % nbuf) * 2 * zd);
      // one barrier per operand buffer: a fast epilogue warp may already be arriving for head h + 1 while a slow one
      // has not arrived for head h, and the two must not be counted into the same phase
      if (h % nbuf) wait_bar(&sm.rb_ready[1], ph_rb1);
      else wait_bar(&sm.rb_ready[0], ph_rb0);
      if (zd == 96) gemm_chunk(I16{}, I96{}, ACC0 + 16 * h, base, base + 96, false);
      else gemm_chunk(I16{}, I128{}, ACC0 + 16 * h, base, base + 128, false);
      if (h < H - nbuf) commit(&sm.a_free);
      if (h == H - 1) commit(&sm.acc_done[0]);
    }
    TCP_MARK(17);
    // 2. gate (A = agg) -> ACC1 ; 3. out projection (A = u) -> ACC0
    wait_bar(&sm.a_ready, ph_a);
    TCP_MARK(18);
    gemm128(ACC1);
    commit(&sm.acc_done[1]);
    TCP_MARK(19);
    wait_bar(&sm.a_ready, ph_a);
    TCP_MARK(20);
    gemm128(ACC0);
    commit(&sm.acc_done[0]);
    TCP_MARK(21);
    // 4. FFN in 32-column hidden chunks: up_j -> UP[j&1]; the epilogue turns it into H; down_j accumulates into ACC1
    wait_bar(&sm.a_ready, ph_a);
    TCP_MARK(22);
    if (leader && blockIdx.x == 0) { g_tcp_dbg[28] = w_full; g_tcp_dbg[29] = w_other; }
    // up-projections run two chunks ahead of the down-projections, so the epilogue's relu / split of chunk j is
    // covered by up_{j+1} and up_{j+2} on the tensor pipe
    gemm_chunk(I32{}, I128{}, UP0, AHI, ALO, false);
    commit(&sm.up_done[0]);
    gemm_chunk(I32{}, I128{}, UP0 + 32, AHI, ALO, false);
    commit(&sm.up_done[1]);
#pragma unroll 1
    for (int j = 0; j < 16; ++j) {
      if (j + 2 < 16) {
        const int b = j & 1;                       // UP[b] is free once the epilogue of chunk j has read it
        if (b) wait_bar(&sm.up_free[1], ph_uf1);
        else wait_bar(&sm.up_free[0], ph_uf0);
        gemm_chunk(I32{}, I128{}, UP0 + 32 * b, AHI, ALO, false);
        commit(&sm.up_done[b]);
      }
      wait_bar(&sm.h_ready, ph_h);
      gemm_chunk(I128{}, I32{}, ACC1, HHI, HLO, j > 0);
      commit(j < 15 ? &sm.h_free[(j + 1) & 1] : &sm.acc_done[1]);   // H is free for the group that owns chunk j + 1
    }
    TCP_MARK(23);
    if (leader && blockIdx.x == 0) { g_tcp_dbg[30] = w_full; g_tcp_dbg[31] = w_other; }
    if (has_next) {
      // 5. next layer's destination-side projections: s -> ACC0, gx -> ACC1, q -> ACC0, then Qhat_h alternating
      wait_bar(&sm.a_ready, ph_a);                 // A = LN_dst'(out); ACC0 (out stash) and ACC1 (y) are consumed
      TCP_MARK(24);
      gemm128(ACC0);                               // s
      commit(&sm.acc_done[0]);
      gemm128(ACC1);                               // gx
      commit(&sm.acc_done[1]);
      wait_bar(&sm.acc_free[0], ph_af0);
      gemm128(ACC0);                               // q
      commit(&sm.acc_done[0]);
      TCP_MARK(25);
      wait_bar(&sm.a_ready, ph_a);                 // A = q (split by the q epilogue)
      TCP_MARK(26);
#pragma unroll 1
      for (int h = 0; h < H; ++h) {
        const int b = (h & 1) ^ 1;                 // h = 0 -> ACC1 (gx consumed), h = 1 -> ACC0 (q consumed), ...
        if (b) wait_bar(&sm.acc_free[1], ph_af1);
        else wait_bar(&sm.acc_free[0], ph_af0);
        gemm_chunk(I128{}, I16{}, b ? ACC1 : ACC0, AHI + 16 * h, ALO + 16 * h, false);
        commit(&sm.acc_done[b]);
      }
      TCP_MARK(27);
    }
  } else {
    // ================================================================== epilogue warps: thread = (row, column group)
    // group g = warp / 4 owns the 32-column chunks {g, g + 2} of every 128-wide row and the input tiles of parity g
    const int grp = warp >> 2;
    const int r = tid & 127;                            // row = TMEM lane
    const uint32_t lb = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t a_ready = e4::smem_u32(&sm.a_ready);
    const float* vec = sm.vec;
    uint32_t ph_done0 = 0, ph_done1 = 0, ph_afree = 0, ph_up = 0, ph_hfree = 0;
    int in_k = 0, xc = 0;                               // tiles this group has consumed
    const bool issuer = r == 0;
    auto publish_a = [&]() {
      tmem_st_wait();
      tc::fence_before_sync();
      mbar_arrive(a_ready);
    };
    auto wait_acc = [&](int b) {
      if (b) { mbar_wait(&sm.acc_done[1], ph_done1); ph_done1 ^= 1; }
      else { mbar_wait(&sm.acc_done[0], ph_done0); ph_done0 ^= 1; }
      tc::fence_after_sync();
    };
    // next input tile of this group -> this thread's 32 floats; the slot is handed back to the producer
    auto in_get = [&](float (&dst)[32]) {
      const int s = grp * NIN_G + in_k % NIN_G;
      mbar_wait(&sm.in_full[s], (in_k / NIN_G) & 1);
      tile_get32(sm.in[s], r, dst);
      mbar_arrive(e4::smem_u32(&sm.in_empty[s]));
      ++in_k;
    };
    // this thread's 32 floats -> the group's staging tile -> TMA store of the [128 x 32] box at column col
    auto out_put = [&](const CUtensorMap* tm, int col, const float (&src)[32]) {
      if (issuer) tma_store_wait_read<0>();     // the group's previous store has read the tile out
      group_barrier(grp);
      tile_put32(sm.stage[grp], r, src);
      e4::fence_proxy_async();
      group_barrier(grp);
      if (issuer) tma_store_tile(tm, col, row0, e4::smem_u32(sm.stage[grp]));
    };
    // phase 5 only: every input tile has been consumed, so the group's two input slots join its staging tile and
    // three stores can be in flight (a single tile serialises on the store's shared-memory read: 1.5 k cycles per chunk)
    int rot = 0;
    auto out_put3 = [&](auto issue, const float (&src)[32]) {
      uint8_t* tile = rot == 2 ? sm.stage[grp] : sm.in[grp * NIN_G + rot];
      rot = rot == 2 ? 0 : rot + 1;
      if (issuer) tma_store_wait_read<2>();
      group_barrier(grp);
      tile_put32(tile, r, src);
      e4::fence_proxy_async();
      group_barrier(grp);
      if (issuer) issue(e4::smem_u32(tile));
    };
    // row total of a per-thread partial (this group's 64 columns + the other group's)
    auto row_sum = [&](float part) -> float {
      float* x = &sm.xchg[xc & 1][0][0];
      x[grp * 128 + r] = part;
      epi_barrier();
      const float tot = x[r] + x[128 + r];
      ++xc;
      return tot;
    };
    float v[32], t[32];
#define EPI_MARK(slot) do { if (tid == 0) TCP_MARK(slot); } while (0)
    EPI_MARK(0);
    // 1. Rbar_h -> A, head by head (alternating buffers); tiles of parity grp
    {
      int tl = grp;                                      // this group's next Rbar tile
#pragma unroll 1
      for (int h = 0; h < H; ++h) {
        if (h >= nbuf) {
          mbar_wait(&sm.a_free, ph_afree);
          ph_afree ^= 1;
          tc::fence_after_sync();
        }
        const uint32_t base = rb_base + (uint32_t)((h % nbuf) * 2 * zd);
#pragma unroll 1
        for (; tl < (h + 1) * nz; tl += 2) {
          const int c0 = (tl - h * nz) * 32;
          in_get(v);
          split_store(lb, base + c0, base + zd + c0, v);
        }
        tmem_st_wait();
        tc::fence_before_sync();
        mbar_arrive(e4::smem_u32(&sm.rb_ready[h % nbuf]));
      }
    }
    //    agg = ACC0 + AggV -> A
    EPI_MARK(1);
    wait_acc(0);
    EPI_MARK(2);
#pragma unroll 1
    for (int c0 = grp * 32; c0 < 128; c0 += 64) {
      tc::tmem_ld32(lb + ACC0 + c0, v);
      in_get(t);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += t[i];
      split_store(lb, AHI + c0, ALO + c0, v);
    }
    publish_a();
    // 2. gate: g = sigmoid(ACC1 + Gx) ; u = agg + g (S - agg) -> A      (agg re-read from ACC0 + AggV)
    EPI_MARK(3);
    wait_acc(1);
    EPI_MARK(4);
#pragma unroll 1
    for (int c0 = grp * 32; c0 < 128; c0 += 64) {
      float a[32];
      tc::tmem_ld32(lb + ACC0 + c0, a);
      in_get(t);
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] += t[i];
      tc::tmem_ld32(lb + ACC1 + c0, v);
      in_get(t);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 1.0f / (1.0f + expf(-(v[i] + t[i])));
      in_get(t);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = a[i] + v[i] * (t[i] - a[i]);
      split_store(lb, AHI + c0, ALO + c0, v);
    }
    publish_a();
    // 3. o = ACC0 + bo ; x1 = x + LN_post(o) -> parked in the output buffer ; LN_ffpre(x1) -> A
    EPI_MARK(5);
    wait_acc(0);
    EPI_MARK(6);
    {
      float part = 0.f;
#pragma unroll 1
      for (int c0 = grp * 32; c0 < 128; c0 += 64) {
        tc::tmem_ld32(lb + ACC0 + c0, v);
        vec_get32(vec + V_BO + c0, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) part += v[i] + t[i];
      }
      float mean = row_sum(part) * (1.0f / 128.0f);
      part = 0.f;
#pragma unroll 1
      for (int c0 = grp * 32; c0 < 128; c0 += 64) {
        tc::tmem_ld32(lb + ACC0 + c0, v);
        vec_get32(vec + V_BO + c0, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = (v[i] + t[i]) - mean;
          part = fmaf(d, d, part);
        }
      }
      float rstd = 1.0f / sqrtf(row_sum(part) * (1.0f / 128.0f) + LN_EPS);
      // x1 goes back into ACC0 (its o columns are dead once read) so that the LN_ffpre statistics need no second copy
      part = 0.f;
#pragma unroll 1
      for (int c0 = grp * 32; c0 < 128; c0 += 64) {
        float g[32];
        tc::tmem_ld32(lb + ACC0 + c0, v);
        vec_get32(vec + V_BO + c0, t);
        vec_get32(vec + V_LNPOST_G + c0, g);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = ((v[i] + t[i]) - mean) * rstd * g[i];
        vec_get32(vec + V_LNPOST_B + c0, t);
        in_get(g);                                  // x
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] = g[i] + (v[i] + t[i]);
          part += v[i];
        }
        tmem_st32(lb + ACC0 + c0, v);
        out_put(&maps.out, c0, v);
      }
      tmem_st_wait();
      mean = row_sum(part) * (1.0f / 128.0f);
      part = 0.f;
#pragma unroll 1
      for (int c0 = grp * 32; c0 < 128; c0 += 64) {
        tc::tmem_ld32(lb + ACC0 + c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = v[i] - mean;
          part = fmaf(d, d, part);
        }
      }
      rstd = 1.0f / sqrtf(row_sum(part) * (1.0f / 128.0f) + LN_EPS);
#pragma unroll 1
      for (int c0 = grp * 32; c0 < 128; c0 += 64) {
        float g[32];
        tc::tmem_ld32(lb + ACC0 + c0, v);
        vec_get32(vec + V_LNFFPRE_G + c0, g);
        vec_get32(vec + V_LNFFPRE_B + c0, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd * g[i] + t[i];
        split_store(lb, AHI + c0, ALO + c0, v);
      }
    }
    publish_a();
    if (issuer) {
      tma_store_wait<0>();                          // x1 is in global memory: the input producer may fetch it again
      mbar_arrive(e4::smem_u32(&sm.x1_stored));
    }
    // 4. FFN hidden chunks j = grp, grp + 2, ...: h_j = relu(UP[grp] + b1_j) -> H (hi | lo)
    EPI_MARK(7);
#pragma unroll 1
    for (int j = grp; j < 16; j += 2) {
      mbar_wait(&sm.up_done[grp], ph_up);
      ph_up ^= 1;
      tc::fence_after_sync();
      tc::tmem_ld32(lb + UP0 + 32 * grp, v);
      tc::fence_before_sync();
      mbar_arrive(e4::smem_u32(&sm.up_free[grp]));
      vec_get32(vec + V_B1 + 32 * j, t);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + t[i], 0.f);
      if (j > 0) {
        mbar_wait(&sm.h_free[grp], ph_hfree);     // down_{j-1} has consumed H
        ph_hfree ^= 1;
        tc::fence_after_sync();
      }
      split_store(lb, HHI, HLO, v);
      tmem_st_wait();
      tc::fence_before_sync();
      mbar_arrive(e4::smem_u32(&sm.h_ready));
    }
    //    y = ACC1 + b2 ; out = x1 + LN_ffpost(y) -> global and ACC0 ; LN_dst'(out) -> A
    EPI_MARK(8);
    wait_acc(1);
    EPI_MARK(9);
    {
      float part = 0.f;
#pragma unroll 1
      for (int c0 = grp * 32; c0 < 128; c0 += 64) {
        tc::tmem_ld32(lb + ACC1 + c0, v);
        vec_get32(vec + V_B2 + c0, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) part += v[i] + t[i];
      }
      float mean = row_sum(part) * (1.0f / 128.0f);
      part = 0.f;
#pragma unroll 1
      for (int c0 = grp * 32; c0 < 128; c0 += 64) {
        tc::tmem_ld32(lb + ACC1 + c0, v);
        vec_get32(vec + V_B2 + c0, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = (v[i] + t[i]) - mean;
          part = fmaf(d, d, part);
        }
      }
      float rstd = 1.0f / sqrtf(row_sum(part) * (1.0f / 128.0f) + LN_EPS);
      part = 0.f;
#pragma unroll 1
      for (int c0 = grp * 32; c0 < 128; c0 += 64) {
        float g[32];
        tc::tmem_ld32(lb + ACC1 + c0, v);
        vec_get32(vec + V_B2 + c0, t);
        vec_get32(vec + V_LNFFPOST_G + c0, g);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = ((v[i] + t[i]) - mean) * rstd * g[i];
        vec_get32(vec + V_LNFFPOST_B + c0, t);
        in_get(g);                                  // x1
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] = g[i] + (v[i] + t[i]);
          part += v[i];
        }
        if (has_next) tmem_st32(lb + ACC0 + c0, v);
        out_put(&maps.out, c0, v);
      }
      if (has_next) {
        tmem_st_wait();
        mean = row_sum(part) * (1.0f / 128.0f);
        part = 0.f;
#pragma unroll 1
        for (int c0 = grp * 32; c0 < 128; c0 += 64) {
          tc::tmem_ld32(lb + ACC0 + c0, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float d = v[i] - mean;
            part = fmaf(d, d, part);
          }
        }
        rstd = 1.0f / sqrtf(row_sum(part) * (1.0f / 128.0f) + LN_EPS);
#pragma unroll 1
        for (int c0 = grp * 32; c0 < 128; c0 += 64) {
          float g[32];
          tc::tmem_ld32(lb + ACC0 + c0, v);
          vec_get32(vec + V_LNDST_G + c0, g);
          vec_get32(vec + V_LNDST_B + c0, t);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd * g[i] + t[i];
          split_store(lb, AHI + c0, ALO + c0, v);
        }
        publish_a();
      }
    }
    EPI_MARK(10);
    if (has_next) {
      // 5. s (ACC0), gx (ACC1), q (ACC0; also split into A), then the eight Qhat_h (ACC1, ACC0, ...)
      auto out_proj = [&](int b, int bias, const CUtensorMap* tm, bool to_a) {
        wait_acc(b);
#pragma unroll 1
        for (int c0 = grp * 32; c0 < 128; c0 += 64) {
          tc::tmem_ld32(lb + (b ? ACC1 : ACC0) + c0, v);
          vec_get32(vec + bias + c0, t);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += t[i];
          if (to_a) split_store(lb, AHI + c0, ALO + c0, v);   // every MMA that read A = LN_dst'(out) has completed
          out_put3([&](uint32_t src) { tma_store_tile(tm, c0, row0, src); }, v);
        }
        tc::fence_before_sync();
        mbar_arrive(e4::smem_u32(&sm.acc_free[b]));
      };
      out_proj(0, V_BS, &maps.s_n, false);
      out_proj(1, V_BG, &maps.gx_n, false);
      out_proj(0, V_BQ, &maps.q_n, true);
      publish_a();
      EPI_MARK(11);
#pragma unroll 1
      for (int h = 0; h < H; ++h) {
        const int b = (h & 1) ^ 1;
        wait_acc(b);
#pragma unroll 1
        for (int c0 = grp * 32; c0 < 128; c0 += 64) {
          tc::tmem_ld32(lb + (b ? ACC1 : ACC0) + c0, v);
          out_put3([&](uint32_t src) { tma_store_tile(&maps.qhat_n, h * D + c0, row0, src); }, v);
        }
        tc::fence_before_sync();
        mbar_arrive(e4::smem_u32(&sm.acc_free[b]));
      }
    }
    EPI_MARK(12);
    if (issuer) tma_store_wait<0>();   // all output boxes have landed before the CTA retires
    EPI_MARK(13);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace tcp
}  // namespace prosim
