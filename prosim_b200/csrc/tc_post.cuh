// Node side of the AttentionLayer on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-equivalent via 3xTF32.
//
// Same math as attn_post2_kernel (attn2.cuh; reference prosim/models/layers/attention_layer.py:102-118, :38-43 and the
// next layer's destination-side projections :56-66), one CTA per 128 destination rows:
//   agg = AggV + Rbar_h . Wvr'_h        8 x [128 x zd] . [zd x 16]
//   g   = sigmoid(agg . Wga + Gx) ;  u = agg + g (S - agg)
//   x1  = x + LN(u . Wo + bo) ;  y = W2 relu(W1 LN(x1) + b1) + b2 ;  out = x1 + LN(y)
//   next layer:  q | s | gx = LN_dst'(out) . {Wq, Ws, Wgx} ;  Qhat_h = q_h . (diag(gamma_r) Wkr_h)
// Every product a*w is evaluated as a_lo*w_hi + a_hi*w_lo + a_hi*w_hi with hi = tf32(v), lo = tf32(v - hi)
// (relative error ~2^-21 per product, fp32 accumulate in tensor memory) -- the per-tick parity gate of 1e-5 needs
// fp32-class arithmetic, a single TF32 pass (2^-11) fails it (tests/test_gpu_kernels.py::test_tc_gemm).
//
// Structure (192 threads, 1 CTA/SM, all 512 TMEM columns):
//   warps 0-3  epilogue: thread = row = TMEM lane.  tcgen05.ld the accumulator, apply bias / gate / LayerNorm /
//              residual in registers (a LayerNorm is thread local: no shuffles), split the result into hi/lo and
//              tcgen05.st it back as the A operand of the next GEMM -- activations never touch shared memory
//   warp 4     weight producer: one thread streams the layer's pre-split B operands (aw::TC_*, weights.py) from L2
//              through a 4 x 32 KB shared-memory ring with cp.async.bulk on full/empty mbarriers
//   warp 5     MMA issuer: one thread issues tcgen05.mma kind::tf32 with A from TMEM, B from the ring; tcgen05.commit
//              releases ring stages and publishes accumulators
// TMEM columns: [0,128) ACC0, [128,256) ACC1, [256,384) A hi, [384,512) A lo.  During the FFN the hidden layer is
// produced and consumed in 32-column chunks (ACC0 is re-used as 2 x 32 up-accumulators + 32 hi + 32 lo hidden
// columns) so that the 512-wide hidden activation never has to exist anywhere.
#pragma once
#include "common.cuh"
#include "edge4.cuh"     // e4:: mbarrier / bulk-copy helpers
#include "tc_gemm.cuh"   // tc:: tcgen05 helpers
#include "weights_layout.h"

namespace prosim {
namespace tcp {

constexpr int NSTAGE = 4;
constexpr int STAGE_BYTES = 32768;
constexpr int THREADS = 192;
constexpr uint32_t ACC0 = 0, ACC1 = 128, AHI = 256, ALO = 384;
constexpr uint32_t UP0 = 0, HHI = 64, HLO = 96;   // inside ACC0 during the FFN: up accumulators at 0 / 32

struct Smem {
  uint8_t ring[NSTAGE][STAGE_BYTES];
  float stash[128 * 128];          // one fp32 row per thread (agg, then x1, then out / q), float4-swizzled
  uint64_t full[NSTAGE], empty[NSTAGE];
  uint64_t a_ready, a_free, acc_done[2], acc_free[2], up_done[2], up_free[2], h_ready, h_free;
  uint32_t tmem_base;
};
constexpr size_t SMEM_BYTES = sizeof(Smem) + 1024;

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

// v (32 columns of this thread's row) -> tf32 hi / lo halves -> TMEM columns hi_col.. / lo_col.. of the thread's lane
__device__ __forceinline__ void split_store(uint32_t lane_base, uint32_t hi_col, uint32_t lo_col, const float (&v)[32]) {
  float hi[32], lo[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    hi[i] = tc::to_tf32(v[i]);
    lo[i] = tc::to_tf32(v[i] - hi[i]);
  }
  tmem_st32(lane_base + hi_col, hi);
  tmem_st32(lane_base + lo_col, lo);
}

// stash[row][c4] at float4 granularity, XOR-swizzled so that a warp's 32 rows hit 32 different bank groups
__device__ __forceinline__ float4* stash_ptr(float* stash, int r, int c4) {
  return reinterpret_cast<float4*>(stash) + r * 32 + (c4 ^ (r & 31));
}
__device__ __forceinline__ void stash_put32(float* stash, int r, int c0, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) *stash_ptr(stash, r, (c0 >> 2) + i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void stash_get32(const float* stash, int r, int c0, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = *stash_ptr(const_cast<float*>(stash), r, (c0 >> 2) + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
// 32 consecutive floats of a global row (zeros when !ok) / of a per-column vector shared by all rows
__device__ __forceinline__ void row_get32(const float* __restrict__ p, bool ok, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) t = *reinterpret_cast<const float4*>(p + 4 * i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void row_put32(float* __restrict__ p, bool ok, const float (&v)[32]) {
  if (!ok) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void vec_get32(const float* __restrict__ p, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

// LayerNorm statistics of (acc + bias) over the 128 accumulator columns at `acc_col` of this thread's lane (two pass)
__device__ __forceinline__ void ln_stats_tmem(uint32_t lane_base, uint32_t acc_col, const float* __restrict__ bias,
                                              float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32], b[32];
    tc::tmem_ld32(lane_base + acc_col + c0, v);
    vec_get32(bias + c0, b);
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i] + b[i];
  }
  mean = s * (1.0f / 128.0f);
  float q = 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32], b[32];
    tc::tmem_ld32(lane_base + acc_col + c0, v);
    vec_get32(bias + c0, b);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float d = (v[i] + b[i]) - mean;
      q = fmaf(d, d, q);
    }
  }
  rstd = 1.0f / sqrtf(q * (1.0f / 128.0f) + LN_EPS);
}
// LayerNorm statistics of the thread's stashed row
__device__ __forceinline__ void ln_stats_stash(const float* stash, int r, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32];
    stash_get32(stash, r, c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
  }
  mean = s * (1.0f / 128.0f);
  float q = 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32];
    stash_get32(stash, r, c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float d = v[i] - mean;
      q = fmaf(d, d, q);
    }
  }
  rstd = 1.0f / sqrtf(q * (1.0f / 128.0f) + LN_EPS);
}

__global__ void __launch_bounds__(THREADS, 1)
    attn_post_tc_kernel(const float* __restrict__ Xdst, int N, int zd, const float* __restrict__ Rbar,
                        const float* __restrict__ AggV, const float* __restrict__ Sg, const float* __restrict__ Gxg,
                        const float* __restrict__ W, float* __restrict__ Out, const float* __restrict__ Wn,
                        float* __restrict__ Qg_n, float* __restrict__ Qhat_n, float* __restrict__ Sg_n,
                        float* __restrict__ Gxg_n) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (e4::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * 128;

  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 512);
  if (tid == 32) {
    for (int i = 0; i < NSTAGE; ++i) {
      e4::mbar_init(e4::smem_u32(&sm.full[i]), 1);
      e4::mbar_init(e4::smem_u32(&sm.empty[i]), 1);
    }
    e4::mbar_init(e4::smem_u32(&sm.a_ready), 128);
    e4::mbar_init(e4::smem_u32(&sm.a_free), 1);
    e4::mbar_init(e4::smem_u32(&sm.h_ready), 128);
    e4::mbar_init(e4::smem_u32(&sm.h_free), 1);
    for (int i = 0; i < 2; ++i) {
      e4::mbar_init(e4::smem_u32(&sm.acc_done[i]), 1);
      e4::mbar_init(e4::smem_u32(&sm.acc_free[i]), 128);
      e4::mbar_init(e4::smem_u32(&sm.up_done[i]), 1);
      e4::mbar_init(e4::smem_u32(&sm.up_free[i]), 128);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  const int vr_floats = 16 * zd * 2;                 // one head's Wvr' chunk (hi | lo)
  const int n_chunks = Wn != nullptr ? 68 : 48;

  if (warp == 4) {
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
        } else if (i < 60) {
          src = Wn + aw::TC_Q + (i - 48) * 8192;   // next layer's Wq, Ws, Wgx
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
  } else if (warp == 5) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      int ci = 0;
      uint32_t ph_a = 0, ph_af[2] = {0, 0}, ph_uf[2] = {0, 0}, ph_h = 0;
      auto wait_a = [&]() {
        mbar_wait(&sm.a_ready, ph_a);
        ph_a ^= 1;
        tc::fence_after_sync();
      };
      // one ring chunk = B operand [n x kc] (hi | lo); A = TMEM columns a_hi.. / a_lo.. ; D += A . B^T
      auto gemm_chunk = [&](int n, int kc, uint32_t d_col, uint32_t a_hi, uint32_t a_lo, bool accumulate) {
        const int s = ci % NSTAGE;
        mbar_wait(&sm.full[s], (ci / NSTAGE) & 1);
        tc::fence_after_sync();
        const uint32_t sb = e4::smem_u32(sm.ring[s]);
        const uint32_t idesc = tc::make_idesc_tf32(128, n);
        const uint32_t sbo = kc * 32, lo_off = n * kc * 4;
        for (int ks = 0; ks < kc / 8; ++ks) {
          const uint64_t bh = tc::make_smem_desc(sb + ks * 256, 128, sbo);
          const uint64_t bl = tc::make_smem_desc(sb + lo_off + ks * 256, 128, sbo);
          mma_ts(tmem + d_col, tmem + a_lo + ks * 8, bh, idesc, accumulate || ks > 0);
          mma_ts(tmem + d_col, tmem + a_hi + ks * 8, bl, idesc, true);
          mma_ts(tmem + d_col, tmem + a_hi + ks * 8, bh, idesc, true);
        }
        tc::mma_commit(&sm.empty[s]);   // the stage is free once these MMAs have read it
        ++ci;
      };
      // 1. agg: eight heads, A = Rbar_h (re-staged by the epilogue warps per head)
      for (int h = 0; h < H; ++h) {
        wait_a();
        gemm_chunk(16, zd, ACC0 + 16 * h, AHI, ALO, false);
        tc::mma_commit(h < H - 1 ? &sm.a_free : &sm.acc_done[0]);
      }
      // 2. gate (A = agg) -> ACC1 ; 3. out projection (A = u) -> ACC0
      wait_a();
      for (int c = 0; c < 4; ++c) gemm_chunk(128, 32, ACC1, AHI + 32 * c, ALO + 32 * c, c > 0);
      tc::mma_commit(&sm.acc_done[1]);
      wait_a();
      for (int c = 0; c < 4; ++c) gemm_chunk(128, 32, ACC0, AHI + 32 * c, ALO + 32 * c, c > 0);
      tc::mma_commit(&sm.acc_done[0]);
      // 4. FFN in 32-column hidden chunks: up_j -> UP[j&1]; the epilogue turns it into H; down_j accumulates into ACC1
      wait_a();
      gemm_chunk(32, 128, UP0, AHI, ALO, false);
      tc::mma_commit(&sm.up_done[0]);
      for (int j = 0; j < 16; ++j) {
        if (j + 1 < 16) {
          const int b = (j + 1) & 1;
          if (j + 1 >= 2) {
            mbar_wait(&sm.up_free[b], ph_uf[b]);
            ph_uf[b] ^= 1;
            tc::fence_after_sync();
          }
          gemm_chunk(32, 128, UP0 + 32 * b, AHI, ALO, false);
          tc::mma_commit(&sm.up_done[b]);
        }
        mbar_wait(&sm.h_ready, ph_h);
        ph_h ^= 1;
        tc::fence_after_sync();
        gemm_chunk(128, 32, ACC1, HHI, HLO, j > 0);
        tc::mma_commit(j < 15 ? &sm.h_free : &sm.acc_done[1]);
      }
      if (Wn != nullptr) {
        // 5. next layer's destination-side projections: q -> ACC0, s -> ACC1, gx -> ACC0, then Qhat_h alternating
        auto wait_free = [&](int b) {
          mbar_wait(&sm.acc_free[b], ph_af[b]);
          ph_af[b] ^= 1;
          tc::fence_after_sync();
        };
        wait_a();
        for (int c = 0; c < 4; ++c) gemm_chunk(128, 32, ACC0, AHI + 32 * c, ALO + 32 * c, c > 0);
        tc::mma_commit(&sm.acc_done[0]);
        for (int c = 0; c < 4; ++c) gemm_chunk(128, 32, ACC1, AHI + 32 * c, ALO + 32 * c, c > 0);
        tc::mma_commit(&sm.acc_done[1]);
        wait_free(0);
        for (int c = 0; c < 4; ++c) gemm_chunk(128, 32, ACC0, AHI + 32 * c, ALO + 32 * c, c > 0);
        tc::mma_commit(&sm.acc_done[0]);
        wait_a();   // A = q
        for (int h = 0; h < H; ++h) {
          wait_free(h & 1);
          gemm_chunk(128, 16, (h & 1) ? ACC1 : ACC0, AHI + 16 * h, ALO + 16 * h, false);
          tc::mma_commit(&sm.acc_done[h & 1]);
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================== epilogue warps: thread = row
    const int r = tid;                                  // 0..127 = TMEM lane
    const int row = row0 + r;
    const bool ok = row < N;
    const uint32_t lb = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t a_ready = e4::smem_u32(&sm.a_ready);
    uint32_t ph_done[2] = {0, 0}, ph_afree = 0, ph_up[2] = {0, 0}, ph_hfree = 0;
    auto publish_a = [&]() {
      tmem_st_wait();
      tc::fence_before_sync();
      mbar_arrive(a_ready);
    };
    auto wait_acc = [&](int b) {
      mbar_wait(&sm.acc_done[b], ph_done[b]);
      ph_done[b] ^= 1;
      tc::fence_after_sync();
    };
    float v[32], t[32];

    // 1. Rbar_h -> A, head by head
    const float* rbar = Rbar + (size_t)(ok ? row : 0) * H * zd;
#pragma unroll 1
    for (int h = 0; h < H; ++h) {
      if (h > 0) {
        mbar_wait(&sm.a_free, ph_afree);
        ph_afree ^= 1;
        tc::fence_after_sync();
      }
#pragma unroll 1
      for (int c0 = 0; c0 < zd; c0 += 32) {
        row_get32(rbar + h * zd + c0, ok, v);
        split_store(lb, AHI + c0, ALO + c0, v);
      }
      publish_a();
    }
    //    agg = ACC0 + AggV  -> stash, A
    wait_acc(0);
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      tc::tmem_ld32(lb + ACC0 + c0, v);
      row_get32(AggV + (size_t)(ok ? row : 0) * D + c0, ok, t);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += t[i];
      stash_put32(sm.stash, r, c0, v);
      split_store(lb, AHI + c0, ALO + c0, v);
    }
    publish_a();
    // 2. gate: g = sigmoid(ACC1 + Gx) ; u = agg + g (S - agg) -> A
    wait_acc(1);
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float a[32];
      tc::tmem_ld32(lb + ACC1 + c0, v);
      row_get32(Gxg + (size_t)(ok ? row : 0) * D + c0, ok, t);
      stash_get32(sm.stash, r, c0, a);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 1.0f / (1.0f + expf(-(v[i] + t[i])));
      row_get32(Sg + (size_t)(ok ? row : 0) * D + c0, ok, t);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = a[i] + v[i] * (t[i] - a[i]);
      split_store(lb, AHI + c0, ALO + c0, v);
    }
    publish_a();
    // 3. o = ACC0 + bo ; x1 = x + LN_post(o) -> stash ; LN_ffpre(x1) -> A
    wait_acc(0);
    {
      float mean, rstd;
      ln_stats_tmem(lb, ACC0, W + aw::BO, mean, rstd);
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float g[32];
        tc::tmem_ld32(lb + ACC0 + c0, v);
        vec_get32(W + aw::BO + c0, t);
        vec_get32(W + aw::LN_POST_G + c0, g);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = ((v[i] + t[i]) - mean) * rstd * g[i];
        vec_get32(W + aw::LN_POST_B + c0, t);
        row_get32(Xdst + (size_t)(ok ? row : 0) * D + c0, ok, g);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = g[i] + (v[i] + t[i]);
        stash_put32(sm.stash, r, c0, v);
      }
      ln_stats_stash(sm.stash, r, mean, rstd);
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float g[32];
        stash_get32(sm.stash, r, c0, v);
        vec_get32(W + aw::LN_FFPRE_G + c0, g);
        vec_get32(W + aw::LN_FFPRE_B + c0, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd * g[i] + t[i];
        split_store(lb, AHI + c0, ALO + c0, v);
      }
    }
    publish_a();
    // 4. FFN hidden chunks: h_j = relu(UP[j&1] + b1_j) -> H (hi | lo)
#pragma unroll 1
    for (int j = 0; j < 16; ++j) {
      const int b = j & 1;
      mbar_wait(&sm.up_done[b], ph_up[b]);
      ph_up[b] ^= 1;
      tc::fence_after_sync();
      tc::tmem_ld32(lb + UP0 + 32 * b, v);
      tc::fence_before_sync();
      mbar_arrive(e4::smem_u32(&sm.up_free[b]));
      vec_get32(W + aw::B1 + 32 * j, t);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + t[i], 0.f);
      if (j > 0) {
        mbar_wait(&sm.h_free, ph_hfree);
        ph_hfree ^= 1;
        tc::fence_after_sync();
      }
      split_store(lb, HHI, HLO, v);
      tmem_st_wait();
      tc::fence_before_sync();
      mbar_arrive(e4::smem_u32(&sm.h_ready));
    }
    //    y = ACC1 + b2 ; out = x1 + LN_ffpost(y) -> global, stash ; LN_dst'(out) -> A
    wait_acc(1);
    {
      float mean, rstd;
      ln_stats_tmem(lb, ACC1, W + aw::B2, mean, rstd);
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float g[32];
        tc::tmem_ld32(lb + ACC1 + c0, v);
        vec_get32(W + aw::B2 + c0, t);
        vec_get32(W + aw::LN_FFPOST_G + c0, g);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = ((v[i] + t[i]) - mean) * rstd * g[i];
        vec_get32(W + aw::LN_FFPOST_B + c0, t);
        stash_get32(sm.stash, r, c0, g);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = g[i] + (v[i] + t[i]);
        row_put32(Out + (size_t)(ok ? row : 0) * D + c0, ok, v);
        stash_put32(sm.stash, r, c0, v);
      }
      if (Wn != nullptr) {
        ln_stats_stash(sm.stash, r, mean, rstd);
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          float g[32];
          stash_get32(sm.stash, r, c0, v);
          vec_get32(Wn + aw::LN_DST_G + c0, g);
          vec_get32(Wn + aw::LN_DST_B + c0, t);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd * g[i] + t[i];
          split_store(lb, AHI + c0, ALO + c0, v);
        }
        publish_a();
      }
    }
    if (Wn != nullptr) {
      // 5. q (-> global, stash), s, gx, then q -> A and the eight Qhat_h
      auto out_proj = [&](int b, const float* bias, float* dst, bool keep) {
        wait_acc(b);
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          tc::tmem_ld32(lb + (b ? ACC1 : ACC0) + c0, v);
          vec_get32(bias + c0, t);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += t[i];
          row_put32(dst + (size_t)(ok ? row : 0) * D + c0, ok, v);
          if (keep) stash_put32(sm.stash, r, c0, v);
        }
        tc::fence_before_sync();
        mbar_arrive(e4::smem_u32(&sm.acc_free[b]));
      };
      out_proj(0, Wn + aw::BQ, Qg_n, true);
      out_proj(1, Wn + aw::BS, Sg_n, false);
      out_proj(0, Wn + aw::BG, Gxg_n, false);
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        stash_get32(sm.stash, r, c0, v);
        split_store(lb, AHI + c0, ALO + c0, v);
      }
      publish_a();
#pragma unroll 1
      for (int h = 0; h < H; ++h) {
        const int b = h & 1;
        wait_acc(b);
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          tc::tmem_ld32(lb + (b ? ACC1 : ACC0) + c0, v);
          row_put32(Qhat_n + (size_t)(ok ? row : 0) * H * D + h * D + c0, ok, v);
        }
        tc::fence_before_sync();
        mbar_arrive(e4::smem_u32(&sm.acc_free[b]));
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace tcp
}  // namespace prosim
