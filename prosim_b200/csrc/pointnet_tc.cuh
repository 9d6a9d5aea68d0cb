// PointNet polyline encoder on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-class accuracy via 3xTF32.
//
// Same math as pointnet_kernel (pointnet.cuh; reference prosim/models/scene_encoder/pointnet_encoder.py:24-62): the
// per-point GEMMs are genuinely dense ([points x 128] . [128 x 128], 67 k MAC per map point), the FFMA kernel runs them
// at 27-45 % of the fp32 FMA peak and cannot go further.  Here one CTA (128 threads, thread = point row = TMEM lane)
// owns G polylines of P points (<= 128 rows) and runs the whole chain
//   pre_mlps -> max-pool -> mlps.0 on [point | pooled] (K = 256, the reference's own concatenation) -> mlps.1 -> max-pool
//   -> out_mlps (on the G pooled rows, as M = 128 MMAs with the other rows zero)
// as a sequence of K = 32 chunks: the chunk's B operand (weights, pre-split into tf32 hi / lo and pre-arranged in the
// tcgen05 K-major core-matrix order by weights.py::pack_pointnet_tc, 32 KB) arrives by one cp.async.bulk into a 2-stage
// ring one chunk ahead; the A operand (this thread's 32 activations, split into hi / lo on the fly) is written to shared
// memory by the row's own thread; one thread issues 4 k-steps x 3 MMAs (a_lo*w_hi + a_hi*w_lo + a_hi*w_hi, fp32
// accumulate in tensor memory) and commits to an mbarrier.  Epilogues (bias, LayerNorm, ReLU, mask) are thread local:
// a thread reads its row's 128 accumulator columns with tcgen05.ld and keeps them in registers as the next A operand.
// 128 TMEM columns and ~106 KB of shared memory per CTA: two CTAs per SM overlap each other's epilogues and MMAs.
// Measured alternatives (round 1, per forward: PointNet / K'V' launches): this version 1.77 / 0.91 ms; a 4-stage weight
// ring with one CTA per SM 2.43 / 1.09 ms; 4 stages + two A buffers with deferred MMA waits 2.23 / 1.01 ms -- the chunk
// loop is bound by the per-thread epilogue / staging instruction streams (one warp per scheduler), so the second
// resident CTA is worth more than deeper pipelines inside one CTA.  (A policy-head variant on this pipeline was also
// tried: 96 us against 84 us for the FFMA kernel at 32 CTAs, and its 3xTF32 rounding fed the closed loop directly.)
#pragma once
#include "common.cuh"
#include "edge4.cuh"     // e4:: mbarrier / bulk-copy helpers
#include "tc_gemm.cuh"   // tc:: tcgen05 helpers
#include "weights_layout.h"

namespace prosim {
namespace pntc {

constexpr int KC = 32;                         // K per chunk
constexpr int A_BYTES = 128 * KC * 4;          // one of hi / lo: 16 KB
constexpr int B_STAGE_BYTES = 2 * 128 * KC * 4;   // hi | lo: 32 KB
constexpr int N_VEC = 13;                      // per-column vectors (biases, LayerNorm affine) staged in shared memory
constexpr int POOL_LD = 33;                    // pooling scratch [128 rows][32 + 1] floats inside the A buffers

template <int P>
struct Cfg {
  static constexpr int G = 128 / P;
  static constexpr int GP = (G + 3) & ~3;
  static constexpr size_t smem_bytes = 1024 + 2 * A_BYTES + 2 * B_STAGE_BYTES + (size_t)GP * 128 * 4 + 128 * 4 + 64 + N_VEC * 512;
};

// chunks of the packed tensor-core weight block (weights.py::pack_pointnet_tc), in consumption order
template <int NPRE>
__host__ __device__ constexpr int n_chunks() { return 1 + (NPRE == 3 ? 8 : 0) + 8 + 4 + 4 + 4; }

__device__ __forceinline__ void wait_bar(uint64_t* barp, uint32_t parity) {   // bounded: a protocol bug traps, never hangs
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
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ float tf32_rna_finite(float x) {   // cvt.rna.tf32.f32 for finite values, two integer instructions
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

struct Pipe {
  uint8_t* sAhi;
  uint8_t* sAlo;
  uint8_t* sB;            // [2][B_STAGE_BYTES]
  uint64_t* bfull;        // [2]
  uint64_t* bmma;         // [1]
  const float* wtc;       // packed chunks
  int wmul, woff;         // chunk c of this CTA is packed chunk c * wmul + woff (kv_tc.cuh: K' and V' chunks interleaved)
  uint32_t tmem;
  int chunk;              // chunks consumed so far (stage = chunk & 1, B parity = (chunk >> 1) & 1, MMA parity = chunk & 1)
  int total;
};

// thread 0: start the copy of chunk `c` into its stage
__device__ __forceinline__ void load_b(const Pipe& p, int c) {
  if (c < p.total) {
    const uint32_t fb = e4::smem_u32(&p.bfull[c & 1]);
    e4::mbar_expect_tx(fb, B_STAGE_BYTES);
    e4::bulk_copy(e4::smem_u32(p.sB + (c & 1) * B_STAGE_BYTES), p.wtc + (size_t)(c * p.wmul + p.woff) * (B_STAGE_BYTES / 4),
                  B_STAGE_BYTES, fb);
  }
}

// One K = 32 chunk of the current GEMM: a[] = this thread's 32 activations (row m = threadIdx.x).
// first: the chunk starts a new accumulation (in the 128 accumulator columns at acc_col).  All 128 threads call this in
// lockstep.
__device__ __forceinline__ void chunk_mma(Pipe& p, const float (&a)[32], bool first, uint32_t acc_col = 0) {
  const int m = threadIdx.x;
  // the previous chunk's MMAs have consumed the A buffers (and the stage the NEXT copy goes to)
  if (p.chunk > 0) wait_bar(p.bmma, (p.chunk - 1) & 1);
  if (m == 0) load_b(p, p.chunk + 1);
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    const float4 hi = make_float4(tf32_rna_finite(a[k]), tf32_rna_finite(a[k + 1]), tf32_rna_finite(a[k + 2]),
                                  tf32_rna_finite(a[k + 3]));
    const float4 lo = make_float4(tf32_rna_finite(a[k] - hi.x), tf32_rna_finite(a[k + 1] - hi.y),
                                  tf32_rna_finite(a[k + 2] - hi.z), tf32_rna_finite(a[k + 3] - hi.w));
    const uint32_t off = tc::umma_off(m, k, KC);
    *reinterpret_cast<float4*>(p.sAhi + off) = hi;
    *reinterpret_cast<float4*>(p.sAlo + off) = lo;
  }
  tc::fence_async_smem();          // generic-proxy writes of A -> visible to the tensor core's async proxy
  tc::fence_before_sync();         // orders this thread's earlier tcgen05.ld of the accumulator before the barrier
  __syncthreads();
  if (m == 0) {
    wait_bar(&p.bfull[p.chunk & 1], (p.chunk >> 1) & 1);
    tc::fence_after_sync();
    const uint32_t idesc = tc::make_idesc_tf32(128, 128);
    constexpr uint32_t SBO = KC * 32, LBO = 128;
    const uint32_t bh0 = e4::smem_u32(p.sB + (p.chunk & 1) * B_STAGE_BYTES), bl0 = bh0 + B_STAGE_BYTES / 2;
    const uint32_t ah0 = e4::smem_u32(p.sAhi), al0 = e4::smem_u32(p.sAlo);
#pragma unroll
    for (int ks = 0; ks < KC / 8; ++ks) {
      const uint32_t adv = ks * 2 * LBO;
      const uint64_t ah = tc::make_smem_desc(ah0 + adv, LBO, SBO), al = tc::make_smem_desc(al0 + adv, LBO, SBO);
      const uint64_t bh = tc::make_smem_desc(bh0 + adv, LBO, SBO), bl = tc::make_smem_desc(bl0 + adv, LBO, SBO);
      // hi.hi and the two cross terms accumulate in SEPARATE 128-column accumulators (summed in fp32 by read_acc): the
      // tensor core truncates on every accumulate, and 2^-11-sized terms added one by one into the large sum tripled the
      // number of truncations at the large sum's magnitude (closed-loop drift of the bench-shape parity test, r2)
      tc::mma_tf32(p.tmem + acc_col, ah, bh, idesc, !(first && ks == 0));
      tc::mma_tf32(p.tmem + acc_col + 128, al, bh, idesc, !(first && ks == 0));
      tc::mma_tf32(p.tmem + acc_col + 128, ah, bl, idesc, true);
    }
    tc::mma_commit(p.bmma);
  }
  ++p.chunk;
}

// wait for the last chunk's MMAs and read this thread's 128 accumulator columns (+ bias)
__device__ __forceinline__ void read_acc(Pipe& p, const float* __restrict__ bias, float (&v)[128], uint32_t acc_col = 0) {
  wait_bar(p.bmma, (p.chunk - 1) & 1);
  tc::fence_after_sync();
  const uint32_t lane_base = p.tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
#pragma unroll
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float t[32];
    tc::tmem_ld32(lane_base + acc_col + c0, t);
    {
      float u[32];
      tc::tmem_ld32(lane_base + acc_col + 128 + c0, u);
#pragma unroll
      for (int i = 0; i < 32; ++i) t[i] += u[i];
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b = *reinterpret_cast<const float4*>(bias + c0 + i);    // global or shared
      v[c0 + i] = t[i] + b.x; v[c0 + i + 1] = t[i + 1] + b.y; v[c0 + i + 2] = t[i + 2] + b.z; v[c0 + i + 3] = t[i + 3] + b.w;
    }
  }
  // the bmma wait above is repeated by the next chunk_mma (same parity: passes immediately)
}

__device__ __forceinline__ void ln_relu(float (&v)[128], const float* __restrict__ g, const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 128; ++i) s += v[i];
  const float mean = s * (1.0f / 128.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 128; ++i) {
    const float d = v[i] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = 1.0f / sqrtf(q * (1.0f / 128.0f) + LN_EPS);
#pragma unroll
  for (int i = 0; i < 128; i += 4) {
    const float4 gg = *reinterpret_cast<const float4*>(g + i), bb = *reinterpret_cast<const float4*>(b + i);
    v[i] = fmaxf((v[i] - mean) * rstd * gg.x + bb.x, 0.f);
    v[i + 1] = fmaxf((v[i + 1] - mean) * rstd * gg.y + bb.y, 0.f);
    v[i + 2] = fmaxf((v[i + 2] - mean) * rstd * gg.z + bb.z, 0.f);
    v[i + 3] = fmaxf((v[i + 3] - mean) * rstd * gg.w + bb.w, 0.f);
  }
}

template <int KTOT>
__device__ __forceinline__ void gemm_rows(Pipe& p, const float (&v)[KTOT]) {
#pragma unroll
  for (int c = 0; c < KTOT / 32; ++c) {
    float a[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = v[c * 32 + i];
    chunk_mma(p, a, c == 0);
  }
}

}  // namespace pntc

// X: [n_all][P][IN], mask: uint8 [n_all][P][mask_inner] or nullptr (validity = no NaN feature), rows: [n_poly] polyline
// indices, W: fp32 block (pw::, biases and LayerNorm vectors), Wtc: packed tensor-core chunks, Out: [n_poly][128]
template <int IN, int NPRE, int P>
__global__ void __launch_bounds__(128, 2) pointnet_tc_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask,
                                                             int mask_inner, const int* __restrict__ rows, int n_poly,
                                                             const float* __restrict__ W, const float* __restrict__ Wtc,
                                                             float* __restrict__ Out) {
  using C = pntc::Cfg<P>;
  constexpr int G = C::G;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (e4::smem_u32(smem_raw) & 1023u)) & 1023u);
  pntc::Pipe p;
  p.sAhi = base;
  p.sAlo = base + pntc::A_BYTES;
  p.sB = base + 2 * pntc::A_BYTES;
  float* sPool = reinterpret_cast<float*>(p.sB + 2 * pntc::B_STAGE_BYTES);   // [GP][128]
  int* sValid = reinterpret_cast<int*>(sPool + C::GP * 128);                 // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sValid + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* sVec = reinterpret_cast<float*>(bars + 8);                         // [N_VEC][128]
  p.bfull = bars;
  p.bmma = bars + 2;
  p.wtc = Wtc;
  p.wmul = 1;
  p.woff = 0;
  p.chunk = 0;
  p.total = pntc::n_chunks<NPRE>();
  float* sScr = reinterpret_cast<float*>(base);       // pooling scratch [128][33] floats = 16.9 KB over the (idle) A buffers

  const int m = threadIdx.x, warp = m >> 5;
  const int poly0 = blockIdx.x * G;
  const int g = m / P, pt = m % P;
  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  if (m == 32) {
    tc::mbar_init(&p.bfull[0], 1);
    tc::mbar_init(&p.bfull[1], 1);
    tc::mbar_init(p.bmma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // row validity and inputs
  int valid = 0;
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = 0.f;
  if (g < G && poly0 + g < n_poly) {
    const size_t prow = (size_t)rows[poly0 + g] * P + pt;
    const float* xp = X + prow * IN;
    valid = 1;
    if (mask != nullptr) {
      const uint8_t* mk = mask + prow * mask_inner;
      for (int i = 0; i < mask_inner; ++i) valid &= (mk[i] != 0);
    } else {
      for (int i = 0; i < IN; ++i) valid &= !isnan(xp[i]);
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < IN; ++i) x[i] = xp[i];
    }
  }
  sValid[m] = valid;
  {   // per-column vectors -> shared memory (every thread reads all of them, many times)
    constexpr int src[pntc::N_VEC] = {pw::PRE0_B, pw::PRE0_G, pw::PRE0_BB, pw::PRE1_B, pw::PRE1_G, pw::PRE1_BB, pw::PRE2_B,
                                      pw::MLP0_B, pw::MLP0_G, pw::MLP0_BB, pw::MLP1_B, pw::OUT0_B, pw::OUT1_B};
#pragma unroll
    for (int k = 0; k < pntc::N_VEC; ++k) sVec[k * 128 + m] = __ldg(W + src[k] + m);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  p.tmem = *tmem_slot;
  if (m == 0) pntc::load_b(p, 0);

  float v[128];
  // max over the P points of each polyline (zeros of masked points included) -> sPool[g][128]
  auto pool = [&]() {
    // the A buffers are idle here: read_acc waited for the MMAs that read them
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 32) {      // fully unrolled: v[] must stay in registers
#pragma unroll
      for (int i = 0; i < 32; ++i) sScr[m * pntc::POOL_LD + i] = v[c0 + i];
      __syncthreads();
      for (int o = m; o < G * 32; o += 128) {
        const int gg = o >> 5, c = o & 31;
        float mx = sScr[(gg * P) * pntc::POOL_LD + c];
        for (int q = 1; q < P; ++q) mx = fmaxf(mx, sScr[(gg * P + q) * pntc::POOL_LD + c]);
        sPool[gg * 128 + c0 + c] = mx;
      }
      __syncthreads();
    }
  };
  auto relu_mask = [&]() {
#pragma unroll
    for (int i = 0; i < 128; ++i) v[i] = valid ? fmaxf(v[i], 0.f) : 0.f;
  };

  // ---- the chain as a runtime loop over six stages (each: epilogue of the previous GEMM, then a K = 128 GEMM): one copy of
  // the staging / MMA-issue / accumulator-read / LayerNorm / pooling code instead of one per GEMM.  Fully unrolled, the
  // kernel was 11-16 k instructions (180-260 KB) that every warp walked through exactly once: the top stall was
  // instruction fetch (ncu: 2.4-5.1 "no instruction" cycles per issued instruction).
  //   stage 0, 1 (NPRE = 3 only): LN+ReLU -> pre_mlps.1 / pre_mlps.2
  //   stage 2: ReLU+mask, max-pool -> mlps.0 on [point | pooled] (two K = 128 halves into one accumulator)
  //   stage 3: LN+ReLU -> mlps.1
  //   stage 4: ReLU+mask, max-pool, pooled rows to threads 0..G-1 -> out_mlps.0
  //   stage 5: ReLU -> out_mlps.1
  pntc::chunk_mma(p, x, true);
  pntc::read_acc(p, sVec, v);
#pragma unroll 1
  for (int st = (NPRE == 3 ? 0 : 2); st < 6; ++st) {
    if (st == 0 || st == 1 || st == 3) {
      const int go = st == 0 ? 1 : st == 1 ? 4 : 8;            // sVec rows: gamma, then beta
      pntc::ln_relu(v, sVec + go * 128, sVec + (go + 1) * 128);
    } else if (st == 5) {
#pragma unroll
      for (int i = 0; i < 128; ++i) v[i] = m < G ? fmaxf(v[i], 0.f) : 0.f;
    } else {
      relu_mask();
      if (st == 4) __syncthreads();                            // every thread has read its pooled row of the first pooling
      pool();
      if (st == 4) {                                           // out_mlps run on the G pooled rows (rows >= G are zero)
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m < G) t = *reinterpret_cast<const float4*>(sPool + m * 128 + i);
          v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
        }
      }
    }
#pragma unroll 1
    for (int half = 0; half < (st == 2 ? 2 : 1); ++half) {
      if (half == 1) {                                         // the pooled half of mlps.0's input
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g < G) t = *reinterpret_cast<const float4*>(sPool + g * 128 + i);
          v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float a[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = v[c * 32 + i];
        pntc::chunk_mma(p, a, half == 0 && c == 0);
      }
    }
    const int bo = st == 0 ? 3 : st == 1 ? 6 : st == 2 ? 7 : st == 3 ? 10 : st == 4 ? 11 : 12;   // sVec row of the bias
    pntc::read_acc(p, sVec + bo * 128, v);
  }
  if (m < G && poly0 + m < n_poly) {
    int any = 0;
    for (int q = 0; q < P; ++q) any |= sValid[m * P + q];
    float4* o = reinterpret_cast<float4*>(Out + (size_t)(poly0 + m) * D);
#pragma unroll
    for (int i = 0; i < 128; i += 4)
      o[i >> 2] = any ? make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(p.tmem, 256);
}

}  // namespace prosim
