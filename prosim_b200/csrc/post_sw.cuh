// Node side of the AttentionLayer on tcgen05 / TMEM, "swapped" orientation: the chip-filling variant of tc_post.cuh.
//
// tc_post.cuh puts 128 destination rows on the 128 TMEM lanes of one CTA, so a 4096-row launch (32 scenes x 128 agents,
// the per-GPU shard of BASELINE configs[4]) occupies 32 of 148 SMs and its per-CTA latency chain (80 us) is the launch
// time.  Here the GEMMs are transposed: D^T[feature][row] = W[feature][k] . X[row][k]^T -- the WEIGHTS are the M-side
// operand (128 output features = 128 TMEM lanes, pre-split tf32 hi / lo chunks streamed from L2 by cp.async.bulk, the
// same packed chunks tc_post.cuh consumes as its N side) and the ACTIVATIONS of NR = 32 destination rows are the N-side
// operand (K-major in shared memory, written by the epilogue threads).  A CTA owns 32 rows: 128 CTAs per 4096 rows,
// the tensor work per CTA drops 4x (N = 32: 16 cycles per MMA), the per-thread epilogue work 4x (16 values per phase).
//
// Same math as tc_post.cuh / attn_post2_kernel (reference prosim/models/layers/attention_layer.py:102-118, :38-43 and
// the next layer's destination-side projections :56-66):
//   agg = AggV + Rbar_h . Wvr'_h           (normal orientation: rows on lanes 0..31 of an M = 128 MMA, N = 16 per head)
//   g = sigmoid(Wga agg + Gx) ; u = agg + g (S - agg) ; x1 = x + LN(Wo u + bo)
//   y = W2 relu(W1 LN(x1) + b1) + b2 ; out = x1 + LN(y) ; next layer: s | gx | q = {Ws, Wgx, Wq} LN_dst'(out), Qhat_h
// every product as three tf32 MMAs (a_lo w_hi + a_hi w_lo + a_hi w_hi, fp32 accumulation in TMEM).
//
// Thread roles (320 threads, 1 CTA / SM):
//   warps 0-7  epilogue.  Two index maps over the CTA's [32 rows x 128 features] tile, 16 values per thread:
//                T-map (TMEM native): thread = feature f = 32 (warp % 4) + lane, rows 16 (warp / 4) .. + 15.  Biases are
//                  per-thread scalars; global rows [r][f] are read / written as coalesced 128-byte lines; the next
//                  operand is written with conflict-free 4-byte stores (operand core matrices are 144 bytes apart).
//                R-map (row major): thread = (row 4 warp + lane / 8, feature groups s + 8 i of 4 floats, s = lane % 8):
//                  LayerNorm statistics are three xor-shuffles inside 8 lanes.  T-map -> R-map goes through a padded
//                  [32][132] fp32 scratch tile and one named barrier.
//   warp 8     weight producer (one thread, 4 x 32 KB ring on full / empty mbarriers)
//   warp 9     MMA issuer (one thread)
// TMEM: twelve [128 lanes x 32 columns] accumulator slots + the [32 lanes x 128 columns] agg accumulator; every slot is
// written by one GEMM and read by one epilogue, so there are no accumulator-free barriers; every accumulator has its own
// single-use "done" mbarrier.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "edge4.cuh"     // e4:: mbarrier / bulk-copy helpers
#include "tc_gemm.cuh"   // tc:: tcgen05 helpers
#include "tc_post.cuh"   // tcp:: V_* vector offsets, tf32 split, bounded mbarrier wait
#include "weights_layout.h"

namespace prosim {
namespace psw {

constexpr int NR = 32;                     // destination rows per CTA
constexpr int NST = 4;                     // weight ring stages
constexpr int STAGE_BYTES = 32768;
constexpr int LBO_A = 144;                 // activation operands: K-adjacent core matrices 144 B apart (bank spread)
constexpr int SBO_A = 32 * LBO_A;          // 8-row groups of a [32 x 128] operand
constexpr int OPND_HALF = 4 * SBO_A;       // hi (or lo) part
constexpr int OPND_BYTES = 2 * OPND_HALF;
constexpr int SCR_LD = 132;                // scratch row pitch (floats)
constexpr int EPI_THREADS = 256;
constexpr int THREADS = EPI_THREADS + 64;
constexpr uint32_t AGG_COL = 384;          // agg accumulator [lanes 0..31][128]
constexpr int A_AGG = 0, A_GATE = 1, A_OUT = 2, A_UP = 3, A_DOWN = 7, A_S = 8, A_GX = 9, A_Q = 10, A_QH = 11, NACC = 19;
// accumulator slots (x 32 columns): gate 0, out 1, up_t 2..5, down 6, s 7, gx 8, q 9, Qhat_h: 0..5, 10, 11
__host__ __device__ constexpr uint32_t slot_col(int s) { return 32u * (uint32_t)s; }
__host__ __device__ constexpr int qh_slot(int h) { return h < 6 ? h : 4 + h; }

struct Smem {
  uint8_t actA[OPND_BYTES];                // activation operand (agg / u / LN(x1) / LN_dst(out)); Rbar head buffer 0
  uint8_t actH[OPND_BYTES];                // FFN hidden operand, q operand; Rbar head buffer 1
  uint8_t ring[NST][STAGE_BYTES];          // (the agg MMAs read 96 unused operand rows past the Rbar buffers: into here)
  float scratch[NR * SCR_LD];
  float vec[tcp::V_SIZE];
  uint64_t full[NST], empty[NST];
  uint64_t rb_ready[2], rb_free[2], opnd_ready, h_ready, h_free, acc_done[NACC];
  uint32_t tmem_base;
};
constexpr size_t SMEM_BYTES = sizeof(Smem) + 128;

struct Args {
  const float *x, *rbar, *aggv, *s, *gx;   // [n][128], [n][8 zd], [n][128] ...
  float* out;                              // [n][128]
  float *q_n, *qhat_n, *s_n, *gx_n;        // next layer's destination-side projections (has_next)
  const float *W, *Wn;                     // packed layer weights (aw::), next layer's (nullptr = last layer)
  int n;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { tcp::mbar_arrive(e4::smem_u32(bar)); }

template <int ZD>
__global__ void __launch_bounds__(THREADS, 1) attn_post_sw_kernel(const __grid_constant__ Args a) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((128u - (e4::smem_u32(smem_raw) & 127u)) & 127u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * NR;
  const float* __restrict__ W = a.W;
  const float* __restrict__ Wn = a.Wn;
  const bool has_next = Wn != nullptr;
  constexpr int SBO_R = (ZD / 4) * LBO_A;              // 8-row groups of a [rows x ZD] Rbar head operand
  constexpr int RB_HALF = 4 * SBO_R;
  constexpr int NG_R = ZD / 32;                        // 16-byte groups per thread per Rbar head row

  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 512);
  if (tid == 32) {
    for (int i = 0; i < NST; ++i) {
      e4::mbar_init(e4::smem_u32(&sm.full[i]), 1);
      e4::mbar_init(e4::smem_u32(&sm.empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      e4::mbar_init(e4::smem_u32(&sm.rb_ready[i]), EPI_THREADS);
      e4::mbar_init(e4::smem_u32(&sm.rb_free[i]), 1);
    }
    e4::mbar_init(e4::smem_u32(&sm.opnd_ready), EPI_THREADS);
    e4::mbar_init(e4::smem_u32(&sm.h_ready), EPI_THREADS);
    e4::mbar_init(e4::smem_u32(&sm.h_free), 1);
    for (int i = 0; i < NACC; ++i) e4::mbar_init(e4::smem_u32(&sm.acc_done[i]), 1);
  }
  // per-column vectors -> shared memory (same table as tc_post.cuh)
  for (int i = tid; i < tcp::V_SIZE; i += THREADS) {
    float v = 0.f;
    if (i < tcp::V_B1) v = W[aw::BO + i];
    else if (i < tcp::V_B2) v = W[aw::B1 + (i - tcp::V_B1)];
    else if (i < tcp::V_LNDST_G) v = W[aw::B2 + (i - tcp::V_B2)];
    else if (has_next) {
      if (i < tcp::V_BQ) v = Wn[aw::LN_DST_G + (i - tcp::V_LNDST_G)];
      else if (i < tcp::V_BS) v = Wn[aw::BQ + (i - tcp::V_BQ)];
      else if (i < tcp::V_BG) v = Wn[aw::BS + (i - tcp::V_BS)];
      else v = Wn[aw::BG + (i - tcp::V_BG)];
    }
    sm.vec[i] = v;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  const int n_chunks = has_next ? 68 : 48;

  if (warp == 8) {
    // ================================================================== weight producer
    if (lane == 0) {
      constexpr int VR_FLOATS = 16 * ZD * 2;
      for (int i = 0; i < n_chunks; ++i) {
        const float* src;
        uint32_t bytes;
        if (i < 8) {
          src = W + (ZD == 96 ? aw::TC_VR96 : aw::TC_VR128) + i * VR_FLOATS;
          bytes = VR_FLOATS * 4;
        } else if (i < 16) {
          src = W + aw::TC_GA + (i - 8) * 8192;      // Wga (4), Wo (4): contiguous
          bytes = 32768;
        } else if (i < 48) {
          src = W + aw::TC_FF2 + (i - 16) * 8192;    // up_0, up_1, down_0, up_2, down_1, up_3, down_2, down_3 (4 k-chunks each)
          bytes = 32768;
        } else if (i < 56) {
          src = Wn + aw::TC_S + (i - 48) * 8192;     // next layer's Ws, Wgx (contiguous) ...
          bytes = 32768;
        } else if (i < 60) {
          src = Wn + aw::TC_Q + (i - 56) * 8192;     // ... then Wq
          bytes = 32768;
        } else {
          src = Wn + aw::TC_KRG + (i - 60) * 4096;
          bytes = 16384;
        }
        const int s = i % NST;
        if (i >= NST) tcp::mbar_wait(&sm.empty[s], ((i / NST) - 1) & 1);
        const uint32_t fb = e4::smem_u32(&sm.full[s]);
        e4::mbar_expect_tx(fb, bytes);
        e4::bulk_copy(e4::smem_u32(sm.ring[s]), src, bytes, fb);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      int ci = 0;
      uint32_t ph_rb0 = 0, ph_rb1 = 0, ph_op = 0, ph_h = 0;
      auto wait_bar = [&](uint64_t* bar, uint32_t& ph) {
        tcp::mbar_wait(bar, ph);
        ph ^= 1;
        tc::fence_after_sync();
      };
      auto next_chunk = [&]() -> uint32_t {
        const int s = ci % NST;
        tcp::mbar_wait(&sm.full[s], (ci / NST) & 1);
        tc::fence_after_sync();
        return e4::smem_u32(sm.ring[s]);
      };
      auto release_chunk = [&]() {
        tc::mma_commit(&sm.empty[ci % NST]);
        ++ci;
      };
      // D^T[128 features][32 rows] (+)= Wchunk[128 x KC] . X[32 rows][k0 .. k0 + KC)^T
      auto gemm_sw = [&](auto kc_tag, uint32_t d_col, uint32_t opnd, int k0, bool accumulate) {
        constexpr int KC_ = decltype(kc_tag)::value;
        const uint32_t wb = next_chunk();
        constexpr uint32_t idesc = tc::make_idesc_tf32(128, NR);
        const uint64_t wh0 = tc::make_smem_desc(wb, 128, KC_ * 32);
        const uint64_t wl0 = tc::make_smem_desc(wb + 128 * KC_ * 4, 128, KC_ * 32);
        const uint64_t xh0 = tc::make_smem_desc(opnd + (k0 >> 2) * LBO_A, LBO_A, SBO_A);
        const uint64_t xl0 = tc::make_smem_desc(opnd + OPND_HALF + (k0 >> 2) * LBO_A, LBO_A, SBO_A);
#pragma unroll
        for (int ks = 0; ks < KC_ / 8; ++ks) {
          const uint64_t wh = wh0 + (uint64_t)(ks * 16), wl = wl0 + (uint64_t)(ks * 16);
          const uint64_t xh = xh0 + (uint64_t)(ks * (2 * LBO_A / 16)), xl = xl0 + (uint64_t)(ks * (2 * LBO_A / 16));
          tc::mma_tf32(tmem + d_col, wh, xl, idesc, accumulate || ks > 0);
          tc::mma_tf32(tmem + d_col, wl, xh, idesc, true);
          tc::mma_tf32(tmem + d_col, wh, xh, idesc, true);
        }
        release_chunk();
      };
      using I16 = std::integral_constant<int, 16>;
      using I32 = std::integral_constant<int, 32>;
      auto gemm128 = [&](uint32_t d_col, uint32_t opnd) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) gemm_sw(I32{}, d_col, opnd, 32 * c, c > 0);
      };
      const uint32_t actA = e4::smem_u32(sm.actA), actH = e4::smem_u32(sm.actH);
      // 1. agg (normal orientation): D[row lanes][16 h ..] = Rbar_h[rows x ZD] . Wvr'_h[16 x ZD]^T
#pragma unroll 1
      for (int h = 0; h < H; ++h) {
        if (h & 1) wait_bar(&sm.rb_ready[1], ph_rb1);
        else wait_bar(&sm.rb_ready[0], ph_rb0);
        const uint32_t rb = (h & 1) ? actH : actA;
        const uint32_t wb = next_chunk();
        constexpr uint32_t idesc = tc::make_idesc_tf32(128, 16);
        const uint64_t ah0 = tc::make_smem_desc(rb, LBO_A, SBO_R);
        const uint64_t al0 = tc::make_smem_desc(rb + RB_HALF, LBO_A, SBO_R);
        const uint64_t bh0 = tc::make_smem_desc(wb, 128, ZD * 32);
        const uint64_t bl0 = tc::make_smem_desc(wb + 16 * ZD * 4, 128, ZD * 32);
        const uint32_t d = tmem + AGG_COL + 16 * h;
#pragma unroll
        for (int ks = 0; ks < ZD / 8; ++ks) {
          const uint64_t ah = ah0 + (uint64_t)(ks * (2 * LBO_A / 16)), al = al0 + (uint64_t)(ks * (2 * LBO_A / 16));
          const uint64_t bh = bh0 + (uint64_t)(ks * 16), bl = bl0 + (uint64_t)(ks * 16);
          tc::mma_tf32(d, al, bh, idesc, ks > 0);
          tc::mma_tf32(d, ah, bl, idesc, true);
          tc::mma_tf32(d, ah, bh, idesc, true);
        }
        release_chunk();
        tc::mma_commit(&sm.rb_free[h & 1]);
      }
      tc::mma_commit(&sm.acc_done[A_AGG]);
      // 2. gate, 3. out projection
      wait_bar(&sm.opnd_ready, ph_op);
      gemm128(slot_col(0), actA);
      tc::mma_commit(&sm.acc_done[A_GATE]);
      wait_bar(&sm.opnd_ready, ph_op);
      gemm128(slot_col(1), actA);
      tc::mma_commit(&sm.acc_done[A_OUT]);
      // 4. FFN: hidden tile t = 128 features; chunks arrive as up_0, up_1, down_0, up_2, down_1, up_3, down_2, down_3
      wait_bar(&sm.opnd_ready, ph_op);
      gemm128(slot_col(2), actA);
      tc::mma_commit(&sm.acc_done[A_UP + 0]);
      gemm128(slot_col(3), actA);
      tc::mma_commit(&sm.acc_done[A_UP + 1]);
#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        wait_bar(&sm.h_ready, ph_h);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) gemm_sw(I32{}, slot_col(6), actH, 32 * c, t > 0 || c > 0);
        if (t < 3) tc::mma_commit(&sm.h_free);
        else tc::mma_commit(&sm.acc_done[A_DOWN]);
        if (t + 2 < 4) {
          gemm128(slot_col(2 + t + 2), actA);
          tc::mma_commit(&sm.acc_done[A_UP + t + 2]);
        }
      }
      if (has_next) {
        // 5. next layer: s, gx, q from LN_dst'(out); then Qhat_h = Wkr'_h[128 x 16] . q_h^T
        wait_bar(&sm.opnd_ready, ph_op);
        gemm128(slot_col(7), actA);
        tc::mma_commit(&sm.acc_done[A_S]);
        gemm128(slot_col(8), actA);
        tc::mma_commit(&sm.acc_done[A_GX]);
        gemm128(slot_col(9), actA);
        tc::mma_commit(&sm.acc_done[A_Q]);
        wait_bar(&sm.opnd_ready, ph_op);
#pragma unroll 1
        for (int h = 0; h < H; ++h) {
          gemm_sw(I16{}, slot_col(qh_slot(h)), actH, 16 * h, false);
          tc::mma_commit(&sm.acc_done[A_QH + h]);
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================== epilogue warps
    const int q4 = warp & 3, ch = warp >> 2;
    const int f = q4 * 32 + lane;                          // T-map: feature, rows 16 ch .. 16 ch + 15
    const uint32_t lb = tmem + ((uint32_t)(q4 * 32) << 16);
    const int rr = 4 * warp + (lane >> 3), sg = lane & 7;  // R-map: row, feature groups sg + 8 i
    const bool rr_ok = row0 + rr < a.n;
    const float* vec = sm.vec;
    float* scr = sm.scratch;
    uint32_t ph_rbf0 = 0, ph_rbf1 = 0, ph_hfree = 0;
    auto wait_acc = [&](int i) {
      tcp::mbar_wait(&sm.acc_done[i], 0);
      tc::fence_after_sync();
    };
    auto publish = [&](uint64_t* bar) {
      e4::fence_proxy_async();
      tc::fence_before_sync();
      mbar_arrive(bar);
    };
    // T-map: 16 rows of this thread's feature -> hi / lo activation operand at `opnd`
    auto t_store_opnd = [&](uint8_t* opnd, const float (&v)[16]) {
      uint8_t* p = opnd + (f >> 2) * LBO_A + (f & 3) * 4 + (2 * ch) * SBO_A;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float hi = tcp::tf32_rna_finite(v[i]);
        const float lo = tcp::tf32_rna_finite(v[i] - hi);
        uint8_t* pe = p + (i >> 3) * SBO_A + (i & 7) * 16;
        *reinterpret_cast<float*>(pe) = hi;
        *reinterpret_cast<float*>(pe + OPND_HALF) = lo;
      }
    };
    // R-map: this thread's 4 groups of 4 features of row rr -> hi / lo activation operand
    auto r_store_opnd = [&](uint8_t* opnd, const float (&v)[16]) {
      uint8_t* p = opnd + (rr >> 3) * SBO_A + (rr & 7) * 16 + sg * LBO_A;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 hi, lo;
        hi.x = tcp::tf32_rna_finite(v[4 * i]);     lo.x = tcp::tf32_rna_finite(v[4 * i] - hi.x);
        hi.y = tcp::tf32_rna_finite(v[4 * i + 1]); lo.y = tcp::tf32_rna_finite(v[4 * i + 1] - hi.y);
        hi.z = tcp::tf32_rna_finite(v[4 * i + 2]); lo.z = tcp::tf32_rna_finite(v[4 * i + 2] - hi.z);
        hi.w = tcp::tf32_rna_finite(v[4 * i + 3]); lo.w = tcp::tf32_rna_finite(v[4 * i + 3] - hi.w);
        *reinterpret_cast<float4*>(p + 8 * i * LBO_A) = hi;
        *reinterpret_cast<float4*>(p + 8 * i * LBO_A + OPND_HALF) = lo;
      }
    };
    auto r_load_scr = [&](float (&v)[16]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(scr + rr * SCR_LD + 4 * (sg + 8 * i));
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    };
    auto r_load_vec = [&](const float* p, float (&v)[16]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * (sg + 8 * i));
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    };
    auto r_load_glb = [&](const float* base, float (&v)[16]) {   // row rr of a [n][128] global matrix
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr_ok) t = __ldg(reinterpret_cast<const float4*>(base + (size_t)(row0 + rr) * D) + sg + 8 * i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    };
    auto r_store_glb = [&](float* base, const float (&v)[16]) {
      if (rr_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *(reinterpret_cast<float4*>(base + (size_t)(row0 + rr) * D) + sg + 8 * i) =
              make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    };
    auto row_sum8 = [&](float p) -> float {                      // over the 8 lanes that share a row
      p += __shfl_xor_sync(0xffffffffu, p, 1);
      p += __shfl_xor_sync(0xffffffffu, p, 2);
      p += __shfl_xor_sync(0xffffffffu, p, 4);
      return p;
    };
    // two-pass LayerNorm of the row (R-map), affine from vec + g_off / b_off
    auto r_layernorm = [&](float (&v)[16], int g_off, int b_off) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) s += v[i];
      const float mean = row_sum8(s) * (1.0f / 128.0f);
      float qv = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float d = v[i] - mean;
        qv = fmaf(d, d, qv);
      }
      const float rstd = 1.0f / sqrtf(row_sum8(qv) * (1.0f / 128.0f) + LN_EPS);
      float g[16], b[16];
      r_load_vec(vec + g_off, g);
      r_load_vec(vec + b_off, b);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (v[i] - mean) * rstd * g[i] + b[i];
    };
    // T-map global access: element [row0 + 16 ch + i][f]
    auto t_load_glb = [&](const float* base, float (&v)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int r = row0 + 16 * ch + i;
        v[i] = r < a.n ? __ldg(base + (size_t)r * D + f) : 0.f;
      }
    };
    auto t_store_glb = [&](float* base, size_t ld, int col, const float (&v)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int r = row0 + 16 * ch + i;
        if (r < a.n) base[(size_t)r * ld + col + f] = v[i];
      }
    };
    auto t_store_scr = [&](const float (&v)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) scr[(16 * ch + i) * SCR_LD + f] = v[i];
    };

    // ---- 0. Rbar_h -> A-side operand tiles (hi | lo), heads alternate between the two operand buffers
    {
      const float* rb_row = a.rbar + (size_t)(row0 + rr) * (H * ZD);
      float4 cur[NG_R], nxt[NG_R];
      auto load_head = [&](int h, float4 (&dst)[NG_R]) {
#pragma unroll
        for (int i = 0; i < NG_R; ++i) {
          dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rr_ok) dst[i] = __ldg(reinterpret_cast<const float4*>(rb_row + h * ZD) + sg + 8 * i);
        }
      };
      load_head(0, cur);
#pragma unroll 1
      for (int h = 0; h < H; ++h) {
        if (h + 1 < H) load_head(h + 1, nxt);
        if (h >= 2) {
          if (h & 1) { tcp::mbar_wait(&sm.rb_free[1], ph_rbf1); ph_rbf1 ^= 1; }
          else { tcp::mbar_wait(&sm.rb_free[0], ph_rbf0); ph_rbf0 ^= 1; }
        }
        uint8_t* p = ((h & 1) ? sm.actH : sm.actA) + (rr >> 3) * SBO_R + (rr & 7) * 16 + sg * LBO_A;
#pragma unroll
        for (int i = 0; i < NG_R; ++i) {
          float4 hi, lo;
          hi.x = tcp::tf32_rna_finite(cur[i].x); lo.x = tcp::tf32_rna_finite(cur[i].x - hi.x);
          hi.y = tcp::tf32_rna_finite(cur[i].y); lo.y = tcp::tf32_rna_finite(cur[i].y - hi.y);
          hi.z = tcp::tf32_rna_finite(cur[i].z); lo.z = tcp::tf32_rna_finite(cur[i].z - hi.z);
          hi.w = tcp::tf32_rna_finite(cur[i].w); lo.w = tcp::tf32_rna_finite(cur[i].w - hi.w);
          *reinterpret_cast<float4*>(p + 8 * i * LBO_A) = hi;
          *reinterpret_cast<float4*>(p + 8 * i * LBO_A + RB_HALF) = lo;
        }
        publish(&sm.rb_ready[h & 1]);
#pragma unroll
        for (int i = 0; i < NG_R; ++i) cur[i] = nxt[i];
      }
    }
    float v[16], t[16];
    // ---- 1. agg = AGG accumulator (rows on lanes 0..31) + AggV -> scratch (fp32, for the gate) and operand A
    r_load_glb(a.aggv, t);
    wait_acc(A_AGG);
    if (q4 == 0) {   // warps 0 and 4 own TMEM lanes 0..31: thread = row, 64 columns each
      float w[32];
#pragma unroll 1
      for (int c0 = 64 * ch; c0 < 64 * ch + 64; c0 += 32) {
        tc::tmem_ld32(tmem + AGG_COL + c0, w);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(scr + lane * SCR_LD + c0 + 4 * i) = make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
      }
    }
    epi_barrier();
    r_load_scr(v);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += t[i];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(scr + rr * SCR_LD + 4 * (sg + 8 * i)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    r_store_opnd(sm.actA, v);
    publish(&sm.opnd_ready);
    // ---- 2. gate: g = sigmoid(acc + Gx) ; u = agg + g (S - agg) -> operand A        (T-map)
    {
      float gxv[16], sv[16];
      t_load_glb(a.gx, gxv);
      t_load_glb(a.s, sv);
      wait_acc(A_GATE);
      tmem_ld16(lb + slot_col(0) + 16 * ch, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float ag = scr[(16 * ch + i) * SCR_LD + f];
        const float g = 1.0f / (1.0f + expf(-(v[i] + gxv[i])));
        v[i] = ag + g * (sv[i] - ag);
      }
      t_store_opnd(sm.actA, v);
      publish(&sm.opnd_ready);
    }
    // ---- 3. o = acc + bo ; x1 = x + LN_post(o) (kept in registers, R-map) ; LN_ffpre(x1) -> operand A
    float x1[16];
    {
      r_load_glb(a.x, x1);
      const float bo = vec[tcp::V_BO + f];
      wait_acc(A_OUT);
      tmem_ld16(lb + slot_col(1) + 16 * ch, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += bo;
      t_store_scr(v);
      epi_barrier();
      r_load_scr(v);
      r_layernorm(v, tcp::V_LNPOST_G, tcp::V_LNPOST_B);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        x1[i] += v[i];
        v[i] = x1[i];
      }
      r_layernorm(v, tcp::V_LNFFPRE_G, tcp::V_LNFFPRE_B);
      r_store_opnd(sm.actA, v);
      publish(&sm.opnd_ready);
    }
    // ---- 4. FFN hidden tiles: h_t = relu(acc + b1) -> operand H                     (T-map)
#pragma unroll 1
    for (int tt = 0; tt < 4; ++tt) {
      const float b1 = vec[tcp::V_B1 + 128 * tt + f];
      wait_acc(A_UP + tt);
      tmem_ld16(lb + slot_col(2 + tt) + 16 * ch, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + b1, 0.f);
      if (tt > 0) {
        tcp::mbar_wait(&sm.h_free, ph_hfree);   // down_{t-1} has consumed the hidden operand
        ph_hfree ^= 1;
      }
      t_store_opnd(sm.actH, v);
      publish(&sm.h_ready);
    }
    // ---- y = acc + b2 ; out = x1 + LN_ffpost(y) -> global ; LN_dst'(out) -> operand A
    {
      const float b2 = vec[tcp::V_B2 + f];
      wait_acc(A_DOWN);
      tmem_ld16(lb + slot_col(6) + 16 * ch, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += b2;
      t_store_scr(v);
      epi_barrier();
      r_load_scr(v);
      r_layernorm(v, tcp::V_LNFFPOST_G, tcp::V_LNFFPOST_B);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += x1[i];
      r_store_glb(a.out, v);
      if (has_next) {
        r_layernorm(v, tcp::V_LNDST_G, tcp::V_LNDST_B);
        r_store_opnd(sm.actA, v);
        publish(&sm.opnd_ready);
      }
    }
    if (has_next) {
      // ---- 5. s, gx, q (+ bias) -> global (coalesced lines, T-map); q also -> operand H; then the eight Qhat_h
      {
        const float bs = vec[tcp::V_BS + f];
        wait_acc(A_S);
        tmem_ld16(lb + slot_col(7) + 16 * ch, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += bs;
        t_store_glb(a.s_n, D, 0, v);
      }
      {
        const float bg = vec[tcp::V_BG + f];
        wait_acc(A_GX);
        tmem_ld16(lb + slot_col(8) + 16 * ch, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += bg;
        t_store_glb(a.gx_n, D, 0, v);
      }
      {
        const float bq = vec[tcp::V_BQ + f];
        wait_acc(A_Q);
        tmem_ld16(lb + slot_col(9) + 16 * ch, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += bq;
        t_store_opnd(sm.actH, v);             // every MMA that read the hidden operand completed before acc_done[DOWN]
        publish(&sm.opnd_ready);
        t_store_glb(a.q_n, D, 0, v);
      }
#pragma unroll 1
      for (int h = 0; h < H; ++h) {
        wait_acc(A_QH + h);
        tmem_ld16(lb + slot_col(qh_slot(h)) + 16 * ch, v);
        t_store_glb(a.qhat_n, (size_t)H * D, h * D, v);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace psw
}  // namespace prosim
