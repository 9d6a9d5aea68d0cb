// Node side of the AttentionLayer on tcgen05 / TMEM, "swapped" orientation: the chip-filling variant of tc_post.cuh.
//
// tc_post.cuh puts 128 destination rows on the 128 TMEM lanes of one CTA, so a 4096-row launch (32 scenes x 128 agents,
// the per-GPU shard of BASELINE configs[4]) occupies 32 of 148 SMs and its per-CTA latency chain (80 us) is the launch
// time.  Here the GEMMs are transposed: D^T[feature][row] = W[feature][k] . X[row][k]^T -- the WEIGHTS are the M-side
// operand (128 output features = 128 TMEM lanes, pre-split tf32 hi / lo chunks streamed from L2 by cp.async.bulk, the
// same packed chunks tc_post.cuh consumes as its N side) and the ACTIVATIONS of NR = 32 destination rows are the N-side
// operand (K-major, 128B-swizzled, written by the epilogue threads).  A CTA owns 32 rows: 128 CTAs per 4096 rows.
//
// Same math as tc_post.cuh / attn_post2_kernel (reference prosim/models/layers/attention_layer.py:102-118, :38-43 and
// the next layer's destination-side projections :56-66):
//   agg = AggV + Rbar_h . Wvr'_h           (fp32 FFMA: thread = (row, head), ZD x 16 MACs, the row's Rbar_h in registers)
//   g = sigmoid(Wga agg + Gx) ; u = agg + g (S - agg) ; x1 = x + LN(Wo u + bo)
//   y = W2 relu(W1 LN(x1) + b1) + b2 ; out = x1 + LN(y) ; next layer: s | gx | q = {Ws, Wgx, Wq} LN_dst'(out), Qhat_h
// every tensor-core product a w ~ a_hi w_hi + a_lo w_hi + a_hi w_lo (tf32 hi / lo splits, fp32 accumulation in TMEM) in
// TWO MMAs per k-step: the activation operand holds its 32 hi rows and its 32 lo rows as one 64-row operand, so
// W_hi . [X_hi ; X_lo]^T is one N = 64 instruction (columns 0..31 = hi hi, 32..63 = hi lo) and W_lo . X_hi^T accumulates
// onto columns 32..63; the epilogue adds the two column halves.
//
// What was measured on the way (B200, 4096 rows, profiles/r2_*): SS-mode MMAs re-read their operands from shared memory
// per instruction, so at N = 32 the 4 KB weight operand -- not the math -- paces the tensor pipe (65 cycles per MMA with
// three MMAs per k-step); activation operands whose 128-byte core matrices were placed 144 bytes apart (to make the
// epilogue's stores conflict free) doubled that operand's fetch cost -> SWIZZLE_128B operands, conflict free for both
// store patterns AND aligned.  The weight stream is NOT the bound: multicasting it inside clusters of 2 / 4 CTAs changed
// nothing (60.8 / 60.7 / 59.9 us per launch), neither did skipping the copies altogether.  Rbar_h . Wvr'_h as eight
// M = 128 MMAs (96 of 128 operand rows unused) with hi / lo staging cost 13 k cycles per CTA; as fp32 FFMA it is exact
// and shorter.
//
// Thread roles (320 threads, 1 CTA / SM):
//   warps 0-7  epilogue.  Index maps over the CTA's [32 rows x 128 features] tile, 16 values per thread:
//                T-map (TMEM native): thread = feature f = 32 (warp % 4) + lane, rows 16 (warp / 4) .. + 15.  Biases are
//                  per-thread scalars; global rows [r][f] are read / written as coalesced 128-byte lines.
//                R-map (row major): thread = (row 4 warp + lane / 8, feature groups s + 8 i of 4 floats, s = lane % 8):
//                  LayerNorm statistics are three xor-shuffles inside 8 lanes.  T-map -> R-map goes through a padded
//                  [32][132] fp32 scratch tile and one named barrier.
//                H-map (agg only): thread = (row lane, head warp): weight reads are warp-wide broadcasts.
//   warp 8     weight producer (one thread, 4 x 32 KB ring on full / empty mbarriers)
//   warp 9     MMA issuer (one thread)
// TMEM: eight [128 lanes x 64 columns] accumulator slots; every accumulator has its own single-use "done" mbarrier, and
// a slot is rewritten only after every epilogue thread has read its previous contents AND arrived on a barrier the MMA
// thread waits on before that GEMM, so there are no "free" barriers:
//   gate 0, out 1, up_0 2, up_1 3, up_2 2, up_3 3, down 4, q 2, s 0, gx 1, Qhat_h: 3 4 5 6 7 2 0 1
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

constexpr int NR_MAX = 32;                 // destination rows per CTA: 32, or 16 for launches that would leave SMs idle
constexpr int NST = 4;                     // weight ring stages
constexpr int STAGE_BYTES = 32768;
constexpr int OPND_BYTES = 4 * 2 * NR_MAX * 128;   // four 32-wide k slabs of [NR hi rows | NR lo rows] x 128 B
constexpr int SCR_LD = 132;                // scratch row pitch (floats)
constexpr int EPI_THREADS = 256, EPI_WARPS = 8;
constexpr int THREADS = EPI_THREADS + 64;
constexpr int A_GATE = 0, A_OUT = 1, A_UP = 2, A_DOWN = 6, A_S = 7, A_GX = 8, A_Q = 9, A_QH = 10, NACC = 18;
__host__ __device__ constexpr uint32_t slot_col(int s) { return 64u * (uint32_t)s; }
__host__ __device__ constexpr int qh_slot(int h) { return h < 5 ? 3 + h : h == 5 ? 2 : h - 6; }

struct Smem {
  uint8_t actA[OPND_BYTES];                // activation operand: agg / u / LN(x1) / LN_dst(out)
  uint8_t actH[OPND_BYTES];                // FFN hidden operand, q operand
  uint8_t ring[NST][STAGE_BYTES];
  float scratch[NR_MAX * SCR_LD];
  float vec[tcp::V_SIZE];
  uint64_t full[NST], empty[NST];
  uint64_t opnd_ready, sgx_read, h_ready, h_free, acc_done[NACC];
  uint32_t tmem_base;
};
constexpr size_t SMEM_BYTES = sizeof(Smem) + 1024;

struct Args {
  const float *x, *rbar, *aggv, *s, *gx;   // [n][128], [n][8 zd], [n][128] ...
  float* out;                              // [n][128]
  float *q_n, *qhat_n, *s_n, *gx_n;        // next layer's destination-side projections (has_next)
  const float *W, *Wn;                     // packed layer weights (aw::), next layer's (nullptr = last layer);
                                           // W == nullptr: "pre-only" launch -- just Wn's destination-side projections of x
                                           // (the first layer of a stack, which has no previous node kernel to ride on)
  int n;
};

// K-major SWIZZLE_128B shared-memory matrix descriptor (8-row groups 1024 B apart; cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                        // leading byte offset: unused for swizzled K-major operands
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset
  d |= (uint64_t)1 << 46;                        // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// VPT accumulator values = hi.hi half (columns taddr ..) + cross-term half (columns taddr + NR ..), one wait for both loads
template <int VPT, int NR>
__device__ __forceinline__ void tmem_ld_pair(uint32_t taddr, float (&v)[VPT]) {
  uint32_t r[VPT], q[VPT];
  if constexpr (VPT == 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
          "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
        : "r"(taddr + NR)
        : "memory");
  } else {
    static_assert(VPT == 8, "8 or 16 rows per thread");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
                 : "r"(taddr + NR)
                 : "memory");
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < VPT; ++i) v[i] = __uint_as_float(r[i]) + __uint_as_float(q[i]);
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { tcp::mbar_arrive(e4::smem_u32(bar)); }

// phase timestamps of CTA 0 (SM clock) into tcp::g_tcp_dbg: [0..15] epilogue thread 0, [16..27] MMA thread, [28] cycles the
// MMA thread waited for weights, [29] for the epilogue warps (tools/sw_phases.py)
#define PSW_MARK(slot)                                                  \
  do {                                                                  \
    if (blockIdx.x == 0 && has_next) tcp::g_tcp_dbg[slot] = clock64(); \
  } while (0)

template <int ZD, int NR>
__global__ void __launch_bounds__(THREADS, 1) attn_post_sw_kernel(const __grid_constant__ Args a_in) {
  static_assert(NR == 32 || NR == 16, "rows per CTA");
  constexpr int SLAB_BYTES = 2 * NR * 128;   // one 32-wide k block of an activation operand: NR hi rows then NR lo rows, 128 B each
  constexpr int LO_OFF = NR * 128;
  constexpr int VPT = NR / 2;                // values per epilogue thread in every index map
  constexpr int LPR = EPI_THREADS / NR;      // R-map: lanes per row (8 / 16)
  constexpr int GPT = VPT / 4;               // R-map: 4-float feature groups per thread, group sg + LPR i
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (e4::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * NR;
  const float* __restrict__ W = a_in.W;
  const float* __restrict__ Wn = a_in.Wn;
  const bool has_next = Wn != nullptr;
  const bool pre_only = W == nullptr;
  constexpr int VR_HALF_BYTES = (ZD / 2) * D * 4;      // the fp32 Wvr' block [ZD][128] travels as two ring stages

  if (tid == 0 && blockIdx.x == 0 && has_next) tcp::g_tcp_dbg[30] = clock64();
  pdl_launch_dependents();   // the next kernel in the stream may be scheduled; it blocks in its own pdl_wait() until we are done
  if (warp == 0) tc::tmem_alloc(&sm.tmem_base, 512);
  const int n_chunks = has_next ? 58 : 42;
  const int chunk0 = pre_only ? 42 : 0;             // ring positions count from the first chunk of the launch
  // weight chunk i of the launch's script -> ring stage (i - chunk0) % NST (producer thread only)
  auto issue_chunk = [&](int i) {
    const float* src;
    uint32_t bytes;
    if (i < 2) {
      src = W + (ZD == 96 ? aw::WVRG96T : aw::WVRGT) + i * (VR_HALF_BYTES / 4);   // fp32 [ZD][128]: the FFMA agg
      bytes = VR_HALF_BYTES;
    } else if (i < 10) {
      src = W + aw::TC_GA + (i - 2) * 8192;      // Wga (4), Wo (4): contiguous
      bytes = 32768;
    } else if (i < 42) {
      src = W + aw::TC_FF2 + (i - 10) * 8192;    // up_0, up_1, down_0, up_2, down_1, up_3, down_2, down_3 (4 k-chunks each)
      bytes = 32768;
    } else if (i < 46) {
      src = Wn + aw::TC_Q + (i - 42) * 8192;     // next layer's Wq first (q feeds the Qhat GEMMs) ...
      bytes = 32768;
    } else if (i < 54) {
      src = Wn + aw::TC_S + (i - 46) * 8192;     // ... then Ws, Wgx (contiguous)
      bytes = 32768;
    } else {
      src = Wn + aw::TC_KRG + (i - 54) * 8192;   // Wkr' of two heads per stage
      bytes = 32768;
    }
    const int k = i - chunk0, s = k % NST;
    if (k >= NST) tcp::mbar_wait(&sm.empty[s], ((k / NST) - 1) & 1);
    const uint32_t fb = e4::smem_u32(&sm.full[s]);
    e4::mbar_expect_tx(fb, bytes);
    e4::bulk_copy(e4::smem_u32(sm.ring[s]), src, bytes, fb);
  };
  if (tid == EPI_THREADS) {   // the producer thread arms its own ring and starts the first NST copies under the rest of the set-up
    for (int i = 0; i < NST; ++i) {
      e4::mbar_init(e4::smem_u32(&sm.full[i]), 1);
      e4::mbar_init(e4::smem_u32(&sm.empty[i]), 1);
    }
    e4::fence_proxy_async();
    for (int i = chunk0; i < chunk0 + NST; ++i) issue_chunk(i);
  }
  if (tid == 32) {
    e4::mbar_init(e4::smem_u32(&sm.opnd_ready), EPI_WARPS);
    e4::mbar_init(e4::smem_u32(&sm.h_ready), EPI_WARPS);
    e4::mbar_init(e4::smem_u32(&sm.sgx_read), EPI_WARPS);
    e4::mbar_init(e4::smem_u32(&sm.h_free), 1);
    for (int i = 0; i < NACC; ++i) e4::mbar_init(e4::smem_u32(&sm.acc_done[i]), 1);
  }
  // per-column vectors -> shared memory (same table as tc_post.cuh)
  for (int i = tid; i < tcp::V_SIZE; i += THREADS) {
    float v = 0.f;
    if (i < tcp::V_LNDST_G) {
      if (pre_only) v = 0.f;
      else if (i < tcp::V_B1) v = W[aw::BO + i];
      else if (i < tcp::V_B2) v = W[aw::B1 + (i - tcp::V_B1)];
      else v = W[aw::B2 + (i - tcp::V_B2)];
    } else if (has_next) {
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
  // Everything above touched only the weights (constant during a forward) and this CTA's shared / tensor memory; from here on
  // the kernel reads what the previous kernel wrote and writes what it may still be reading.
  pdl_wait();
  // Row buffers are read with ld.global.cg (L2) from here on, never through the non-coherent path (__ldg / ld.global.nc), whose
  // contract ("read-only for the lifetime of the kernel") a CTA that became resident several grids early cannot honour: the
  // ping-pong q / s / gx / x buffers ARE rewritten between its launch and its pdl_wait().  A precaution, not a measured bug.
  Args a = a_in;   // the row buffers are only reachable through pointers acquired after the wait
  a.x = pdl_acquire(a.x); a.rbar = pdl_acquire(a.rbar); a.aggv = pdl_acquire(a.aggv); a.s = pdl_acquire(a.s); a.gx = pdl_acquire(a.gx);
  a.out = pdl_acquire(a.out); a.q_n = pdl_acquire(a.q_n); a.qhat_n = pdl_acquire(a.qhat_n); a.s_n = pdl_acquire(a.s_n);
  a.gx_n = pdl_acquire(a.gx_n);

  if (warp == 8) {
    // ================================================================== weight producer
    if (lane == 0) {
      for (int i = chunk0 + NST; i < n_chunks; ++i) issue_chunk(i);
    }
    __syncwarp();
  } else if (warp == 9) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      int ci = pre_only ? 0 : 2;                   // chunks 0 and 1 (fp32 Wvr') belong to the epilogue warps
      uint32_t ph_op = 0, ph_h = 0;
      long long w_full = 0, w_other = 0;
      auto wait_bar = [&](uint64_t* bar, uint32_t& ph) {
        const long long t0 = clock64();
        tcp::mbar_wait(bar, ph);
        w_other += clock64() - t0;
        ph ^= 1;
        tc::fence_after_sync();
      };
      auto next_chunk = [&]() -> uint32_t {
        const int s = ci % NST;
        const long long t0 = clock64();
        tcp::mbar_wait(&sm.full[s], (ci / NST) & 1);
        w_full += clock64() - t0;
        tc::fence_after_sync();
        return e4::smem_u32(sm.ring[s]);
      };
      // D^T[128 features][hi.hi | cross] (+)= Wchunk[128 x KC] . X[rows][k0 .. k0 + KC)^T ; k0 .. k0 + KC stays inside a 32-wide slab
      // `part` of `parts` weight blocks that share one ring stage
      auto gemm_sw = [&](auto kc_tag, uint32_t d_col, uint32_t opnd, int k0, bool accumulate, int part = 0, int parts = 1) {
        constexpr int KC_ = decltype(kc_tag)::value;
        const uint32_t wb = (part == 0 ? next_chunk() : e4::smem_u32(sm.ring[ci % NST])) + part * (2 * 128 * KC_ * 4);
        constexpr uint32_t idesc64 = tc::make_idesc_tf32(128, 2 * NR), idesc32 = tc::make_idesc_tf32(128, NR);
        const uint64_t wh0 = tc::make_smem_desc(wb, 128, KC_ * 32);
        const uint64_t wl0 = tc::make_smem_desc(wb + 128 * KC_ * 4, 128, KC_ * 32);
        const uint64_t x0 = make_desc_sw128(opnd + (k0 >> 5) * SLAB_BYTES + (k0 & 31) * 4);
#pragma unroll
        for (int ks = 0; ks < KC_ / 8; ++ks) {
          const uint64_t wh = wh0 + (uint64_t)(ks * 16), wl = wl0 + (uint64_t)(ks * 16);
          const uint64_t xd = x0 + (uint64_t)(ks * 2);                          // + 32 bytes inside the swizzled row
          tc::mma_tf32(tmem + d_col, wh, xd, idesc64, accumulate || ks > 0);    // [hi.hi | hi.lo]: rows 0..NR-1 hi, NR..2NR-1 lo
          tc::mma_tf32(tmem + d_col + NR, wl, xd, idesc32, true);               // + lo.hi onto the cross-term half
        }
        if (part == parts - 1) {
          tc::mma_commit(&sm.empty[ci % NST]);
          ++ci;
        }
      };
      using I16 = std::integral_constant<int, 16>;
      using I32 = std::integral_constant<int, 32>;
      auto gemm128 = [&](uint32_t d_col, uint32_t opnd) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) gemm_sw(I32{}, d_col, opnd, 32 * c, c > 0);
      };
      const uint32_t actA = e4::smem_u32(sm.actA), actH = e4::smem_u32(sm.actH);
      PSW_MARK(16);
      if (!pre_only) {
      // 1. gate, 2. out projection
      wait_bar(&sm.opnd_ready, ph_op);
      PSW_MARK(18);
      gemm128(slot_col(0), actA);
      tc::mma_commit(&sm.acc_done[A_GATE]);
      PSW_MARK(19);
      wait_bar(&sm.opnd_ready, ph_op);
      PSW_MARK(20);
      gemm128(slot_col(1), actA);
      tc::mma_commit(&sm.acc_done[A_OUT]);
      PSW_MARK(21);
      // 3. FFN: hidden tile t = 128 features; chunks arrive as up_0, up_1, down_0, up_2, down_1, up_3, down_2, down_3
      wait_bar(&sm.opnd_ready, ph_op);
      PSW_MARK(22);
      gemm128(slot_col(2), actA);
      tc::mma_commit(&sm.acc_done[A_UP + 0]);
      gemm128(slot_col(3), actA);
      tc::mma_commit(&sm.acc_done[A_UP + 1]);
#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        wait_bar(&sm.h_ready, ph_h);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) gemm_sw(I32{}, slot_col(4), actH, 32 * c, t > 0 || c > 0);
        if (t < 3) tc::mma_commit(&sm.h_free);
        else tc::mma_commit(&sm.acc_done[A_DOWN]);
        if (t + 2 < 4) {
          gemm128(slot_col(2 + t), actA);                 // up_{t+2} reuses the slot of up_t (read before h_ready was signalled)
          tc::mma_commit(&sm.acc_done[A_UP + t + 2]);
        }
      }
      }
      PSW_MARK(23);
      if (has_next) {
        // 4. next layer: q, s, gx from LN_dst'(out); Qhat_h = Wkr'_h[128 x 16] . q_h^T starts as soon as the q operand is
        //    written (that epilogue runs under the s / gx MMAs); the last two heads reuse the s / gx slots
        wait_bar(&sm.opnd_ready, ph_op);
        PSW_MARK(24);
        gemm128(slot_col(2), actA);
        tc::mma_commit(&sm.acc_done[A_Q]);
        gemm128(slot_col(0), actA);
        tc::mma_commit(&sm.acc_done[A_S]);
        gemm128(slot_col(1), actA);
        tc::mma_commit(&sm.acc_done[A_GX]);
        PSW_MARK(25);
        wait_bar(&sm.opnd_ready, ph_op);         // q operand written, slot 2 read
        PSW_MARK(26);
#pragma unroll 1
        for (int h = 0; h < H; ++h) {
          if (h == 6) {                                  // s and gx read: slots 0 and 1 are free (its own barrier: a warp may
            uint32_t ph0 = 0;                            // arrive here before another has arrived for the q operand)
            wait_bar(&sm.sgx_read, ph0);
          }
          gemm_sw(I16{}, slot_col(qh_slot(h)), actH, 16 * h, false, h & 1, 2);
          tc::mma_commit(&sm.acc_done[A_QH + h]);
        }
        PSW_MARK(27);
        if (blockIdx.x == 0) { tcp::g_tcp_dbg[28] = w_full; tcp::g_tcp_dbg[29] = w_other; }
      }
    }
    __syncwarp();
  } else {
    // ================================================================== epilogue warps
    const int q4 = warp & 3, ch = warp >> 2;
    const int f = q4 * 32 + lane;                          // T-map: feature, rows VPT ch .. VPT ch + VPT - 1
    const uint32_t lb = tmem + ((uint32_t)(q4 * 32) << 16);
    const int rr = (NR / 8) * warp + lane / LPR, sg = lane % LPR;   // R-map: row, feature groups sg + LPR i
    const bool rr_ok = row0 + rr < a.n;
    const float* vec = sm.vec;
    float* scr = sm.scratch;
    uint32_t ph_hfree = 0;
    auto wait_acc = [&](int i) {
      tcp::mbar_wait(&sm.acc_done[i], 0);
      tc::fence_after_sync();
    };
    // this warp's operand stores are complete and visible to the tensor core: one arrival per warp
    auto publish = [&](uint64_t* bar) {
      e4::fence_proxy_async();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // T-map: VPT rows of this thread's feature -> hi / lo rows of the swizzled activation operand at `opnd`
    auto t_store_opnd = [&](uint8_t* opnd, const float (&v)[VPT]) {
      uint8_t* p = opnd + q4 * SLAB_BYTES + ((VPT / 8) * ch) * 1024 + (lane & 3) * 4;
      const int c = lane >> 2;                             // 16-byte chunk of the row before the swizzle
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const float hi = tcp::tf32_rna_finite(v[i]);
        const float lo = tcp::tf32_rna_finite(v[i] - hi);
        uint8_t* pe = p + (i >> 3) * 1024 + (i & 7) * 128 + ((c ^ (i & 7)) << 4);
        *reinterpret_cast<float*>(pe) = hi;
        *reinterpret_cast<float*>(pe + LO_OFF) = lo;
      }
    };
    // R-map: this thread's GPT groups of 4 features of row rr (group g = sg + LPR i = chunk g % 8 of slab g / 8)
    auto r_store_opnd = [&](uint8_t* opnd, const float (&v)[VPT]) {
      uint8_t* p = opnd + (rr >> 3) * 1024 + (rr & 7) * 128;
#pragma unroll
      for (int i = 0; i < GPT; ++i) {
        const int g = sg + LPR * i;
        uint8_t* pe = p + (g >> 3) * SLAB_BYTES + (((g & 7) ^ (rr & 7)) << 4);
        float4 hi, lo;
        hi.x = tcp::tf32_rna_finite(v[4 * i]);     lo.x = tcp::tf32_rna_finite(v[4 * i] - hi.x);
        hi.y = tcp::tf32_rna_finite(v[4 * i + 1]); lo.y = tcp::tf32_rna_finite(v[4 * i + 1] - hi.y);
        hi.z = tcp::tf32_rna_finite(v[4 * i + 2]); lo.z = tcp::tf32_rna_finite(v[4 * i + 2] - hi.z);
        hi.w = tcp::tf32_rna_finite(v[4 * i + 3]); lo.w = tcp::tf32_rna_finite(v[4 * i + 3] - hi.w);
        *reinterpret_cast<float4*>(pe) = hi;
        *reinterpret_cast<float4*>(pe + LO_OFF) = lo;
      }
    };
    auto r_load_scr = [&](float (&v)[VPT]) {
#pragma unroll
      for (int i = 0; i < GPT; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(scr + rr * SCR_LD + 4 * (sg + LPR * i));
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    };
    auto r_load_vec = [&](const float* p, float (&v)[VPT]) {
#pragma unroll
      for (int i = 0; i < GPT; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * (sg + LPR * i));
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    };
    auto r_load_glb = [&](const float* base, float (&v)[VPT]) {   // row rr of a [n][128] global matrix
#pragma unroll
      for (int i = 0; i < GPT; ++i) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr_ok) t = __ldcg(reinterpret_cast<const float4*>(base + (size_t)(row0 + rr) * D) + sg + LPR * i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    };
    auto r_store_glb = [&](float* base, const float (&v)[VPT]) {
      if (rr_ok) {
#pragma unroll
        for (int i = 0; i < GPT; ++i)
          *(reinterpret_cast<float4*>(base + (size_t)(row0 + rr) * D) + sg + LPR * i) =
              make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    };
    // Sum over the 128 features of a row from per-group partials t[i] (group sg + LPR i), as ONE fixed tree for both row
    // counts: level 1 pairs groups g and g + 16, level 2 pairs g and g + 8 (g < 8), then xor-shuffles 1, 2, 4 -- in registers
    // where a thread holds both partners, by shuffle where it does not.  A row's statistics therefore do not depend on NR.
    auto row_tree = [&](const float (&t)[GPT]) -> float {
      float p;
      if constexpr (NR == 32) {
        p = (t[0] + t[2]) + (t[1] + t[3]);
      } else {
        p = t[0] + t[1];
        p += __shfl_xor_sync(0xffffffffu, p, 8);
      }
      p += __shfl_xor_sync(0xffffffffu, p, 1);
      p += __shfl_xor_sync(0xffffffffu, p, 2);
      p += __shfl_xor_sync(0xffffffffu, p, 4);
      return p;
    };
    // two-pass LayerNorm of the row (R-map), affine from vec + g_off / b_off
    auto r_layernorm = [&](float (&v)[VPT], int g_off, int b_off) {
      float t[GPT];
#pragma unroll
      for (int i = 0; i < GPT; ++i) t[i] = ((v[4 * i] + v[4 * i + 1]) + v[4 * i + 2]) + v[4 * i + 3];
      const float mean = row_tree(t) * (1.0f / 128.0f);
#pragma unroll
      for (int i = 0; i < GPT; ++i) {
        const float d0 = v[4 * i] - mean, d1 = v[4 * i + 1] - mean, d2 = v[4 * i + 2] - mean, d3 = v[4 * i + 3] - mean;
        t[i] = fmaf(d3, d3, fmaf(d2, d2, fmaf(d1, d1, d0 * d0)));
      }
      const float rstd = 1.0f / sqrtf(row_tree(t) * (1.0f / 128.0f) + LN_EPS);
      float g[VPT], b[VPT];
      r_load_vec(vec + g_off, g);
      r_load_vec(vec + b_off, b);
#pragma unroll
      for (int i = 0; i < VPT; ++i) v[i] = (v[i] - mean) * rstd * g[i] + b[i];
    };
    // T-map global access: element [row0 + VPT ch + i][f]
    auto t_load_glb = [&](const float* base, float (&v)[VPT]) {
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int r = row0 + VPT * ch + i;
        v[i] = r < a.n ? __ldcg(base + (size_t)r * D + f) : 0.f;
      }
    };
    auto t_store_glb = [&](float* base, size_t ld, int col, const float (&v)[VPT]) {
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int r = row0 + VPT * ch + i;
        if (r < a.n) base[(size_t)r * ld + col + f] = v[i];
      }
    };
    auto t_store_scr = [&](const float (&v)[VPT]) {
#pragma unroll
      for (int i = 0; i < VPT; ++i) scr[(VPT * ch + i) * SCR_LD + f] = v[i];
    };
#define PSW_EMARK(slot) do { if (tid == 0) PSW_MARK(slot); } while (0)
    PSW_EMARK(0);
    float v[VPT];
    if (pre_only) {            // LN_dst'(x) -> operand A, then straight to the projections
      r_load_glb(a.x, v);
      r_layernorm(v, tcp::V_LNDST_G, tcp::V_LNDST_B);
      r_store_opnd(sm.actA, v);
      publish(&sm.opnd_ready);
    } else {
    // ---- 0. agg[row][16 h ..] = (AggV + sum_{d < ZD/2} Rbar[row][h][d] Wvr'[d][16 h ..]) + sum_{d >= ZD/2} ...     (H-map)
    //         fp32 FFMA: thread = (row, head = warp), for NR = 16 also (half of d) = lane / 16; Rbar_h in registers, the
    //         weights as broadcasts from the ring (two stages hold the fp32 [ZD][128] block).  Two ascending-d chains, one per
    //         weight half, added at the end: the same arithmetic whether one thread runs both (NR = 32) or two lanes one each.
    {
      // Rbar reaches its thread in passes of PW floats through a warp-private staging tile of 32 "slots" (NR = 32: slot =
      // row; NR = 16: slot = 16 half + row): a pass is read from global memory as whole 128-byte lines and re-read slot-wise
      // from shared memory.  (Reading it straight into the owning thread -- 16 bytes per lane at a 3 KB stride -- cost ~7 k
      // cycles of LSU address divergence per CTA: 28 instructions x 8 warps x 32 distinct lines.)  The tile lives in the
      // activation operands, which nobody touches before the agg result is written.
      constexpr int PW = ZD == 96 ? 48 : 32, NPASS = ZD / PW, P4 = PW / 4, PITCH = PW + 4;
      constexpr int PPH = NPASS / 2;                        // passes per weight half (= ring stage)
      constexpr int NLOOP = NR == 32 ? NPASS : PPH;         // NR = 16: both halves advance together
      static_assert(8 * 32 * PITCH * 4 <= 2 * OPND_BYTES, "staging tiles must fit in actA + actH");
      const int arow = NR == 32 ? lane : (lane & 15), ahalf = NR == 32 ? 0 : (lane >> 4);
      const bool ok = row0 + arow < a.n;
      float* stg = reinterpret_cast<float*>(sm.actA) + warp * (32 * PITCH);
      float4 g[P4], rb[P4];
      float2 av[8], av2[8];
      auto load_pass = [&](int pass) {
#pragma unroll
        for (int it = 0; it < P4; ++it) {
          const int idx = it * 32 + lane, slot = idx / P4, c4 = idx % P4;
          const int r = NR == 32 ? slot : (slot & 15);
          const int dpass = NR == 32 ? pass : (slot >> 4) * PPH + pass;
          g[it] = row0 + r < a.n
                      ? __ldcg(reinterpret_cast<const float4*>(a.rbar + (size_t)(row0 + r) * (H * ZD) + warp * ZD + dpass * PW) + c4)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load_pass(0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && ahalf == 0) t4 = __ldcg(reinterpret_cast<const float4*>(a.aggv + (size_t)(row0 + arow) * D + 16 * warp) + i);
        av[2 * i] = make_float2(t4.x, t4.y);
        av[2 * i + 1] = make_float2(t4.z, t4.w);
        av2[2 * i] = av2[2 * i + 1] = make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int pass = 0; pass < NLOOP; ++pass) {
#pragma unroll
        for (int it = 0; it < P4; ++it) {
          const int idx = it * 32 + lane, slot = idx / P4, c4 = idx % P4;
          *reinterpret_cast<float4*>(stg + slot * PITCH + 4 * c4) = g[it];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < P4; ++i) rb[i] = *reinterpret_cast<const float4*>(stg + lane * PITCH + 4 * i);
        __syncwarp();
        if (pass + 1 < NLOOP) load_pass(pass + 1);          // in flight under this pass's FMAs
        const int half = NR == 32 ? pass / PPH : ahalf;
        if (NR == 32 ? (pass % PPH == 0) : (pass == 0)) {
          if (NR == 32) {
            tcp::mbar_wait(&sm.full[half], 0);
          } else {
            tcp::mbar_wait(&sm.full[0], 0);
            tcp::mbar_wait(&sm.full[1], 0);
          }
          if (half == 0) PSW_EMARK(2); else PSW_EMARK(15);
        }
        const float* wv = reinterpret_cast<const float*>(sm.ring[half]) + (pass % PPH) * PW * D + 16 * warp;
        auto chain = [&](float2 (&acc)[8]) {
#pragma unroll
          for (int dq = 0; dq < P4; ++dq) {
            const float4 r4 = rb[dq];
            const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float* wr = wv + (dq * 4 + j) * D;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 w4 = *reinterpret_cast<const float4*>(wr + 4 * i);
                const float2 rr2 = make_float2(rv[j], rv[j]);      // packed FFMA2: two columns per instruction (same rounding)
                acc[2 * i] = __ffma2_rn(rr2, make_float2(w4.x, w4.y), acc[2 * i]);
                acc[2 * i + 1] = __ffma2_rn(rr2, make_float2(w4.z, w4.w), acc[2 * i + 1]);
              }
            }
          }
        };
        if (NR == 32 && half == 1) chain(av2); else chain(av);
      }
      PSW_EMARK(14);
      if constexpr (NR == 16) {   // the second-half chain lives in lane + 16
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          av2[i].x = __shfl_xor_sync(0xffffffffu, av[i].x, 16);
          av2[i].y = __shfl_xor_sync(0xffffffffu, av[i].y, 16);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        av[i].x += av2[i].x;
        av[i].y += av2[i].y;
      }
      if (NR == 32 || lane < 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(scr + arow * SCR_LD + 16 * warp + 4 * i) = make_float4(av[2 * i].x, av[2 * i].y, av[2 * i + 1].x, av[2 * i + 1].y);
      }
      epi_barrier();
      if (tid == 0) {          // every epilogue thread has read the two weight stages
        mbar_arrive(&sm.empty[0]);
        mbar_arrive(&sm.empty[1]);
      }
      PSW_EMARK(1);
      r_load_scr(v);           // agg stays in the scratch tile (fp32) for the gate
      r_store_opnd(sm.actA, v);
      publish(&sm.opnd_ready);
      PSW_EMARK(3);
    }
    // ---- 1. gate: g = sigmoid(acc + Gx) ; u = agg + g (S - agg) -> operand A        (T-map)
    {
      float gxv[VPT], sv[VPT];
      t_load_glb(a.gx, gxv);
      t_load_glb(a.s, sv);
      wait_acc(A_GATE);
      PSW_EMARK(4);
      tmem_ld_pair<VPT, NR>(lb + slot_col(0) + VPT * ch, v);
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const float ag = scr[(VPT * ch + i) * SCR_LD + f];
        const float g = 1.0f / (1.0f + expf(-(v[i] + gxv[i])));
        v[i] = ag + g * (sv[i] - ag);
      }
      t_store_opnd(sm.actA, v);
      publish(&sm.opnd_ready);
      PSW_EMARK(5);
    }
    // ---- 2. o = acc + bo ; x1 = x + LN_post(o) (kept in registers, R-map) ; LN_ffpre(x1) -> operand A
    float x1[VPT];
    {
      r_load_glb(a.x, x1);
      const float bo = vec[tcp::V_BO + f];
      wait_acc(A_OUT);
      PSW_EMARK(6);
      tmem_ld_pair<VPT, NR>(lb + slot_col(1) + VPT * ch, v);
#pragma unroll
      for (int i = 0; i < VPT; ++i) v[i] += bo;
      t_store_scr(v);
      epi_barrier();
      r_load_scr(v);
      r_layernorm(v, tcp::V_LNPOST_G, tcp::V_LNPOST_B);
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        x1[i] += v[i];
        v[i] = x1[i];
      }
      r_layernorm(v, tcp::V_LNFFPRE_G, tcp::V_LNFFPRE_B);
      r_store_opnd(sm.actA, v);
      publish(&sm.opnd_ready);
      PSW_EMARK(7);
    }
    // ---- 3. FFN hidden tiles: h_t = relu(acc + b1) -> operand H                     (T-map)
#pragma unroll 1
    for (int tt = 0; tt < 4; ++tt) {
      const float b1 = vec[tcp::V_B1 + 128 * tt + f];
      wait_acc(A_UP + tt);
      tmem_ld_pair<VPT, NR>(lb + slot_col(2 + (tt & 1)) + VPT * ch, v);
#pragma unroll
      for (int i = 0; i < VPT; ++i) v[i] = fmaxf(v[i] + b1, 0.f);
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
      PSW_EMARK(8);
      wait_acc(A_DOWN);
      PSW_EMARK(9);
      tmem_ld_pair<VPT, NR>(lb + slot_col(4) + VPT * ch, v);
#pragma unroll
      for (int i = 0; i < VPT; ++i) v[i] += b2;
      t_store_scr(v);
      epi_barrier();
      r_load_scr(v);
      r_layernorm(v, tcp::V_LNFFPOST_G, tcp::V_LNFFPOST_B);
#pragma unroll
      for (int i = 0; i < VPT; ++i) v[i] += x1[i];
      r_store_glb(a.out, v);
      if (has_next) {
        r_layernorm(v, tcp::V_LNDST_G, tcp::V_LNDST_B);
        r_store_opnd(sm.actA, v);
        publish(&sm.opnd_ready);
      }
      PSW_EMARK(10);
    }
    }
    if (has_next) {
      // ---- 4. q, s, gx (+ bias) -> global as coalesced lines (T-map); q also -> operand H
      {
        const float bq = vec[tcp::V_BQ + f];
        wait_acc(A_Q);
        PSW_EMARK(11);
        tmem_ld_pair<VPT, NR>(lb + slot_col(2) + VPT * ch, v);
#pragma unroll
        for (int i = 0; i < VPT; ++i) v[i] += bq;
        t_store_opnd(sm.actH, v);             // every MMA that read the hidden operand completed before acc_done[DOWN]
        publish(&sm.opnd_ready);              // q operand written, slot 2 read
        PSW_EMARK(12);
        t_store_glb(a.q_n, D, 0, v);
      }
      {
        const float bs = vec[tcp::V_BS + f];
        wait_acc(A_S);
        tmem_ld_pair<VPT, NR>(lb + slot_col(0) + VPT * ch, v);
#pragma unroll
        for (int i = 0; i < VPT; ++i) v[i] += bs;
        t_store_glb(a.s_n, D, 0, v);
      }
      {
        const float bg = vec[tcp::V_BG + f];
        wait_acc(A_GX);
        tmem_ld_pair<VPT, NR>(lb + slot_col(1) + VPT * ch, v);
#pragma unroll
        for (int i = 0; i < VPT; ++i) v[i] += bg;
        publish(&sm.sgx_read);                // slots 0 and 1 read: the last two Qhat GEMMs may overwrite them
        t_store_glb(a.gx_n, D, 0, v);
      }
      // ---- 5. Qhat_h: the eight small GEMMs are issued back to back; stream them out as they complete
#pragma unroll 1
      for (int h = 0; h < H; ++h) {
        wait_acc(A_QH + h);
        tmem_ld_pair<VPT, NR>(lb + slot_col(qh_slot(h)) + VPT * ch, v);
        t_store_glb(a.qhat_n, (size_t)H * D, h * D, v);
      }
      PSW_EMARK(13);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace psw
}  // namespace prosim
