// Row-tile fp32 GEMM core, version 2: weights staged through shared memory.
//
//   acc[r][c] += sum_k A[row(r)][k] * Wt[k][col(c)]      M = 16*RT rows per CTA, N = 128 columns
//
// 256 threads as a 16 x 16 grid: tx = tid & 15 owns columns {4tx..4tx+3} and {64+4tx..64+4tx+3},
// ty = tid >> 4 owns rows {ty*RT .. ty*RT+RT-1}  ->  RT x 8 fp32 accumulators per thread.
//   * A (activations) lives in shared memory, row-major with a padded stride; reads are LDS.128 with
//     two distinct addresses per warp (broadcast).
//   * Wt (K-major weights, n contiguous) is streamed from L2 in chunks of KC = 32 k-rows (16 KB) through a
//     3-stage cp.async ring shared by the CTA: each weight element is fetched ONCE per CTA and reused by all
//     16*RT rows, instead of once per thread row-group as in gemm_tile_acc (v1).  The ring keeps running
//     across consecutive GEMMs of a fused kernel: the last iteration of one GEMM already prefetches the first
//     chunk of the next (WPipe::primed), so a chain of small GEMMs has no pipeline bubbles.
//   * B reads are conflict-free LDS.128 (16 lanes x 16 B contiguous), one __syncthreads per chunk.
// Accumulation over k is in ascending order for every output: deterministic and batch invariant.
#pragma once
#include "common.cuh"

namespace prosim {

constexpr int KC = 32;                     // k rows per weight chunk
constexpr int WSTAGES = 3;
constexpr int WCHUNK_FLOATS = KC * 128;    // one stage
constexpr size_t WPIPE_BYTES = (size_t)WSTAGES * WCHUNK_FLOATS * sizeof(float);   // 48 KB

struct WPipe {
  float* buf;    // smem [WSTAGES][KC][128]
  int st;        // stage holding (or about to hold) the next chunk to consume
  bool primed;   // that chunk's cp.async group has already been issued by the previous GEMM
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// all 256 threads: copy rows [k0, k0+kc) x 128 columns of Wt (row stride ldw) into one stage
__device__ __forceinline__ void wpipe_issue(float* stage, const float* __restrict__ Wt, int ldw, int k0, int kc) {
#pragma unroll
  for (int i = 0; i < (KC * 32) / 256; ++i) {
    const int idx = threadIdx.x + 256 * i;
    const int row = idx >> 5, c4 = idx & 31;
    if (row < kc) cp_async16(stage + row * 128 + c4 * 4, Wt + (size_t)(k0 + row) * ldw + c4 * 4);
  }
}

template <int RT>
__device__ __forceinline__ void acc2_init(float (&acc)[RT][8], float v) {
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = v;
}

// bias (or any per-column vector) into every row's accumulators
template <int RT>
__device__ __forceinline__ void acc2_init_bias(float (&acc)[RT][8], const float* __restrict__ bias) {
  const int tx = threadIdx.x & 15;
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + 4 * tx));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + 64 + 4 * tx));
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    acc[r][0] = b0.x; acc[r][1] = b0.y; acc[r][2] = b0.z; acc[r][3] = b0.w;
    acc[r][4] = b1.x; acc[r][5] = b1.y; acc[r][6] = b1.z; acc[r][7] = b1.w;
  }
}

// A0 feeds columns 0..63 of this thread, A1 columns 64..127 (normally A0 == A1; they differ only for the
// block-diagonal Wvr' contraction where the A row depends on the output head).
// next_Wt != nullptr: prefetch chunk 0 of the next GEMM (next_K rows, clipped to KC) while finishing this one.
template <int RT, bool TWO_A>
__device__ __forceinline__ void gemm_tile2(float (&acc)[RT][8], const float* A0, const float* A1, int lda, int K,
                                           const float* __restrict__ Wt, int ldw, WPipe& p,
                                           const float* __restrict__ next_Wt, int next_ldw, int next_K) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* a0 = A0 + ty * RT * lda;
  const float* a1 = A1 + ty * RT * lda;
  const int nch = (K + KC - 1) / KC;
  if (!p.primed) {
    wpipe_issue(p.buf + p.st * WCHUNK_FLOATS, Wt, ldw, 0, min(KC, K));
    cp_async_commit();
  }
  for (int c = 0; c < nch; ++c) {
    const int nxt = (p.st + 1) % WSTAGES;
    if (c + 1 < nch) wpipe_issue(p.buf + nxt * WCHUNK_FLOATS, Wt, ldw, (c + 1) * KC, min(KC, K - (c + 1) * KC));
    else if (next_Wt != nullptr) wpipe_issue(p.buf + nxt * WCHUNK_FLOATS, next_Wt, next_ldw, 0, min(KC, next_K));
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* Bs = p.buf + p.st * WCHUNK_FLOATS;
    const int kc = min(KC, K - c * KC);
    const int kbase = c * KC;
#pragma unroll 2
    for (int kk = 0; kk < kc; kk += 4) {
      float4 av0[RT], av1[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        av0[r] = *reinterpret_cast<const float4*>(a0 + r * lda + kbase + kk);
        if (TWO_A) av1[r] = *reinterpret_cast<const float4*>(a1 + r * lda + kbase + kk);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b0 = *reinterpret_cast<const float4*>(Bs + (kk + j) * 128 + 4 * tx);
        const float4 b1 = *reinterpret_cast<const float4*>(Bs + (kk + j) * 128 + 64 + 4 * tx);
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const float x0 = j == 0 ? av0[r].x : j == 1 ? av0[r].y : j == 2 ? av0[r].z : av0[r].w;
          const float x1 = !TWO_A ? x0 : (j == 0 ? av1[r].x : j == 1 ? av1[r].y : j == 2 ? av1[r].z : av1[r].w);
          acc[r][0] = fmaf(x0, b0.x, acc[r][0]);
          acc[r][1] = fmaf(x0, b0.y, acc[r][1]);
          acc[r][2] = fmaf(x0, b0.z, acc[r][2]);
          acc[r][3] = fmaf(x0, b0.w, acc[r][3]);
          acc[r][4] = fmaf(x1, b1.x, acc[r][4]);
          acc[r][5] = fmaf(x1, b1.y, acc[r][5]);
          acc[r][6] = fmaf(x1, b1.z, acc[r][6]);
          acc[r][7] = fmaf(x1, b1.w, acc[r][7]);
        }
      }
    }
    p.st = nxt;
  }
  p.primed = next_Wt != nullptr;
}

// convenience wrapper for the common single-A case
template <int RT>
__device__ __forceinline__ void gemm2(float (&acc)[RT][8], const float* A, int lda, int K, const float* __restrict__ Wt,
                                      int ldw, WPipe& p, const float* __restrict__ next_Wt = nullptr, int next_ldw = 128,
                                      int next_K = KC) {
  gemm_tile2<RT, false>(acc, A, A, lda, K, Wt, ldw, p, next_Wt, next_ldw, next_K);
}

// ---- epilogue helpers for the (tx, ty) accumulator layout
template <int RT>
__device__ __forceinline__ void acc2_store_smem(const float (&acc)[RT][8], float* dst, int ld, bool relu) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    float4 v0 = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    float4 v1 = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
    if (relu) {
      v0 = make_float4(fmaxf(v0.x, 0.f), fmaxf(v0.y, 0.f), fmaxf(v0.z, 0.f), fmaxf(v0.w, 0.f));
      v1 = make_float4(fmaxf(v1.x, 0.f), fmaxf(v1.y, 0.f), fmaxf(v1.z, 0.f), fmaxf(v1.w, 0.f));
    }
    float* d = dst + (ty * RT + r) * ld;
    *reinterpret_cast<float4*>(d + 4 * tx) = v0;
    *reinterpret_cast<float4*>(d + 64 + 4 * tx) = v1;
  }
}

// rows row0 + ty*RT + r < N are written; dst row stride ldg floats, column offset col0
template <int RT>
__device__ __forceinline__ void acc2_store_global(const float (&acc)[RT][8], float* __restrict__ dst, size_t ldg,
                                                  int col0, int row0, int N) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    const int row = row0 + ty * RT + r;
    if (row < N) {
      float* d = dst + (size_t)row * ldg + col0;
      *reinterpret_cast<float4*>(d + 4 * tx) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      *reinterpret_cast<float4*>(d + 64 + 4 * tx) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
    }
  }
}

// acc[r][c] = src[row][col(c)] for valid rows, 0 otherwise
template <int RT>
__device__ __forceinline__ void acc2_load_global(float (&acc)[RT][8], const float* __restrict__ src, size_t ldg, int row0,
                                                 int N) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    const int row = row0 + ty * RT + r;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (row < N) {
      v0 = *reinterpret_cast<const float4*>(src + (size_t)row * ldg + 4 * tx);
      v1 = *reinterpret_cast<const float4*>(src + (size_t)row * ldg + 64 + 4 * tx);
    }
    acc[r][0] = v0.x; acc[r][1] = v0.y; acc[r][2] = v0.z; acc[r][3] = v0.w;
    acc[r][4] = v1.x; acc[r][5] = v1.y; acc[r][6] = v1.z; acc[r][7] = v1.w;
  }
}

}  // namespace prosim
