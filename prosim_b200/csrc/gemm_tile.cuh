// Row-tile fp32 GEMM core, version 2: weights staged through shared memory.
//
//   acc[r][c] += sum_k A[row(r)][k] * Wt[k][col(c)]      M = 16*RT rows per CTA, N = 128 columns
//
// 8 warps as 4 row groups x 2 column halves; inside a warp tx = lane & 15 owns 4 columns, ty = lane >> 4 owns
// TR = 2*RT rows: a warp covers 4*RT rows x 64 columns, a thread keeps TR x 4 fp32 accumulators.  (A first
// layout gave every warp all 128 columns of 2*RT rows; ncu showed it shared-memory bound -- 82 % of the
// LDS bandwidth at 31 % FMA -- because each warp re-read the whole weight row for only 16 FMAs.)
//   * A (activations) lives in shared memory, row-major with a padded stride; reads are LDS.128 with
//     two distinct addresses per warp (broadcast).
//   * Wt (K-major weights, n contiguous) is streamed from L2 in chunks of KC = 32 k-rows (16 KB) through a
//     3-stage cp.async ring shared by the CTA: each weight element is fetched ONCE per CTA and reused by all
//     16*RT rows, instead of once per thread row-group as in gemm_tile_acc (v1).  The ring keeps running
//     across consecutive GEMMs of a fused kernel: the last iteration of one GEMM already prefetches the first
//     chunk of the next (WPipe::primed), so a chain of small GEMMs has no pipeline bubbles.
//   * B reads are conflict-free LDS.128 (16 lanes x 16 B contiguous: 2 wavefronts per k feed 8*RT FMA instructions
//     of the warp), one __syncthreads per chunk.
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

// thread -> tile coordinates
struct TileCoord {
  int row;   // first row of this thread inside the CTA tile
  int col;   // first of its 4 columns
};
template <int RT>
__device__ __forceinline__ TileCoord tile_coord() {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TileCoord t;
  t.row = (warp >> 1) * (4 * RT) + (lane >> 4) * (2 * RT);
  t.col = (warp & 1) * 64 + (lane & 15) * 4;
  return t;
}

template <int RT>
__device__ __forceinline__ void acc2_init(float (&acc)[2 * RT][4], float v) {
#pragma unroll
  for (int r = 0; r < 2 * RT; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = v;
}

// bias (or any per-column vector) into every row's accumulators
template <int RT>
__device__ __forceinline__ void acc2_init_bias(float (&acc)[2 * RT][4], const float* __restrict__ bias) {
  const float4 b = __ldg(reinterpret_cast<const float4*>(bias + tile_coord<RT>().col));
#pragma unroll
  for (int r = 0; r < 2 * RT; ++r) {
    acc[r][0] = b.x; acc[r][1] = b.y; acc[r][2] = b.z; acc[r][3] = b.w;
  }
}

// All four columns of a thread lie inside one attention head, so the block-diagonal Wvr' contraction is the same
// routine with a per-thread A base (attn2.cuh).
// next_Wt != nullptr: prefetch chunk 0 of the next GEMM (next_K rows, clipped to KC) while finishing this one.
template <int RT>
__device__ __forceinline__ void gemm_tile2(float (&acc)[2 * RT][4], const float* A, int lda, int K,
                                           const float* __restrict__ Wt, int ldw, WPipe& p,
                                           const float* __restrict__ next_Wt, int next_ldw, int next_K) {
  constexpr int TR = 2 * RT;
  const TileCoord tc = tile_coord<RT>();
  const float* a0 = A + tc.row * lda;
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
    const float* Bs = p.buf + p.st * WCHUNK_FLOATS + tc.col;
    const int kc = min(KC, K - c * KC);
    const int kbase = c * KC;
#pragma unroll 2
    for (int kk = 0; kk < kc; kk += 4) {
      float4 av[TR];
#pragma unroll
      for (int r = 0; r < TR; ++r) av[r] = *reinterpret_cast<const float4*>(a0 + r * lda + kbase + kk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(Bs + (kk + j) * 128);
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          const float x = j == 0 ? av[r].x : j == 1 ? av[r].y : j == 2 ? av[r].z : av[r].w;
          // Blackwell packed fp32 FMA (SASS FFMA2, scalar-broadcast first operand): two IEEE fmas per issue slot
          const float2 lo = __ffma2_rn(make_float2(x, x), make_float2(b.x, b.y), make_float2(acc[r][0], acc[r][1]));
          const float2 hi = __ffma2_rn(make_float2(x, x), make_float2(b.z, b.w), make_float2(acc[r][2], acc[r][3]));
          acc[r][0] = lo.x; acc[r][1] = lo.y; acc[r][2] = hi.x; acc[r][3] = hi.y;
        }
      }
    }
    p.st = nxt;
  }
  p.primed = next_Wt != nullptr;
}

// convenience wrapper with default "no prefetch" arguments
template <int RT>
__device__ __forceinline__ void gemm2(float (&acc)[2 * RT][4], const float* A, int lda, int K, const float* __restrict__ Wt,
                                      int ldw, WPipe& p, const float* __restrict__ next_Wt = nullptr, int next_ldw = 128,
                                      int next_K = KC) {
  gemm_tile2<RT>(acc, A, lda, K, Wt, ldw, p, next_Wt, next_ldw, next_K);
}

// ---- epilogue helpers for the accumulator layout
template <int RT>
__device__ __forceinline__ void acc2_store_smem(const float (&acc)[2 * RT][4], float* dst, int ld, bool relu) {
  const TileCoord tc = tile_coord<RT>();
#pragma unroll
  for (int r = 0; r < 2 * RT; ++r) {
    float4 v = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    *reinterpret_cast<float4*>(dst + (tc.row + r) * ld + tc.col) = v;
  }
}

// rows row0 + r < N are written; dst row stride ldg floats, column offset col0
template <int RT>
__device__ __forceinline__ void acc2_store_global(const float (&acc)[2 * RT][4], float* __restrict__ dst, size_t ldg,
                                                  int col0, int row0, int N) {
  const TileCoord tc = tile_coord<RT>();
#pragma unroll
  for (int r = 0; r < 2 * RT; ++r) {
    const int row = row0 + tc.row + r;
    if (row < N)
      *reinterpret_cast<float4*>(dst + (size_t)row * ldg + col0 + tc.col) =
          make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
}

// acc[r][c] = src[row][col(c)] for valid rows, 0 otherwise
template <int RT>
__device__ __forceinline__ void acc2_load_global(float (&acc)[2 * RT][4], const float* __restrict__ src, size_t ldg,
                                                 int row0, int N) {
  const TileCoord tc = tile_coord<RT>();
#pragma unroll
  for (int r = 0; r < 2 * RT; ++r) {
    const int row = row0 + tc.row + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < N) v = *reinterpret_cast<const float4*>(src + (size_t)row * ldg + tc.col);
    acc[r][0] = v.x; acc[r][1] = v.y; acc[r][2] = v.z; acc[r][3] = v.w;
  }
}

}  // namespace prosim
