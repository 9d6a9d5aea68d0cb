// Row-tile fp32 GEMM core, version 2: weights staged through shared memory.
//
//   acc[r][c] += sum_k A[row(r)][k] * Wt[k][col(c)]      M = NW*TR rows per CTA, N = 128 columns
//
// NW warps (8 or 16) as NW/2 row groups x 2 column halves; inside a warp tx = lane & 15 owns 4 columns,
// ty = lane >> 4 owns TR rows: a warp covers 2*TR rows x 64 columns, the CTA M = NW*TR rows, a thread keeps
// TR x 4 fp32 accumulators.  Measured on B200 (policy layer, 4096 rows): 8 warps x TR = 4 -> 104 us, 16 warps x
// TR = 2 -> 128 us: halving the per-warp weight-row reuse costs more shared-memory bandwidth than the extra
// warps buy in latency hiding, so api.cu instantiates NW = 8.  (A first
// layout gave every warp all 128 columns of 2*RT rows; ncu showed it shared-memory bound -- 82 % of the
// LDS bandwidth at 31 % FMA -- because each warp re-read the whole weight row for only 16 FMAs.)
//   * A (activations) lives in shared memory, row-major with a padded stride; reads are LDS.128 with
//     two distinct addresses per warp (broadcast).
//   * Wt (K-major weights, n contiguous) is streamed from L2 in chunks of KC = 32 k-rows (16 KB) through a
//     4-stage cp.async ring shared by the CTA: each weight element is fetched ONCE per CTA and reused by all
//     16*RT rows, instead of once per thread row-group as in gemm_tile_acc (v1).  The ring runs 2 chunks ahead
//     of the math across the whole chain of GEMMs of a fused kernel (WPipe), so a chain of small GEMMs has no
//     pipeline bubbles.
//   * B reads are conflict-free LDS.128 (16 lanes x 16 B contiguous: 2 wavefronts per k feed 8*RT FMA instructions
//     of the warp), one __syncthreads per chunk.
// Accumulation over k is in ascending order for every output: deterministic and batch invariant.
#pragma once
#include "common.cuh"

namespace prosim {

constexpr int KC = 32;                     // k rows per weight chunk
constexpr int WSTAGES = 4;                 // ring stages
constexpr int WAHEAD = 2;                  // chunks in flight ahead of the one being consumed
constexpr int WCHUNK_FLOATS = KC * 128;    // one stage
constexpr int WMAXSEG = 24;                // GEMMs (weight segments) a fused kernel may chain
constexpr size_t WPIPE_BYTES = (size_t)WSTAGES * WCHUNK_FLOATS * sizeof(float) + WMAXSEG * 16;   // 64 KB + table

// One GEMM's weight operand: K rows x 128 columns starting at w, row stride ldw.
struct WSeg {
  const float* w;
  int ldw;
  int K;
};

// Weight stream of a fused kernel: the kernel lists every GEMM it will run, in order, once; the producer side
// then runs WAHEAD chunks ahead of the consumer ACROSS GEMM boundaries, so neither the L2 latency (~1 us, longer
// than one 32-row chunk of math) nor the start of a small GEMM ever stalls the FMA pipes.
struct WPipe {
  float* buf;        // smem [WSTAGES][KC][128]
  WSeg* segs;        // smem table
  int nseg;
  int pseg, pk;      // producer cursor: segment, k offset inside it
  int issued;        // chunks issued so far (global chunk counter -> stage = counter % WSTAGES)
  int consumed;      // chunks consumed so far
  int cseg;          // consumer cursor: next segment to multiply
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// all NW*32 threads: copy rows [k0, k0+kc) x 128 columns of Wt (row stride ldw) into one stage
template <int NW>
__device__ __forceinline__ void wpipe_issue(float* stage, const float* __restrict__ Wt, int ldw, int k0, int kc) {
#pragma unroll
  for (int i = 0; i < (KC * 32) / (NW * 32); ++i) {
    const int idx = threadIdx.x + NW * 32 * i;
    const int row = idx >> 5, c4 = idx & 31;
    if (row < kc) cp_async16(stage + row * 128 + c4 * 4, Wt + (size_t)(k0 + row) * ldw + c4 * 4);
  }
}

// issue the producer's next chunk (if any); always commits a group so group counting stays uniform
template <int NW>
__device__ __forceinline__ void wpipe_produce(WPipe& p) {
  if (p.pseg < p.nseg) {
    const WSeg sg = p.segs[p.pseg];
    const int kc = min(KC, sg.K - p.pk);
    wpipe_issue<NW>(p.buf + (p.issued % WSTAGES) * WCHUNK_FLOATS, sg.w, sg.ldw, p.pk, kc);
    ++p.issued;
    p.pk += kc;
    if (p.pk >= sg.K) { ++p.pseg; p.pk = 0; }
  }
  cp_async_commit();
}

// smem: [ring | table].  `fill(segs)` (run by thread 0) writes the kernel's GEMM list and returns its length.
template <int NW, typename F>
__device__ __forceinline__ WPipe wpipe_init(float* smem, F fill) {
  WPipe p;
  p.buf = smem;
  p.segs = reinterpret_cast<WSeg*>(smem + WSTAGES * WCHUNK_FLOATS);
  __shared__ int s_nseg;
  if (threadIdx.x == 0) s_nseg = fill(p.segs);
  __syncthreads();
  p.nseg = s_nseg;
  p.pseg = p.pk = p.issued = p.consumed = p.cseg = 0;
#pragma unroll
  for (int i = 0; i < WAHEAD; ++i) wpipe_produce<NW>(p);
  return p;
}

// thread -> tile coordinates
struct TileCoord {
  int row;   // first row of this thread inside the CTA tile
  int col;   // first of its 4 columns
};
template <int TR>
__device__ __forceinline__ TileCoord tile_coord() {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TileCoord t;
  t.row = (warp >> 1) * (2 * TR) + (lane >> 4) * TR;
  t.col = (warp & 1) * 64 + (lane & 15) * 4;
  return t;
}

template <int TR>
__device__ __forceinline__ void acc2_init(float (&acc)[TR][4], float v) {
#pragma unroll
  for (int r = 0; r < TR; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = v;
}

// bias (or any per-column vector) into every row's accumulators
template <int TR>
__device__ __forceinline__ void acc2_init_bias(float (&acc)[TR][4], const float* __restrict__ bias) {
  const float4 b = __ldg(reinterpret_cast<const float4*>(bias + tile_coord<TR>().col));
#pragma unroll
  for (int r = 0; r < TR; ++r) {
    acc[r][0] = b.x; acc[r][1] = b.y; acc[r][2] = b.z; acc[r][3] = b.w;
  }
}

// All four columns of a thread lie inside one attention head, so the block-diagonal Wvr' contraction is the same
// routine with a per-thread A base (attn2.cuh).  Multiplies by the pipe's NEXT weight segment.
template <int TR, int NW>
__device__ __forceinline__ void gemm2(float (&acc)[TR][4], const float* A, int lda, WPipe& p) {
  const TileCoord tc = tile_coord<TR>();
  const float* a0 = A + tc.row * lda;
  const int K = p.segs[p.cseg].K;
  ++p.cseg;
  for (int kbase = 0; kbase < K; kbase += KC) {
    wpipe_produce<NW>(p);             // keeps WAHEAD chunks in flight beyond the one consumed below
    cp_async_wait<WAHEAD>();
    __syncthreads();
    const float* Bs = p.buf + (p.consumed % WSTAGES) * WCHUNK_FLOATS + tc.col;
    ++p.consumed;
    const int kc = min(KC, K - kbase);
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
  }
}

// ---- epilogue helpers for the accumulator layout
template <int TR>
__device__ __forceinline__ void acc2_store_smem(const float (&acc)[TR][4], float* dst, int ld, bool relu) {
  const TileCoord tc = tile_coord<TR>();
#pragma unroll
  for (int r = 0; r < TR; ++r) {
    float4 v = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    *reinterpret_cast<float4*>(dst + (tc.row + r) * ld + tc.col) = v;
  }
}

// rows row0 + r < N are written; dst row stride ldg floats, column offset col0
template <int TR>
__device__ __forceinline__ void acc2_store_global(const float (&acc)[TR][4], float* __restrict__ dst, size_t ldg,
                                                  int col0, int row0, int N) {
  const TileCoord tc = tile_coord<TR>();
#pragma unroll
  for (int r = 0; r < TR; ++r) {
    const int row = row0 + tc.row + r;
    if (row < N)
      *reinterpret_cast<float4*>(dst + (size_t)row * ldg + col0 + tc.col) =
          make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
}

// acc[r][c] = src[row][col(c)] for valid rows, 0 otherwise
template <int TR>
__device__ __forceinline__ void acc2_load_global(float (&acc)[TR][4], const float* __restrict__ src, size_t ldg,
                                                 int row0, int N) {
  const TileCoord tc = tile_coord<TR>();
#pragma unroll
  for (int r = 0; r < TR; ++r) {
    const int row = row0 + tc.row + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < N) v = *reinterpret_cast<const float4*>(src + (size_t)row * ldg + tc.col);
    acc[r][0] = v.x; acc[r][1] = v.y; acc[r][2] = v.z; acc[r][3] = v.w;
  }
}

}  // namespace prosim
