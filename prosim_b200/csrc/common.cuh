// Shared device helpers for the prosim_b200 kernels (sm_100a).
//
// Numerics policy (DESIGN.md "Numerics"): everything is IEEE fp32 -- no fast-math, no TF32/BF16 on
// the value path -- because the closed loop amplifies rounding by x1.5-3 per tick and parity with
// the reference's fp32 PyTorch path is gated at 1e-5 per tick.  Reductions run in a fixed order
// (no float atomics), so results are bit-reproducible run to run and independent of batch size.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

namespace prosim {

constexpr int D = 128;        // hidden dim (waymo_demo.yaml:217)
constexpr int H = 8;          // heads
constexpr int DH = 16;        // head dim
constexpr int LDS_PAD = 132;  // smem row stride for [rows][128] tiles: 16B aligned, conflict-free float4 rows
constexpr float LN_EPS = 1e-5f;

#define PROSIM_CHECK_LAUNCH()                       \
  do {                                              \
    cudaError_t e__ = cudaGetLastError();           \
    if (e__ != cudaSuccess) return (int)e__;        \
  } while (0)

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may be scheduled as soon as every CTA of its
// predecessor in the stream has executed pdl_launch_dependents() (or exited); it must execute pdl_wait() -- which returns when
// the predecessor grid has completed and its memory operations are visible -- before it reads or writes anything the
// predecessor (or anything before it) touches.  Used for the two kernels that alternate in the single-scene layer loop
// (edge_row.cuh, post_sw.cuh): the next kernel's launch latency and set-up (barrier init, TMEM allocation, the first weight
// copies) run under the tail of the current one.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// A pointer to data the predecessor produced, "acquired" after pdl_wait(): the value passes through a volatile asm, so no load
// through it -- not even a read-only (ld.global.nc / __restrict__) one, which the compiler may otherwise move freely because
// the kernel promises the data is constant during its lifetime -- can be scheduled before the wait.
template <typename T>
__device__ __forceinline__ T* pdl_acquire(T* p) {
  asm volatile("" : "+l"(p)::"memory");
  return p;
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {   // PROSIM_NO_PDL=1: plain stream-ordered launches (A/B and fault isolation)
  static const bool on = std::getenv("PROSIM_NO_PDL") == nullptr;
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  // sm_100a: one CREDUX.MAX.F32 into a uniform register instead of five SHFL + FMNMX pairs (max is exact: same result)
  float m;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}

// torch.remainder semantics for a positive divisor: result in [0, b)
__device__ __forceinline__ float py_mod(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.0f && m < 0.0f) m += b;
  return m;
}
// models/utils/geometry.py:13-17 in fp32: -pi + (a + pi) % (2 pi); constants rounded to fp32 like torch does
__device__ __forceinline__ float wrap_angle(float a) {
  const float PI_F = 3.14159265358979323846f;
  const float TWO_PI_F = 6.28318530717958647692f;
  return __fadd_rn(-PI_F, py_mod(__fadd_rn(a, PI_F), TWO_PI_F));
}

// LayerNorm of one 128-wide row held 4 values per lane (lane l owns columns 4l..4l+3), two-pass.
// Returns the normalised values WITHOUT affine; mean/rstd out for callers that need them.
__device__ __forceinline__ float4 ln_row_noaffine(float4 v) {
  float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.0f / D);
  float4 c = make_float4(v.x - mean, v.y - mean, v.z - mean, v.w - mean);
  float var = warp_sum((c.x * c.x + c.y * c.y) + (c.z * c.z + c.w * c.w)) * (1.0f / D);
  float rstd = 1.0f / sqrtf(var + LN_EPS);
  return make_float4(c.x * rstd, c.y * rstd, c.z * rstd, c.w * rstd);
}
__device__ __forceinline__ float4 ln_row(float4 v, const float* __restrict__ g, const float* __restrict__ b, int lane) {
  float4 n = ln_row_noaffine(v);
  float4 gg = *reinterpret_cast<const float4*>(g + 4 * lane);
  float4 bb = *reinterpret_cast<const float4*>(b + 4 * lane);
  return make_float4(n.x * gg.x + bb.x, n.y * gg.y + bb.y, n.z * gg.z + bb.z, n.w * gg.w + bb.w);
}

// In-place LayerNorm (+ optional ReLU) over `rows` rows of a [rows][ld] smem tile of width W (64 or 128),
// one warp per row, all warps of the CTA cooperating.  Caller syncs before and after.
template <int W>
__device__ __forceinline__ void ln_tile_inplace(float* tile, int ld, int rows, const float* __restrict__ g,
                                                const float* __restrict__ b, bool relu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  constexpr int PER = W / 32;
  for (int r = warp; r < rows; r += nwarps) {
    float v[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[i] = tile[r * ld + lane * PER + i]; s += v[i]; }
    float mean = warp_sum(s) * (1.0f / W);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[i] -= mean; q += v[i] * v[i]; }
    float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / W) + LN_EPS);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      float o = v[i] * rstd * g[lane * PER + i] + b[lane * PER + i];
      tile[r * ld + lane * PER + i] = relu ? fmaxf(o, 0.f) : o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row-tile GEMM on CUDA cores:  Y[r][n] = sum_k X[r][k] * Wt[k][n]   (r < 2*RPT, n < 128)
//   X  : shared memory, row stride ldx floats (16B aligned rows), K multiple of 4
//   Wt : global, K-major "transposed" weight [K][ldw] (n contiguous => coalesced across the CTA)
// 256 threads: thread = (column n = tid & 127, row group tid >> 7); each thread keeps RPT fp32
// accumulators, every weight it loads is reused RPT times, X reads are warp-broadcast LDS.128.
// The k loop runs in ascending order for every output => deterministic, batch-invariant sums.
// ---------------------------------------------------------------------------------------------
template <int RPT>
__device__ __forceinline__ void gemm_tile_acc(float (&acc)[RPT], const float* __restrict__ Xs, int ldx, int K,
                                              const float* __restrict__ Wt, int ldw) {
  const int n = threadIdx.x & 127;
  const float* xrow = Xs + (threadIdx.x >> 7) * RPT * ldx;
  const float* w = Wt + n;
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    float w0 = __ldg(w + (k + 0) * ldw);
    float w1 = __ldg(w + (k + 1) * ldw);
    float w2 = __ldg(w + (k + 2) * ldw);
    float w3 = __ldg(w + (k + 3) * ldw);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      float4 x = *reinterpret_cast<const float4*>(xrow + r * ldx + k);
      acc[r] = fmaf(x.x, w0, acc[r]);
      acc[r] = fmaf(x.y, w1, acc[r]);
      acc[r] = fmaf(x.z, w2, acc[r]);
      acc[r] = fmaf(x.w, w3, acc[r]);
    }
  }
}

// Same contraction with packed fp32 FMAs: thread = (column pair cp = tid & 63, row group tid >> 6), RPT rows x 2
// columns per thread, 256 threads cover 4*RPT rows x 128 columns.  One FFMA2 (scalar-broadcast x, a natural weight
// pair, a natural accumulator pair) does the work of two FFMAs, and a k step costs RPT/4 LDS.128 + one LDG.64 per
// thread instead of RPT/4... twice that per output column: half the issue slots per MAC.  Ascending k per output, so
// results are bit-identical to gemm_tile_acc.
template <int RPT>
__device__ __forceinline__ void gemm_tile_acc2(float2 (&acc)[RPT], const float* __restrict__ Xs, int ldx, int K,
                                               const float* __restrict__ Wt, int ldw) {
  const int cp = threadIdx.x & 63;
  const float* xrow = Xs + (threadIdx.x >> 6) * RPT * ldx;
  const float2* w = reinterpret_cast<const float2*>(Wt) + cp;
  const int ldw2 = ldw >> 1;
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    const float2 w0 = __ldg(w + (k + 0) * ldw2);
    const float2 w1 = __ldg(w + (k + 1) * ldw2);
    const float2 w2 = __ldg(w + (k + 2) * ldw2);
    const float2 w3 = __ldg(w + (k + 3) * ldw2);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(xrow + r * ldx + k);
      acc[r] = __ffma2_rn(make_float2(x.x, x.x), w0, acc[r]);
      acc[r] = __ffma2_rn(make_float2(x.y, x.y), w1, acc[r]);
      acc[r] = __ffma2_rn(make_float2(x.z, x.z), w2, acc[r]);
      acc[r] = __ffma2_rn(make_float2(x.w, x.w), w3, acc[r]);
    }
  }
}

template <int RPT>
__device__ __forceinline__ void acc_init(float (&acc)[RPT], float v) {
#pragma unroll
  for (int r = 0; r < RPT; ++r) acc[r] = v;
}

// Store this thread's accumulators into a smem tile column (row group aware).
template <int RPT>
__device__ __forceinline__ void acc_store_smem(const float (&acc)[RPT], float* Ys, int ldy, bool relu) {
  const int n = threadIdx.x & 127;
  float* y = Ys + (threadIdx.x >> 7) * RPT * ldy + n;
#pragma unroll
  for (int r = 0; r < RPT; ++r) y[r * ldy] = relu ? fmaxf(acc[r], 0.f) : acc[r];
}

// Load a [rows x 128] tile from global (row stride 128, optional row index list) into smem (stride ld),
// zero-filling rows >= nrows.  float4 per thread, coalesced.
__device__ __forceinline__ void load_tile128(float* dst, int ld, const float* __restrict__ src, int row0, int nrows,
                                             int tile_rows) {
  for (int i = threadIdx.x; i < tile_rows * 32; i += blockDim.x) {
    int r = i >> 5, c = (i & 31) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < nrows) v = *reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * D + c);
    *reinterpret_cast<float4*>(dst + r * ld + c) = v;
  }
}

}  // namespace prosim
