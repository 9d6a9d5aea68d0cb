// tcgen05 (5th-gen tensor core) building blocks for the dense node-level GEMMs, fp32-equivalent via 3xTF32.
//
//   C[128 x N] = A[128 x K] * W[N x K]^T      A, W fp32; every product a*w is evaluated as
//   a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with a_hi = tf32(a), a_lo = tf32(a - a_hi) (same for w), fp32 accumulate
//   in tensor memory: relative error ~2^-21 per product, i.e. fp32-class, which the 1e-5 per-tick parity gate needs
//   (a single TF32 pass, 2^-11, does not pass it).
//
// Operands sit in shared memory in the canonical K-major, no-swizzle UMMA layout (8-row x 16-byte core matrices):
//   byte(m, k) = (m / 8) * SBO + (k / 4) * 128 + (m % 8) * 16 + (k % 4) * 4,   SBO = (KC / 4) * 128,  LBO = 128
// One elected thread issues tcgen05.mma (kind::tf32, cta_group::1, M = 128, N <= 256, K = 8 per instruction);
// the accumulator lives in TMEM (lane = row, column = n) and comes back with tcgen05.ld.32x32b.
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prosim {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                        // descriptor version 1 (Blackwell)
  return d;                                      // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive accumulator columns of this thread's TMEM lane (lane = 32 * (warp % 4) + lane id)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// byte offset of element (m, k) of a [rows x KC] K-major no-swizzle operand tile
__device__ __forceinline__ uint32_t umma_off(int m, int k, int kc) {
  return (uint32_t)((m >> 3) * (kc * 32) + (k >> 2) * 128 + (m & 7) * 16 + (k & 3) * 4);
}

// ---------------------------------------------------------------------------------------------------------------
// Validation kernel: C[M x 128] = A[M x 128] * W[128 x 128]^T, one CTA (128 threads) per 128-row tile.
// split3 = 0: single TF32 pass; 1: 3xTF32.  K is processed in two chunks of 64 (operands: 4 x 32 KB).
// ---------------------------------------------------------------------------------------------------------------
constexpr int TEST_KC = 64;
constexpr size_t TEST_SMEM = 4 * 128 * TEST_KC * sizeof(float) + 64;

__global__ void __launch_bounds__(128) tc_gemm_test_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                           float* __restrict__ C, int M, int split3) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sAhi = smem;
  uint8_t* sAlo = sAhi + 128 * TEST_KC * 4;
  uint8_t* sBhi = sAlo + 128 * TEST_KC * 4;
  uint8_t* sBlo = sBhi + 128 * TEST_KC * 4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBlo + 128 * TEST_KC * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * 128;

  if (warp == 0) tmem_alloc(tmem_slot, 128);
  if (tid == 0) mbar_init(bar, 1);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = make_idesc_tf32(128, 128);
  constexpr uint32_t SBO = TEST_KC * 32, LBO = 128;
  uint32_t phase = 0;

  for (int kc0 = 0; kc0 < 128; kc0 += TEST_KC) {
    // stage operands (conflict-free mapping: 8 rows x 4 k-vectors per warp instruction)
    for (int i = tid; i < 128 * (TEST_KC / 4); i += 128) {
      const int blk = i >> 5, l = i & 31;                 // 32 float4 per block: rows (l & 7), k-vectors (l >> 3)
      const int rg = blk % 16, kvg = blk / 16;            // row group of 8, group of 4 k-vectors
      const int m = rg * 8 + (l & 7), k = (kvg * 4 + (l >> 3)) * 4;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + m < M) a = *reinterpret_cast<const float4*>(A + (size_t)(row0 + m) * 128 + kc0 + k);
      const float4 w = *reinterpret_cast<const float4*>(W + (size_t)m * 128 + kc0 + k);
      const float4 ah = make_float4(to_tf32(a.x), to_tf32(a.y), to_tf32(a.z), to_tf32(a.w));
      const float4 wh = make_float4(to_tf32(w.x), to_tf32(w.y), to_tf32(w.z), to_tf32(w.w));
      const uint32_t off = umma_off(m, k, TEST_KC);
      *reinterpret_cast<float4*>(sAhi + off) = ah;
      *reinterpret_cast<float4*>(sBhi + off) = wh;
      *reinterpret_cast<float4*>(sAlo + off) =
          make_float4(to_tf32(a.x - ah.x), to_tf32(a.y - ah.y), to_tf32(a.z - ah.z), to_tf32(a.w - ah.w));
      *reinterpret_cast<float4*>(sBlo + off) =
          make_float4(to_tf32(w.x - wh.x), to_tf32(w.y - wh.y), to_tf32(w.z - wh.z), to_tf32(w.w - wh.w));
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      for (int ks = 0; ks < TEST_KC / 8; ++ks) {
        const uint32_t adv = ks * 2 * LBO;              // 8 tf32 = two 16-byte k-vectors
        const uint64_t ah = make_smem_desc(smem_u32(sAhi) + adv, LBO, SBO), al = make_smem_desc(smem_u32(sAlo) + adv, LBO, SBO);
        const uint64_t bh = make_smem_desc(smem_u32(sBhi) + adv, LBO, SBO), bl = make_smem_desc(smem_u32(sBlo) + adv, LBO, SBO);
        const bool first = kc0 == 0 && ks == 0;
        if (split3) {
          mma_tf32(tmem, al, bh, idesc, !first);
          mma_tf32(tmem, ah, bl, idesc, true);
          mma_tf32(tmem, ah, bh, idesc, true);
        } else {
          mma_tf32(tmem, ah, bh, idesc, !first);
        }
      }
      mma_commit(bar);
    }
    mbar_wait(bar, phase);      // MMAs of this chunk have consumed the operand tiles
    phase ^= 1;
    fence_after_sync();
    __syncthreads();
  }
  // epilogue: warp q owns TMEM lanes 32q..32q+31 = rows; 4 x 32 columns
  const int m = warp * 32 + lane;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    if (row0 + m < M) {
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(C + (size_t)(row0 + m) * 128 + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace tc
}  // namespace prosim
