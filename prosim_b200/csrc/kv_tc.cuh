// K' | V' of an attention layer's sources on the tensor cores (tcgen05 + TMEM, 3xTF32).
//
// Same math as attn_kv2_kernel (attn2.cuh; reference prosim/models/layers/attention_layer.py:67-72 with the folds of
// weights_layout.h):  K' = Wk LN_src(x) + Wkr beta_r,  V' = Wv LN_src(x) + b_v + Wvr beta_r + b_vr.
// One CTA (128 threads, thread = source row = TMEM lane) per 128 rows and layer, built on the chunk pipeline of
// pointnet_tc.cuh: the row's LayerNorm is thread local; K' and V' are separate CTAs (blockIdx.z) that take every second
// chunk of the interleaved weight block [Wk c0, Wv c0, Wk c1, ...] (aw::TC_KV).  Rows enter and leave through a padded shared-memory tile so that
// global accesses are whole 512-byte rows.
#pragma once
#include "pointnet_tc.cuh"

namespace prosim {
namespace kvtc {
constexpr int TILE_LD = 132;
constexpr size_t SMEM_BYTES = 1024 + 2 * pntc::A_BYTES + 2 * pntc::B_STAGE_BYTES + 64 + 4 * 512;
static_assert(2 * pntc::A_BYTES + 2 * pntc::B_STAGE_BYTES >= 128 * TILE_LD * 4, "row tile aliases the operand buffers");
}  // namespace kvtc

__global__ void __launch_bounds__(128, 2) attn_kv_tc_kernel(const float* __restrict__ X, int N, const float* __restrict__ Wbase,
                                                            size_t w_layer_stride, float* __restrict__ KV,
                                                            size_t kv_layer_stride) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (e4::smem_u32(smem_raw) & 1023u)) & 1023u);
  const float* W = Wbase + (size_t)blockIdx.y * w_layer_stride;
  float* kv = KV + (size_t)blockIdx.y * kv_layer_stride;
  pntc::Pipe p;
  p.sAhi = base;
  p.sAlo = base + pntc::A_BYTES;
  p.sB = base + 2 * pntc::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p.sB + 2 * pntc::B_STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* sVec = reinterpret_cast<float*>(bars + 8);    // LN_src gamma, beta, K' bias, V' bias: read by every thread
  p.bfull = bars;
  p.bmma = bars + 2;
  const int half = blockIdx.z;            // 0: K' (kv columns 0..127), 1: V' (128..255) -- one CTA each: twice the CTAs,
                                          // half the serial chain (a 4096-row x 6-layer launch is 192 tiles on 296 CTA slots)
  p.wtc = W + aw::TC_KV;
  p.wmul = 2;
  p.woff = half;
  p.chunk = 0;
  p.total = 4;
  float* tile = reinterpret_cast<float*>(base);     // [128][132] fp32 over the (idle) operand buffers
  const int m = threadIdx.x, warp = m >> 5;
  const int row0 = blockIdx.x * 128;

  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  if (m == 32) {
    tc::mbar_init(&p.bfull[0], 1);
    tc::mbar_init(&p.bfull[1], 1);
    tc::mbar_init(p.bmma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  sVec[m] = __ldg(W + aw::LN_SRC_G + m);
  sVec[128 + m] = __ldg(W + aw::LN_SRC_B + m);
  sVec[256 + m] = __ldg(W + aw::KB + m);
  sVec[384 + m] = __ldg(W + aw::VB + m);
  // rows -> tile (coalesced float4), then each thread takes its own row
  for (int i = m; i < 128 * 32; i += 128) {
    const int r = i >> 5, c4 = i & 31;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) t = __ldg(reinterpret_cast<const float4*>(X + (size_t)(row0 + r) * D) + c4);
    *reinterpret_cast<float4*>(tile + r * kvtc::TILE_LD + 4 * c4) = t;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  p.tmem = *tmem_slot;
  float v[128];
#pragma unroll
  for (int i = 0; i < 128; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(tile + m * kvtc::TILE_LD + i);
    v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
  }
  {   // LN_src with affine, thread local
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
      const float4 g = *reinterpret_cast<const float4*>(sVec + i);
      const float4 b = *reinterpret_cast<const float4*>(sVec + 128 + i);
      v[i] = (v[i] - mean) * rstd * g.x + b.x;
      v[i + 1] = (v[i + 1] - mean) * rstd * g.y + b.y;
      v[i + 2] = (v[i + 2] - mean) * rstd * g.z + b.z;
      v[i + 3] = (v[i + 3] - mean) * rstd * g.w + b.w;
    }
  }
  // The tile is about to become operand buffers again, and the first weight chunk is written into it by the bulk-copy
  // engine (async proxy).  A barrier alone does not order another thread's still-pending shared-memory LOADS before that
  // write: with the copy issued right after the row reads, ~1 % of the launches returned a few corrupted rows 62..127 (the
  // part of the tile under weight stage 0).  The LayerNorm above has consumed every loaded value; the proxy fence orders
  // this thread's generic accesses before the async ones that follow the barrier.
  tc::fence_async_smem();
  __syncthreads();
  if (m == 0) pntc::load_b(p, 0);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float a[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = v[c * 32 + i];
    pntc::chunk_mma(p, a, c == 0);
  }
  // accumulator -> tile -> global
  pntc::read_acc(p, sVec + 256 + half * 128, v);
#pragma unroll
  for (int i = 0; i < 128; i += 4)
    *reinterpret_cast<float4*>(tile + m * kvtc::TILE_LD + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  __syncthreads();
  for (int i = m; i < 128 * 32; i += 128) {
    const int r = i >> 5, c4 = i & 31;
    if (row0 + r < N)
      *(reinterpret_cast<float4*>(kv + (size_t)(row0 + r) * 256 + half * 128) + c4) =
          *reinterpret_cast<const float4*>(tile + r * kvtc::TILE_LD + 4 * c4);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(p.tmem, 256);
}

}  // namespace prosim
