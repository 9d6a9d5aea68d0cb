// C-ABI entry points (include/prosim_b200.h): argument checks + kernel launches, nothing else.
#include "../../include/prosim_b200.h"

#include "attn.cuh"
#include "attn2.cuh"
#include "edge4.cuh"
#include "common.cuh"
#include "graph.cuh"
#include "pointnet.cuh"
#include "pointnet_tc.cuh"
#include "kv_tc.cuh"
#include <cstdlib>
#include "rollout.cuh"
#include "tc_gemm.cuh"
#include "tc_post.cuh"
#include "post_sw.cuh"
#include "edge_row.cuh"
#include "weights_layout.h"

#include <cudaTypedefs.h>

using namespace prosim;

namespace {

constexpr int g_rt2_min_rows = 128;   // row count from which the 32-row gemm_tile kernels (dstpre2, head2; bit-identical) replace the
                                      // 2..16-row ones: measured on one 128-agent scene, 12.07 ms per forward at 1024, 11.82 at 128
int g_split = 1;        // row-split chains of prosim_attn_stack_fwd (prosim_set_stack_split); measured 31.4 / 30.9 / 30.6 / 31.4 ms at 1..4 parts
bool g_use_tc = true;   // node-side GEMMs on tcgen05 (tc_post.cuh); prosim_set_tensor_core(0) selects the FFMA kernels
// Kernel selection mask (prosim_set_tensor_core; initial value from the environment variable PROSIM_TC_MASK for fault isolation
// from a fresh process): bit 0 node kernels on tcgen05, bit 1 K'|V', bit 2 PointNet, bit 3 the 16 / 32-row "swapped" node kernel
// (post_sw.cuh) incl. its pre-only mode, bit 4 the one-launch edge kernel of small launches (edge_row.cuh)
int g_tc_mask = std::getenv("PROSIM_TC_MASK") ? std::atoi(std::getenv("PROSIM_TC_MASK")) & 31 : 31;
constexpr int SW_MAX_ROWS = 148 * 32 * 2;   // above two waves of 32-row CTAs the 128-row kernel streams 4x less weight per row
// 16 rows per CTA while that still fits one wave: twice the CTAs, half the per-CTA epilogue work (bit-identical to 32 rows)
inline bool sw_rows16(int n) { return (n + 15) / 16 <= 148; }
constexpr int FUSED_EDGE_MAX_ROWS = 592;    // small launches: one fused edge launch instead of three (see launch_edge)
constexpr int ERR_ARG = -1;
constexpr int ERR_WORKSPACE = -2;

inline cudaStream_t S(prosim_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- launch accounting (bench.py: "gpu_launches") and optional per-kernel-class CUDA-event timing (roofline).
// Kernel classes: see PROSIM_K_* in the header.  Timing is off unless prosim_profile_enable() was called.
constexpr int N_CLASSES = 16;
constexpr int MAX_REC = 8192;
long long g_launches[N_CLASSES] = {0};
int g_prof_class = -1;
int g_prof_n = 0;
cudaEvent_t g_ev0[MAX_REC], g_ev1[MAX_REC];
bool g_ev_made = false;

struct LaunchScope {
  int cls;
  cudaStream_t st;
  bool timed;
  LaunchScope(int c, cudaStream_t s) : cls(c), st(s), timed(false) {
    ++g_launches[c];
    if (c == g_prof_class && g_prof_n < MAX_REC) {
      timed = true;
      cudaEventRecord(g_ev0[g_prof_n], st);
    }
  }
  ~LaunchScope() {
    if (timed) cudaEventRecord(g_ev1[g_prof_n++], st);
  }
};

// Row-tile height (2*RPT rows per CTA): the largest tile that still yields >= one CTA per SM.
inline int pick_rpt(int n_rows) {
  const int sms = 148;
  for (int rpt = 8; rpt > 1; rpt >>= 1)
    if ((n_rows + 2 * rpt - 1) / (2 * rpt) >= sms) return rpt;
  return 1;
}

// v2 (gemm_tile.cuh) kernels for throughput-sized launches: rows per CTA = 16*RT; 0 => use the v1 latency kernels
inline int pick_rt(int n_rows) {
  if (n_rows >= 64 * 120) return 4;
  if (n_rows >= g_rt2_min_rows) return 2;
  return 0;
}

#define DISPATCH_RPT(rpt, ...)                     \
  switch (rpt) {                                   \
    case 8: { constexpr int RPT = 8; __VA_ARGS__; } break; \
    case 4: { constexpr int RPT = 4; __VA_ARGS__; } break; \
    case 2: { constexpr int RPT = 2; __VA_ARGS__; } break; \
    default: { constexpr int RPT = 1; __VA_ARGS__; } break; \
  }

template <typename K>
inline cudaError_t allow_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int setup_attributes() {
  // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device: one state per device
  static int states[64] = {0};   // 0 = not done, 1 = ok, > 1 = error + 1000
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return ERR_ARG;
  int& state = states[dev];
  if (state != 0) return state == 1 ? 0 : state - 1000;
  cudaError_t e = cudaSuccess;
  auto acc = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
#define E4_ATTR(ZD, NW) acc(allow_smem(attn_edge4_kernel<ZD, NW>, Edge4Cfg<ZD>::smem_bytes(NW)))
  E4_ATTR(96, 1); E4_ATTR(96, 2); E4_ATTR(96, 4); E4_ATTR(96, 8); E4_ATTR(96, 12);
  E4_ATTR(128, 1); E4_ATTR(128, 2); E4_ATTR(128, 4); E4_ATTR(128, 8); E4_ATTR(128, 10);
#undef E4_ATTR
  acc(allow_smem(attn_edge_row_kernel<96>, EdgeRowCfg<96>::smem_bytes()));
  acc(allow_smem(attn_edge_row_kernel<128>, EdgeRowCfg<128>::smem_bytes()));
  acc(allow_smem(attn_post_kernel<8>, PostSmem<8>::bytes));
  acc(allow_smem(attn_post_kernel<4>, PostSmem<4>::bytes));
  acc(allow_smem(attn_post_kernel<2>, PostSmem<2>::bytes));
  acc(allow_smem(attn_post_kernel<1>, PostSmem<1>::bytes));
  acc(allow_smem(pointnet_kernel<24, 24, 1, 11>, PointNetCfg<11>::smem_bytes));
  acc(allow_smem(pointnet_kernel<11, 12, 3, 19>, PointNetCfg<19>::smem_bytes));
  acc(allow_smem(pointnet_kernel<2, 4, 1, 16>, PointNetCfg<16>::smem_bytes));
  acc(allow_smem(pointnet_kernel<2, 4, 1, 8>, PointNetCfg<8>::smem_bytes));
  acc(allow_smem(attn_kv_tc_kernel, kvtc::SMEM_BYTES));
  acc(allow_smem(policy_head2_kernel<2>, Head2Smem<2>::bytes));
  acc(allow_smem(pointnet_tc_kernel<24, 1, 11>, pntc::Cfg<11>::smem_bytes));
  acc(allow_smem(pointnet_tc_kernel<11, 3, 19>, pntc::Cfg<19>::smem_bytes));
  acc(allow_smem(pointnet_tc_kernel<2, 1, 16>, pntc::Cfg<16>::smem_bytes));
  acc(allow_smem(pointnet_tc_kernel<2, 1, 8>, pntc::Cfg<8>::smem_bytes));
  acc(allow_smem(knn_kernel, 64 * 1024));
  acc(allow_smem(attn_kv2_kernel<4, 8>, Kv2Smem<4, 8>::bytes));
  acc(allow_smem(attn_kv2_kernel<8, 8>, Kv2Smem<8, 8>::bytes));
  acc(allow_smem(attn_dstpre2_kernel<4, 8>, Pre2Smem<4, 8>::bytes));
  acc(allow_smem(attn_dstpre2_kernel<8, 8>, Pre2Smem<8, 8>::bytes));
  acc(allow_smem(attn_post2_kernel<4, 8>, Post2Smem<4, 8>::bytes));
  acc(allow_smem(tcp::attn_post_tc_kernel, tcp::SMEM_BYTES));
  acc(allow_smem(psw::attn_post_sw_kernel<96, 32>, psw::SMEM_BYTES));
  acc(allow_smem(psw::attn_post_sw_kernel<128, 32>, psw::SMEM_BYTES));
  acc(allow_smem(psw::attn_post_sw_kernel<96, 16>, psw::SMEM_BYTES));
  acc(allow_smem(psw::attn_post_sw_kernel<128, 16>, psw::SMEM_BYTES));
  state = e == cudaSuccess ? 1 : (int)e + 1000;
  return e == cudaSuccess ? 0 : (int)e;
}

struct DstScratch {
  float *q, *qhat, *s, *gx;
};

struct StackWs {
  DstScratch set[2];
  float *rbar, *aggv, *x0, *x1, *sk, *pw, *ft, *kv;
  int* counter;
};

constexpr size_t WS_PER_DST = 2 * (size_t)(D + H * D + D + D) + H * D + D + 2 * D;

StackWs carve(float* ws, int n_dst, int max_stride) {
  StackWs w;
  float* p = ws;
  for (int i = 0; i < 2; ++i) {
    w.set[i].q = p;    p += (size_t)n_dst * D;
    w.set[i].qhat = p; p += (size_t)n_dst * H * D;
    w.set[i].s = p;    p += (size_t)n_dst * D;
    w.set[i].gx = p;   p += (size_t)n_dst * D;
  }
  w.rbar = p; p += (size_t)n_dst * H * D;
  w.aggv = p; p += (size_t)n_dst * D;
  w.x0 = p;   p += (size_t)n_dst * D;
  w.x1 = p;   p += (size_t)n_dst * D;
  w.sk = p;   p += (size_t)n_dst * max_stride * 8;
  w.pw = p;   p += (size_t)n_dst * max_stride * 8;
  w.ft = p;   p += (size_t)n_dst * ((max_stride + 31) / 32) * 8;
  w.counter = reinterpret_cast<int*>(p); p += 16;
  w.kv = p;
  return w;
}

int launch_kv(const float* x, int n, const float* w, size_t wstride, int layers, float* kv, size_t kvstride,
              cudaStream_t st) {
  if (n <= 0) return 0;
  const int rpt = pick_rpt(n * layers);
  LaunchScope ls(PROSIM_K_ATTN_KV, st);
  const int rt = pick_rt(n);
  if (g_tc_mask & 2) {   // tcgen05 / TMEM 3xTF32 (kv_tc.cuh) for every launch size: a row's K'|V' never depends on the batch it is in
    if (int e = setup_attributes()) return e;
    attn_kv_tc_kernel<<<dim3((n + 127) / 128, layers, 2), 128, kvtc::SMEM_BYTES, st>>>(x, n, w, wstride, kv, kvstride);
    PROSIM_CHECK_LAUNCH();
    return 0;
  }
  if (rt == 4) {
    attn_kv2_kernel<8, 8><<<dim3((n + 63) / 64, layers), 256, Kv2Smem<8, 8>::bytes, st>>>(x, n, w, wstride, kv, kvstride);
    PROSIM_CHECK_LAUNCH();
    return 0;
  }
  if (rt == 2) {
    attn_kv2_kernel<4, 8><<<dim3((n + 31) / 32, layers), 256, Kv2Smem<4, 8>::bytes, st>>>(x, n, w, wstride, kv, kvstride);
    PROSIM_CHECK_LAUNCH();
    return 0;
  }
  DISPATCH_RPT(rpt, attn_kv_kernel<RPT><<<dim3((n + 2 * RPT - 1) / (2 * RPT), layers), 256, 0, st>>>(x, n, w, wstride, kv, kvstride));
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int launch_dstpre(const float* x, int n, const float* w, const DstScratch& d, cudaStream_t st) {
  if (n <= 0) return 0;
  const int rpt = pick_rpt(n);
  LaunchScope ls(PROSIM_K_ATTN_DSTPRE, st);
  if ((g_tc_mask & 1) && (g_tc_mask & 8)) {
    // the first layer of a stack has no previous node kernel to carry its destination-side projections: the 32-row tcgen05
    // kernel in its "pre-only" mode runs just that tail (16 of its 58 weight chunks) -- the same arithmetic every later
    // layer of the stack gets, for any launch size (the FFMA kernels below: 127 us per 16 k-row launch, 39 us for 128 rows)
    psw::Args a{};
    a.x = x; a.q_n = d.q; a.qhat_n = d.qhat; a.s_n = d.s; a.gx_n = d.gx;
    a.W = nullptr; a.Wn = w; a.n = n;
    cudaError_t le;
    if (sw_rows16(n)) le = launch_pdl(psw::attn_post_sw_kernel<96, 16>, dim3((n + 15) / 16), dim3(psw::THREADS), psw::SMEM_BYTES, st, a);
    else le = launch_pdl(psw::attn_post_sw_kernel<96, 32>, dim3((n + 31) / 32), dim3(psw::THREADS), psw::SMEM_BYTES, st, a);
    if (le != cudaSuccess) return (int)le;
    PROSIM_CHECK_LAUNCH();
    return 0;
  }
  const int rt = pick_rt(n);
  if (rt == 4) {
    attn_dstpre2_kernel<8, 8><<<(n + 63) / 64, 256, Pre2Smem<8, 8>::bytes, st>>>(x, n, w, d.q, d.qhat, d.s, d.gx);
    PROSIM_CHECK_LAUNCH();
    return 0;
  }
  if (rt == 2) {
    attn_dstpre2_kernel<4, 8><<<(n + 31) / 32, 256, Pre2Smem<4, 8>::bytes, st>>>(x, n, w, d.q, d.qhat, d.s, d.gx);
    PROSIM_CHECK_LAUNCH();
    return 0;
  }
  DISPATCH_RPT(rpt, attn_dstpre_kernel<RPT><<<(n + 2 * RPT - 1) / (2 * RPT), 256, 0, st>>>(x, n, w, d.q, d.qhat, d.s, d.gx));
  PROSIM_CHECK_LAUNCH();
  return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// a row-major fp32 matrix [rows][cols] as a 2-D tensor fetched in boxes of box_rows x 32 floats with the 128-byte swizzle
int make_map2d(CUtensorMap* tm, const float* base, size_t rows, int cols, int box_rows) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = tensor_map_encoder();
  if (!enc) return ERR_ARG - 10;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : ERR_ARG - 11;
}
// z of a graph: [n_dst * stride][zd], boxes of 8 edges (edge4.cuh)
int make_z_map(CUtensorMap* tm, const float* z, size_t rows, int zd) { return make_map2d(tm, z, rows, zd, 8); }

// Tensor maps of the node kernel's [n][cols] row buffers (128-row boxes).  The buffers are a handful of workspace
// pointers that recur every layer and every forward, so encoded maps are kept in a small table.
struct MapEntry {
  const float* base;
  size_t rows;
  int cols;
  CUtensorMap tm;
};
int cached_map128(CUtensorMap* out, const float* base, size_t rows, int cols) {
  constexpr int CAP = 64;
  static thread_local MapEntry table[CAP];   // per host thread: the entry points stay re-entrant
  static thread_local int used = 0, next = 0;
  for (int i = 0; i < used; ++i)
    if (table[i].base == base && table[i].rows == rows && table[i].cols == cols) {
      *out = table[i].tm;
      return 0;
    }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || rows == 0 || rows > 0x7fffffffull) return ERR_ARG;
  MapEntry& e = table[next];
  if (int err = make_map2d(&e.tm, base, rows, cols, 128)) return err;
  e.base = base; e.rows = rows; e.cols = cols;
  *out = e.tm;
  next = (next + 1) % CAP;
  if (used < CAP) ++used;
  return 0;
}

template <int ZD, int NW>
int launch_edge4(const CUtensorMap& tm, const CUtensorMap& tm32, const DstScratch& d, const prosim_graph_t& g, int n_dst, float* rbar, float* sk,
                 float* pw, float* ft, int ft_tiles, int* counter, cudaStream_t st) {
  const int grid = (n_dst + NW - 1) / NW < 148 ? (n_dst + NW - 1) / NW : 148;   // one persistent CTA per SM
  attn_edge4_kernel<ZD, NW><<<grid, NW * 32, Edge4Cfg<ZD>::smem_bytes(NW), st>>>(tm, tm32, d.qhat, sk, g.deg, g.stride, n_dst, rbar,
                                                                              pw, ft, ft_tiles, counter);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int launch_edge(const DstScratch& d, const float* kv, const prosim_graph_t& g, int n_dst, float* rbar, float* aggv,
                float* sk, float* pw, float* ft, int* counter, cudaStream_t st) {
  if (n_dst <= 0) return 0;
  if ((g.zd != 96 && g.zd != 128) || g.stride > 32 * Edge4Cfg<96>::MT_TILES) return ERR_ARG;
  const size_t z_rows = (size_t)n_dst * g.stride;
  if (z_rows > 0x7fffffffull || (reinterpret_cast<uintptr_t>(g.z) & 15) != 0) return ERR_ARG;
  alignas(64) CUtensorMap tm, tm32;
  if (int e = make_z_map(&tm, g.z, z_rows, g.zd)) return e;
  if (int e = make_map2d(&tm32, g.z, z_rows, g.zd, 32)) return e;
  const int ft_tiles = (g.stride + 31) / 32;
  if (n_dst <= FUSED_EDGE_MAX_ROWS && (g_tc_mask & 16)) {
    // small launch: q.K' scores, the z pass and the V' aggregation in ONE launch, one CTA of four warps per row (edge_row.cuh)
    LaunchScope ls(PROSIM_K_ATTN_EDGE, st);
    cudaError_t le;
    if (g.zd == 96)
      le = launch_pdl(attn_edge_row_kernel<96>, dim3(n_dst), dim3(128), EdgeRowCfg<96>::smem_bytes(), st, tm, tm32, d.qhat, d.q, kv, g.nbr,
                      g.deg, g.stride, n_dst, sk, rbar, pw, ft, ft_tiles, aggv);
    else
      le = launch_pdl(attn_edge_row_kernel<128>, dim3(n_dst), dim3(128), EdgeRowCfg<128>::smem_bytes(), st, tm, tm32, d.qhat, d.q, kv,
                      g.nbr, g.deg, g.stride, n_dst, sk, rbar, pw, ft, ft_tiles, aggv);
    if (le != cudaSuccess) return (int)le;
    PROSIM_CHECK_LAUNCH();
    return 0;
  }
  {
    LaunchScope ls(PROSIM_K_EDGE_QK, st);
    edge_qk_kernel<<<(n_dst + 7) / 8, 256, 0, st>>>(d.q, kv, g.nbr, g.deg, g.stride, n_dst, sk, counter);
    PROSIM_CHECK_LAUNCH();
  }
  {
    // one warp per destination row; fewer warps per CTA when the launch has fewer rows than the chip has warp slots
    LaunchScope ls(PROSIM_K_ATTN_EDGE, st);
    const int per_sm = (n_dst + 147) / 148;
    int e = ERR_ARG;
    if (g.zd == 96) {
      if (per_sm <= 1) e = launch_edge4<96, 1>(tm, tm32, d, g, n_dst, rbar, sk, pw, ft, ft_tiles, counter, st);
      else if (per_sm <= 2) e = launch_edge4<96, 2>(tm, tm32, d, g, n_dst, rbar, sk, pw, ft, ft_tiles, counter, st);
      else if (per_sm <= 4) e = launch_edge4<96, 4>(tm, tm32, d, g, n_dst, rbar, sk, pw, ft, ft_tiles, counter, st);
      else if (per_sm <= 8) e = launch_edge4<96, 8>(tm, tm32, d, g, n_dst, rbar, sk, pw, ft, ft_tiles, counter, st);
      else e = launch_edge4<96, 12>(tm, tm32, d, g, n_dst, rbar, sk, pw, ft, ft_tiles, counter, st);
    } else {
      if (per_sm <= 1) e = launch_edge4<128, 1>(tm, tm32, d, g, n_dst, rbar, sk, pw, ft, ft_tiles, counter, st);
      else if (per_sm <= 2) e = launch_edge4<128, 2>(tm, tm32, d, g, n_dst, rbar, sk, pw, ft, ft_tiles, counter, st);
      else if (per_sm <= 4) e = launch_edge4<128, 4>(tm, tm32, d, g, n_dst, rbar, sk, pw, ft, ft_tiles, counter, st);
      else if (per_sm <= 8) e = launch_edge4<128, 8>(tm, tm32, d, g, n_dst, rbar, sk, pw, ft, ft_tiles, counter, st);
      else e = launch_edge4<128, 10>(tm, tm32, d, g, n_dst, rbar, sk, pw, ft, ft_tiles, counter, st);
    }
    if (e) return e;
  }
  LaunchScope ls(PROSIM_K_EDGE_AV, st);
  edge_av_kernel<<<(n_dst + 7) / 8, 256, 0, st>>>(pw, ft, ft_tiles, kv, g.nbr, g.deg, g.stride, n_dst, aggv);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int launch_post(const float* x, int n, int zd, const float* rbar, const float* aggv, const DstScratch& cur, const float* w,
                float* out, const float* w_next, const DstScratch& nxt, cudaStream_t st) {
  if (n <= 0) return 0;
  const int rpt = pick_rpt(n);
  LaunchScope ls(PROSIM_K_ATTN_POST, st);
  // The 32-row tcgen05 kernel serves every launch size up to two waves of CTAs: for a single 128-agent scene (4 CTAs) its
  // per-CTA chain (~40 us) is less than half of what the 2-rows-per-CTA FFMA kernel needs to stream 1.1 MB of weights per CTA
  // (94 us measured per launch, 58 % of a single-scene forward).  Rows never interact, so results do not depend on the size.
  if ((g_tc_mask & 1) && (g_tc_mask & 8) && n <= SW_MAX_ROWS && (zd == 96 || zd == 128)) {
    LaunchScope ls_sw(PROSIM_K_ATTN_POST_SW, st);   // counted (and timed) under both classes
    psw::Args a;
    a.x = x; a.rbar = rbar; a.aggv = aggv; a.s = cur.s; a.gx = cur.gx; a.out = out;
    a.q_n = nxt.q; a.qhat_n = nxt.qhat; a.s_n = nxt.s; a.gx_n = nxt.gx;
    a.W = w; a.Wn = w_next; a.n = n;
    cudaError_t le;
    if (sw_rows16(n)) {
      const dim3 grid((n + 15) / 16);
      if (zd == 96) le = launch_pdl(psw::attn_post_sw_kernel<96, 16>, grid, dim3(psw::THREADS), psw::SMEM_BYTES, st, a);
      else le = launch_pdl(psw::attn_post_sw_kernel<128, 16>, grid, dim3(psw::THREADS), psw::SMEM_BYTES, st, a);
    } else {
      const dim3 grid((n + 31) / 32);
      if (zd == 96) le = launch_pdl(psw::attn_post_sw_kernel<96, 32>, grid, dim3(psw::THREADS), psw::SMEM_BYTES, st, a);
      else le = launch_pdl(psw::attn_post_sw_kernel<128, 32>, grid, dim3(psw::THREADS), psw::SMEM_BYTES, st, a);
    }
    if (le != cudaSuccess) return (int)le;
    PROSIM_CHECK_LAUNCH();
    return 0;
  }
  if (pick_rt(n) != 0 && (g_tc_mask & 1)) {
    alignas(64) tcp::Maps m;
    const bool nx = w_next != nullptr;
    int e = 0;
    auto mk = [&](CUtensorMap* tm, const float* p, int cols) { if (!e) e = cached_map128(tm, p, (size_t)n, cols); };
    mk(&m.rbar, rbar, H * zd); mk(&m.aggv, aggv, D); mk(&m.s, cur.s, D); mk(&m.gx, cur.gx, D); mk(&m.x, x, D); mk(&m.out, out, D);
    mk(&m.q_n, nx ? nxt.q : out, D); mk(&m.s_n, nx ? nxt.s : out, D); mk(&m.gx_n, nx ? nxt.gx : out, D);
    mk(&m.qhat_n, nx ? nxt.qhat : out, nx ? H * D : D);
    if (e) return e;
    tcp::attn_post_tc_kernel<<<(n + 127) / 128, tcp::THREADS, tcp::SMEM_BYTES, st>>>(m, n, zd, w, w_next);
    PROSIM_CHECK_LAUNCH();
    return 0;
  }
  if (pick_rt(n) != 0 && zd == 96) {
    attn_post2_kernel<4, 8><<<(n + 31) / 32, 256, Post2Smem<4, 8>::bytes, st>>>(x, n, zd, rbar, aggv, cur.s, cur.gx, w, out, w_next,
                                                                         nxt.q, nxt.qhat, nxt.s, nxt.gx);
    PROSIM_CHECK_LAUNCH();
    return 0;
  }
  DISPATCH_RPT(rpt, attn_post_kernel<RPT><<<(n + 2 * RPT - 1) / (2 * RPT), 256, PostSmem<RPT>::bytes, st>>>(
                        x, n, zd, rbar, aggv, cur.s, cur.gx, w, out, w_next, nxt.q, nxt.qhat, nxt.s, nxt.gx));
  PROSIM_CHECK_LAUNCH();
  return 0;
}

}  // namespace

// test support (prosim_debug_scrub): one CTA per SM (227 KB of dynamic shared memory), 128 threads = 128 TMEM lanes
__global__ void __launch_bounds__(128, 1) scrub_kernel(float pattern, int smem_floats) {
  extern __shared__ float scrub_smem[];
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < smem_floats; i += 128) scrub_smem[i] = pattern;
  if (threadIdx.x < 32) tc::tmem_alloc(&tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = tmem_slot + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = pattern;
  for (int c = 0; c < 512; c += 32) tcp::tmem_st32(taddr + c, v);
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_slot, 512);
  // keep the CTA alive long enough for the other SMs to receive theirs (one CTA per SM by shared-memory size)
  const long long t0 = clock64();
  while (clock64() - t0 < 200000) {}
  if (scrub_smem[smem_floats - 1 - threadIdx.x] != pattern && pattern == pattern) tmem_slot = 0;   // keep the stores
}

extern "C" {

int prosim_abi_version(void) { return 9; }
int prosim_debug_scrub(float pattern, void* stream) {
  constexpr int BYTES = 226 * 1024;
  cudaError_t e = cudaFuncSetAttribute(scrub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int rep = 0; rep < 2; ++rep) scrub_kernel<<<sms, 128, BYTES, static_cast<cudaStream_t>(stream)>>>(pattern, BYTES / 4);
  PROSIM_CHECK_LAUNCH();
  return 0;
}
int prosim_tc_debug_read(long long* out32) {
  if (!out32) return ERR_ARG;
  return (int)cudaMemcpyFromSymbol(out32, tcp::g_tcp_dbg, 32 * sizeof(long long));
}
int prosim_set_stack_split(int parts) {
  if (parts < 1 || parts > 4) return ERR_ARG;
  g_split = parts;
  return 0;
}
int prosim_set_tensor_core(int on) {
  g_use_tc = on != 0;
  g_tc_mask = on == 1 ? 31 : (on & 31);  // 1 = everything (the default); other values select kernels by bit
  return 0;
}

long long prosim_launch_count(int kernel_class) {
  if (kernel_class >= 0 && kernel_class < N_CLASSES) return g_launches[kernel_class];
  long long t = 0;
  for (int i = 0; i < N_CLASSES; ++i)
    if (i != PROSIM_K_ATTN_POST_SW) t += g_launches[i];   // a sub-class of PROSIM_K_ATTN_POST: not a launch of its own
  return t;
}

int prosim_profile_enable(int kernel_class) {
  if (kernel_class >= N_CLASSES) return ERR_ARG;
  if (kernel_class >= 0 && !g_ev_made) {
    for (int i = 0; i < MAX_REC; ++i) {
      if (cudaEventCreate(&g_ev0[i]) != cudaSuccess || cudaEventCreate(&g_ev1[i]) != cudaSuccess) return (int)cudaGetLastError();
    }
    g_ev_made = true;
  }
  g_prof_class = kernel_class;
  g_prof_n = 0;
  return 0;
}

int prosim_profile_read(double* total_ms, int* count) {
  if (!total_ms || !count) return ERR_ARG;
  double t = 0.0;
  for (int i = 0; i < g_prof_n; ++i) {
    if (cudaEventSynchronize(g_ev1[i]) != cudaSuccess) return (int)cudaGetLastError();
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_ev0[i], g_ev1[i]) != cudaSuccess) return (int)cudaGetLastError();
    t += ms;
  }
  *total_ms = t;
  *count = g_prof_n;
  g_prof_n = 0;
  return 0;
}
int prosim_attn_layer_floats(void) { return aw::SIZE; }
int prosim_pointnet_floats(void) { return pw::SIZE; }
int prosim_head_floats(void) { return hw::SIZE; }
int prosim_mlp2_floats(void) { return mw::SIZE; }
size_t prosim_attn_workspace_floats(int n_dst, int n_src, int max_stride) {
  const size_t nd = n_dst < 0 ? 0 : n_dst, ns = n_src < 0 ? 0 : n_src, st = max_stride < 1 ? 1 : max_stride;
  return (WS_PER_DST + 16 * st + 8 * ((st + 31) / 32)) * nd + 256 * ns + 64;
}

int prosim_pointnet_fwd(int kind, const float* x, const uint8_t* mask, const int32_t* rows, int n_poly, const float* w,
                        const float* w_tc, float* out, prosim_stream_t stream) {
  if (n_poly < 0 || kind < 0 || kind > 3) return ERR_ARG;
  if (n_poly == 0) return 0;
  if (!x || (!mask && kind < 2) || !rows || !w || !out) return ERR_ARG;
  if (int e = setup_attributes()) return e;
  LaunchScope ls(PROSIM_K_POINTNET, S(stream));
  if (w_tc != nullptr && (g_tc_mask & 4)) {   // tcgen05 / TMEM 3xTF32 kernel (pointnet_tc.cuh)
    if ((reinterpret_cast<uintptr_t>(w_tc) & 15) != 0) return ERR_ARG;
    if (kind == 0)
      pointnet_tc_kernel<24, 1, 11><<<(n_poly + pntc::Cfg<11>::G - 1) / pntc::Cfg<11>::G, 128, pntc::Cfg<11>::smem_bytes, S(stream)>>>(
          x, mask, 24, rows, n_poly, w, w_tc, out);
    else if (kind == 1)
      pointnet_tc_kernel<11, 3, 19><<<(n_poly + pntc::Cfg<19>::G - 1) / pntc::Cfg<19>::G, 128, pntc::Cfg<19>::smem_bytes, S(stream)>>>(
          x, mask, 1, rows, n_poly, w, w_tc, out);
    else if (kind == 2)
      pointnet_tc_kernel<2, 1, 16><<<(n_poly + pntc::Cfg<16>::G - 1) / pntc::Cfg<16>::G, 128, pntc::Cfg<16>::smem_bytes, S(stream)>>>(
          x, mask, 1, rows, n_poly, w, w_tc, out);
    else
      pointnet_tc_kernel<2, 1, 8><<<(n_poly + pntc::Cfg<8>::G - 1) / pntc::Cfg<8>::G, 128, pntc::Cfg<8>::smem_bytes, S(stream)>>>(
          x, mask, 1, rows, n_poly, w, w_tc, out);
    PROSIM_CHECK_LAUNCH();
    return 0;
  }
  if (kind == 0) {
    constexpr int G = PointNetCfg<11>::G;
    pointnet_kernel<24, 24, 1, 11><<<(n_poly + G - 1) / G, 256, PointNetCfg<11>::smem_bytes, S(stream)>>>(
        x, mask, 24, rows, n_poly, w, out);
  } else if (kind == 1) {
    constexpr int G = PointNetCfg<19>::G;
    pointnet_kernel<11, 12, 3, 19><<<(n_poly + G - 1) / G, 256, PointNetCfg<19>::smem_bytes, S(stream)>>>(
        x, mask, 1, rows, n_poly, w, out);
  } else if (kind == 2) {
    constexpr int G = PointNetCfg<16>::G;
    pointnet_kernel<2, 4, 1, 16><<<(n_poly + G - 1) / G, 256, PointNetCfg<16>::smem_bytes, S(stream)>>>(
        x, mask, 1, rows, n_poly, w, out);
  } else {
    constexpr int G = PointNetCfg<8>::G;
    pointnet_kernel<2, 4, 1, 8><<<(n_poly + G - 1) / G, 256, PointNetCfg<8>::smem_bytes, S(stream)>>>(
        x, mask, 1, rows, n_poly, w, out);
  }
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_build_radius_edges(const float* qpos, const int32_t* qscene, int n_q, const float* spos, const int32_t* seg,
                              float r, int cap, int drop_self, int32_t* nbr, int32_t* deg, int stride,
                              prosim_stream_t stream) {
  if (n_q < 0 || cap < 0 || stride <= 0) return ERR_ARG;
  if (n_q == 0) return 0;
  if (!qpos || !qscene || !spos || !seg || !nbr || !deg) return ERR_ARG;
  const float r2 = r * r;
  LaunchScope ls(PROSIM_K_RADIUS, S(stream));
  radius_kernel<<<(n_q + 7) / 8, 256, 0, S(stream)>>>(reinterpret_cast<const float2*>(qpos), qscene, n_q,
                                                      reinterpret_cast<const float2*>(spos),
                                                      reinterpret_cast<const int4*>(seg), r2, cap, drop_self, nbr, deg,
                                                      stride);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_build_knn_edges(const float* qpos, const int32_t* qscene, int n_q, const float* spos, const int32_t* seg,
                           int k, int nmax, int32_t* nbr, int32_t* deg, int stride, prosim_stream_t stream) {
  if (n_q < 0 || k <= 0 || nmax <= 0 || stride <= 0) return ERR_ARG;
  if (n_q == 0) return 0;
  if (!qpos || !qscene || !spos || !seg || !nbr || !deg) return ERR_ARG;
  if (int e = setup_attributes()) return e;
  const size_t smem = (size_t)nmax * 8 * sizeof(unsigned);   // 8 query warps x nmax keys
  if (smem > 64 * 1024) return ERR_ARG;
  LaunchScope ls(PROSIM_K_KNN, S(stream));
  knn_kernel<<<(n_q + 7) / 8, 256, smem, S(stream)>>>(reinterpret_cast<const float2*>(qpos), qscene, n_q,
                                                      reinterpret_cast<const float2*>(spos), reinterpret_cast<const int4*>(seg), k,
                                                      nmax, nbr, deg, stride);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_edge_pe(const float* dpos, const float* dori, int n_dst, const float* spos, const float* sori,
                   const int32_t* nbr, const int32_t* deg, int stride, const float* dim_t16, const float* extra, int zd,
                   float* z, prosim_stream_t stream) {
  if (n_dst < 0 || stride <= 0 || (zd != 96 && zd != 128) || (extra && zd != 128)) return ERR_ARG;
  if (n_dst == 0) return 0;
  if (!dpos || !dori || !spos || !sori || !nbr || !deg || !dim_t16 || !z) return ERR_ARG;
  LaunchScope ls(PROSIM_K_EDGE_PE, S(stream));
  if (zd == 96)
    edge_pe_kernel<96><<<n_dst, 128, 0, S(stream)>>>(reinterpret_cast<const float2*>(dpos), dori,
                                                     reinterpret_cast<const float2*>(spos), sori, nbr, deg, stride,
                                                     dim_t16, extra, z);
  else
    edge_pe_kernel<128><<<n_dst, 128, 0, S(stream)>>>(reinterpret_cast<const float2*>(dpos), dori,
                                                      reinterpret_cast<const float2*>(spos), sori, nbr, deg, stride,
                                                      dim_t16, extra, z);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_attn_kv(const float* x_src, int n_src, const float* w, size_t w_layer_stride, int n_layers, float* kv,
                   size_t kv_layer_stride, prosim_stream_t stream) {
  if (n_src < 0 || n_layers <= 0) return ERR_ARG;
  if (n_src == 0) return 0;
  if (!x_src || !w || !kv) return ERR_ARG;
  return launch_kv(x_src, n_src, w, w_layer_stride, n_layers, kv, kv_layer_stride, S(stream));
}

int prosim_attn_layer_fwd(const float* x_src, int n_src, const float* x_dst, int n_dst, const prosim_graph_t* g,
                          const float* w, float* workspace, size_t workspace_floats, float* out,
                          prosim_stream_t stream) {
  if (n_src < 0 || n_dst < 0 || !g) return ERR_ARG;
  if (n_dst == 0) return 0;
  if (!x_src || !x_dst || !w || !workspace || !out || !g->z || !g->nbr || !g->deg) return ERR_ARG;
  if (workspace_floats < prosim_attn_workspace_floats(n_dst, n_src, g->stride)) return ERR_WORKSPACE;
  if (int e = setup_attributes()) return e;
  cudaStream_t st = S(stream);
  StackWs ws = carve(workspace, n_dst, g->stride);
  if (int e = launch_kv(x_src, n_src, w, 0, 1, ws.kv, 0, st)) return e;
  if (int e = launch_dstpre(x_dst, n_dst, w, ws.set[0], st)) return e;
  if (int e = launch_edge(ws.set[0], ws.kv, *g, n_dst, ws.rbar, ws.aggv, ws.sk, ws.pw, ws.ft, ws.counter, st)) return e;
  return launch_post(x_dst, n_dst, g->zd, ws.rbar, ws.aggv, ws.set[0], w, out, nullptr, ws.set[1], st);
}

// One chain of n_layers x (side A [, side B]) over destination rows [r0, r0 + n): every array of the workspace, of the
// graphs and of x / out is row indexed, so a row range is just an offset.
static int run_stack_rows(const float* x, int r0, int n, int n_layers, const prosim_stack_side_t* side_a,
                          const prosim_stack_side_t* side_b, const StackWs& ws_all, float* out, int part, cudaStream_t st) {
  auto rows = [&](float* p, size_t per_row) { return p + (size_t)r0 * per_row; };
  const int max_stride = side_b && side_b->graph.stride > side_a->graph.stride ? side_b->graph.stride : side_a->graph.stride;
  StackWs ws = ws_all;
  for (int i = 0; i < 2; ++i) {
    ws.set[i].q = rows(ws_all.set[i].q, D);
    ws.set[i].qhat = rows(ws_all.set[i].qhat, H * D);
    ws.set[i].s = rows(ws_all.set[i].s, D);
    ws.set[i].gx = rows(ws_all.set[i].gx, D);
  }
  ws.rbar = rows(ws_all.rbar, H * D);
  ws.aggv = rows(ws_all.aggv, D);
  ws.x0 = rows(ws_all.x0, D);
  ws.x1 = rows(ws_all.x1, D);
  ws.sk = rows(ws_all.sk, (size_t)max_stride * 8);
  ws.pw = rows(ws_all.pw, (size_t)max_stride * 8);
  ws.ft = rows(ws_all.ft, (size_t)((max_stride + 31) / 32) * 8);
  ws.counter = ws_all.counter + part;
  const prosim_stack_side_t* sides[2] = {side_a, side_b};
  prosim_graph_t g[2];
  const int n_sides = side_b ? 2 : 1;
  for (int i = 0; i < n_sides; ++i) {
    g[i] = sides[i]->graph;
    g[i].z += (size_t)r0 * g[i].stride * g[i].zd;
    g[i].nbr += (size_t)r0 * g[i].stride;
    g[i].deg += r0;
  }
  const int total = n_layers * n_sides;
  // x buffers: input -> x0 -> x1 -> x0 ... ; the last layer writes `out`
  const float* cur_x = x + (size_t)r0 * D;
  float* out_rows = out + (size_t)r0 * D;
  int cur_set = 0;
  if (int e = launch_dstpre(cur_x, n, side_a->w, ws.set[cur_set], st)) return e;
  for (int i = 0; i < total; ++i) {
    const int l = i / n_sides;
    const prosim_stack_side_t* sd = sides[i % n_sides];
    const float* w = sd->w + (size_t)l * aw::SIZE;
    const float* kv;
    if (sd->kv) {
      kv = sd->kv + (size_t)l * sd->kv_layer_stride;
    } else {   // self-source layer (never split: r0 == 0 and n == all rows)
      if (int e = launch_kv(cur_x, n, w, 0, 1, ws.kv, 0, st)) return e;
      kv = ws.kv;
    }
    if (int e = launch_edge(ws.set[cur_set], kv, g[i % n_sides], n, ws.rbar, ws.aggv, ws.sk, ws.pw, ws.ft, ws.counter, st)) return e;
    const bool last = i == total - 1;
    const float* w_next = nullptr;
    if (!last) {
      const int ni = i + 1;
      w_next = sides[ni % n_sides]->w + (size_t)(ni / n_sides) * aw::SIZE;
    }
    float* dst_x = last ? out_rows : (cur_x == ws.x0 ? ws.x1 : ws.x0);
    if (int e = launch_post(cur_x, n, sd->graph.zd, ws.rbar, ws.aggv, ws.set[cur_set], w, dst_x, w_next, ws.set[cur_set ^ 1], st))
      return e;
    cur_x = dst_x;
    cur_set ^= 1;
  }
  return 0;
}

// Side streams for the row-split schedule (one set per device, created on first use)
struct SplitStreams {
  cudaStream_t st[3];
  cudaEvent_t fork, join[3];
  bool ok;
};
static SplitStreams* split_streams() {
  static SplitStreams per_dev[16];
  static bool made[16] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SplitStreams& s = per_dev[dev];
  if (!made[dev]) {
    made[dev] = true;
    s.ok = cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 3 && s.ok; ++i)
      s.ok = cudaStreamCreateWithFlags(&s.st[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&s.join[i], cudaEventDisableTiming) == cudaSuccess;
  }
  return s.ok ? &s : nullptr;
}

int prosim_attn_stack_fwd(const float* x, int n_dst, int n_layers, const prosim_stack_side_t* side_a,
                          const prosim_stack_side_t* side_b, float* workspace, size_t workspace_floats, float* out,
                          prosim_stream_t stream) {
  if (n_dst < 0 || n_layers <= 0 || !side_a) return ERR_ARG;
  if (n_dst == 0) return 0;
  if (!x || !workspace || !out || !side_a->w) return ERR_ARG;
  const bool self_src = side_a->kv == nullptr || (side_b && side_b->kv == nullptr);
  const int max_stride = side_b && side_b->graph.stride > side_a->graph.stride ? side_b->graph.stride : side_a->graph.stride;
  if (workspace_floats < prosim_attn_workspace_floats(n_dst, self_src ? n_dst : 0, max_stride)) return ERR_WORKSPACE;
  if (int e = setup_attributes()) return e;
  cudaStream_t st = S(stream);
  StackWs ws = carve(workspace, n_dst, max_stride);
  // Row-split schedule: with fixed sources (K'|V' precomputed) the destination rows never interact, so the stack is
  // run as up to g_split independent chains of >= 1024 rows (multiples of 128: identical tiles, bit-identical
  // results) on side streams.  The tcgen05 node kernel occupies one SM per 128 rows -- alone it leaves 3/4 of the chip
  // idle -- and the chains fill each other's gaps (edge kernels take their rows from a dynamic queue).
  int parts = 1;
  if (!self_src && g_split > 1) {
    parts = n_dst / 1024;
    if (parts > g_split) parts = g_split;
    if (parts < 1) parts = 1;
  }
  SplitStreams* ss = parts > 1 ? split_streams() : nullptr;
  if (!ss) return run_stack_rows(x, 0, n_dst, n_layers, side_a, side_b, ws, out, 0, st);
  const int per = ((n_dst + parts - 1) / parts + 127) / 128 * 128;
  if (cudaEventRecord(ss->fork, st) != cudaSuccess) return (int)cudaGetLastError();
  int err = 0;
  for (int p = 0; p < parts && !err; ++p) {
    const int r0 = p * per;
    const int n = r0 + per <= n_dst ? per : n_dst - r0;
    if (n <= 0) break;
    cudaStream_t sp = p == 0 ? st : ss->st[p - 1];
    if (p > 0 && cudaStreamWaitEvent(sp, ss->fork, 0) != cudaSuccess) err = (int)cudaGetLastError();
    if (!err) err = run_stack_rows(x, r0, n, n_layers, side_a, side_b, ws, out, p * 4, sp);
    if (p > 0 && !err) {
      if (cudaEventRecord(ss->join[p - 1], sp) != cudaSuccess || cudaStreamWaitEvent(st, ss->join[p - 1], 0) != cudaSuccess)
        err = (int)cudaGetLastError();
    }
  }
  return err;
}

int prosim_policy_head_fwd(const float* feat, const int32_t* agent_type, int P, const float* w, const float* noise,
                           float noise_std, float* motion_pred, prosim_stream_t stream) {
  if (P < 0 || (noise && !(noise_std > 0.f))) return ERR_ARG;
  if (P == 0) return 0;
  if (!feat || !agent_type || !w || !motion_pred) return ERR_ARG;
  const int rpt = pick_rpt(P);
  LaunchScope ls(PROSIM_K_HEAD, S(stream));
  if (pick_rt(P) != 0) {   // >= 1024 rows: weights through the shared-memory stream (bit-identical to the kernel below)
    if (int e = setup_attributes()) return e;
    policy_head2_kernel<2><<<(P + 15) / 16, 256, Head2Smem<2>::bytes, S(stream)>>>(feat, agent_type, P, w, noise, noise_std,
                                                                             motion_pred);
    PROSIM_CHECK_LAUNCH();
    return 0;
  }
  DISPATCH_RPT(rpt, policy_head_kernel<RPT><<<(P + 2 * RPT - 1) / (2 * RPT), 256, 0, S(stream)>>>(feat, agent_type, P, w, noise, noise_std, motion_pred));
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_reconst_fwd(const float* emd, int P, const float* w, float* out, prosim_stream_t stream) {
  if (P < 0) return ERR_ARG;
  if (P == 0) return 0;
  if (!emd || !w || !out) return ERR_ARG;
  const int rpt = pick_rpt(P);
  LaunchScope ls(PROSIM_K_HEAD, S(stream));
  DISPATCH_RPT(rpt, reconst_kernel<RPT><<<(P + 2 * RPT - 1) / (2 * RPT), 256, 0, S(stream)>>>(emd, P, w, out));
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_mlp2_fwd(const float* in, int ld_in, int k0, int n, int use_ln, const float* w, const float* tpe_t, int tpe_ld,
                    const float* dim_t128, float* out, prosim_stream_t stream) {
  if (n < 0 || k0 <= 0 || k0 > 8 || ld_in < k0) return ERR_ARG;
  if (n == 0) return 0;
  if (!in || !w || !out || (tpe_t && !dim_t128)) return ERR_ARG;
  const int rpt = pick_rpt(n);
  LaunchScope ls(PROSIM_K_MLP2, S(stream));
  DISPATCH_RPT(rpt, mlp2_kernel<RPT><<<(n + 2 * RPT - 1) / (2 * RPT), 256, 0, S(stream)>>>(in, ld_in, k0, n, use_ln, w, tpe_t, tpe_ld, dim_t128, out));
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_obs_fuse_floats(void) { return fw::SIZE; }
int prosim_obs_fuse_fwd(const float* x_old, const int32_t* idx_old, float* x_new, const int32_t* idx_new, int n, const float* w,
                        prosim_stream_t stream) {
  if (n < 0) return ERR_ARG;
  if (n == 0) return 0;
  if (!x_old || !idx_old || !x_new || !idx_new || !w) return ERR_ARG;
  const int rpt = pick_rpt(n);
  LaunchScope ls(PROSIM_K_MLP2, S(stream));
  DISPATCH_RPT(rpt, obs_fuse_kernel<RPT><<<(n + 2 * RPT - 1) / (2 * RPT), 256, 0, S(stream)>>>(x_old, idx_old, x_new, idx_new, n, w));
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_tag_embed_fwd(const int64_t* tags, int n, int n_tags, const float* table, const float* dim_t64, float* out,
                         prosim_stream_t stream) {
  if (n < 0 || n_tags < 0 || n_tags > 16) return ERR_ARG;
  if (n == 0) return 0;
  if (!tags || !table || !dim_t64 || !out) return ERR_ARG;
  LaunchScope ls(PROSIM_K_MLP2, S(stream));
  tag_embed_kernel<<<(n * D + 255) / 256, 256, 0, S(stream)>>>(reinterpret_cast<const long long*>(tags), n, n_tags, table,
                                                               dim_t64, out);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_cond_pool_fwd(const float* emb, const int32_t* slot, int P, int n_slots, float* extra, int32_t* has,
                         prosim_stream_t stream) {
  if (P < 0 || n_slots <= 0) return ERR_ARG;
  if (P == 0) return 0;
  if (!emb || !slot || !extra || !has) return ERR_ARG;
  LaunchScope ls(PROSIM_K_MLP2, S(stream));
  cond_pool_kernel<<<(P * D + 255) / 256, 256, 0, S(stream)>>>(emb, slot, P, n_slots, extra, has);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_init_traj(const float* obs_in, const float* obs_pos, const float* obs_head, const int32_t* p_slot,
                     const int32_t* p_row, int P, int T, float* traj, float* vel, float* init_pos, float* init_heading,
                     prosim_stream_t stream) {
  if (P < 0 || T < HIST) return ERR_ARG;
  if (P == 0) return 0;
  if (!obs_in || !obs_pos || !obs_head || !p_slot || !p_row || !traj || !vel || !init_pos || !init_heading) return ERR_ARG;
  LaunchScope ls(PROSIM_K_STATE, S(stream));
  init_traj_kernel<<<(P + 127) / 128, 128, 0, S(stream)>>>(obs_in, obs_pos, obs_head, p_slot, p_row, P, T, traj, vel,
                                                           init_pos, init_heading);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_step_env(const float* traj, const float* vel, const float* init_pos, const float* init_heading,
                    const int32_t* p_row, const int32_t* p_slot, int P, int T, int tidx, float* p_pos, float* p_ori,
                    float* fut_in, uint8_t* fut_mask, float* fut_pos, float* fut_head, prosim_stream_t stream) {
  if (P < 0 || tidx < HIST || tidx > T) return ERR_ARG;
  if (fut_in && tidx < HIST + 2) return ERR_ARG;
  if (P == 0) return 0;
  if (!traj || !vel || !init_pos || !init_heading || !p_row || !p_pos || !p_ori) return ERR_ARG;
  if (fut_in && (!fut_mask || !fut_pos || !fut_head || !p_slot)) return ERR_ARG;
  LaunchScope ls(PROSIM_K_STATE, S(stream));
  step_env_kernel<<<(P + 7) / 8, 128, 0, S(stream)>>>(traj, vel, init_pos, init_heading, p_row, p_slot, P, T, tidx,
                                                          p_pos, p_ori, fut_in, fut_mask, fut_pos, fut_head);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_gather_pose(const float* pos, const float* head, const int32_t* rows, int n, float* out_pos, float* out_ori,
                       prosim_stream_t stream) {
  if (n < 0) return ERR_ARG;
  if (n == 0) return 0;
  if (!pos || !head || !rows || !out_pos || !out_ori) return ERR_ARG;
  LaunchScope ls(PROSIM_K_STATE, S(stream));
  gather_pose_kernel<<<(n + 127) / 128, 128, 0, S(stream)>>>(pos, head, rows, n, out_pos, out_ori);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_step_agent_traj(const float* motion_pred, const int32_t* p_row, int P, int T, int tidx, float* traj, float* vel,
                           prosim_stream_t stream) {
  if (P < 0 || tidx < 1 || tidx + STEP > T) return ERR_ARG;
  if (P == 0) return 0;
  if (!motion_pred || !p_row || !traj || !vel) return ERR_ARG;
  LaunchScope ls(PROSIM_K_STATE, S(stream));
  step_agent_traj_kernel<<<(P + 127) / 128, 128, 0, S(stream)>>>(motion_pred, p_row, P, T, tidx, traj, vel);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

int prosim_rollout_to_world(const float* traj, const float* init_pos, const float* init_heading, const int32_t* p_row,
                            int P, int T, int t0, int steps, const float* tf, float* out, prosim_stream_t stream) {
  if (P < 0 || steps < 0 || t0 < 0 || t0 + steps > T) return ERR_ARG;
  if (P == 0 || steps == 0) return 0;
  if (!traj || !init_pos || !init_heading || !p_row || !tf || !out) return ERR_ARG;
  LaunchScope ls(PROSIM_K_STATE, S(stream));
  to_world_kernel<<<(P * steps + 255) / 256, 256, 0, S(stream)>>>(traj, init_pos, init_heading, p_row, P, T, t0, steps, tf,
                                                                  out);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

// ---- one policy tick: scratch layout inside the caller's workspace (all offsets 256-byte aligned)
namespace {
struct TickWs {
  size_t nbr_a, deg_a, nbr_m, deg_m, z_a, z_m, kv_a, attn, total;
  int stride_a, stride_m;
  size_t attn_floats;
};
inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }
bool tick_layout(const prosim_cfg_t& c, TickWs& w) {
  if (c.n_policy_rows < 0 || c.n_agent_tokens < 0 || c.n_map_tokens < 0 || c.max_neigh <= 0 || c.n_layers <= 0) return false;
  const size_t P = (size_t)c.n_policy_rows;
  w.stride_a = c.max_agents_per_scene < c.max_neigh ? c.max_agents_per_scene : c.max_neigh;
  w.stride_m = c.max_map_per_scene < c.max_neigh ? c.max_map_per_scene : c.max_neigh;
  if (w.stride_a < 1) w.stride_a = 1;
  if (w.stride_m < 1) w.stride_m = 1;
  size_t off = 0;
  w.nbr_a = off; off = al256(off + P * w.stride_a * 4);
  w.deg_a = off; off = al256(off + P * 4);
  w.nbr_m = off; off = al256(off + P * w.stride_m * 4);
  w.deg_m = off; off = al256(off + P * 4);
  w.z_a = off;   off = al256(off + P * w.stride_a * 96 * 4);
  w.z_m = off;   off = al256(off + P * w.stride_m * 96 * 4);
  w.kv_a = off;  off = al256(off + (size_t)c.n_layers * c.n_agent_tokens * 256 * 4);
  w.attn_floats = prosim_attn_workspace_floats(c.n_policy_rows, 0, w.stride_a > w.stride_m ? w.stride_a : w.stride_m);
  w.attn = off;  off = al256(off + w.attn_floats * 4);
  w.total = off;
  return true;
}
}  // namespace

size_t prosim_workspace_bytes(const prosim_cfg_t* cfg) {
  TickWs w;
  if (!cfg || !tick_layout(*cfg, w)) return 0;
  return w.total;
}

int prosim_policy_tick(const prosim_tick_t* t, void* workspace, size_t workspace_bytes, prosim_stream_t stream) {
  if (!t) return ERR_ARG;
  TickWs w;
  if (!tick_layout(t->cfg, w)) return ERR_ARG;
  const int P = t->cfg.n_policy_rows, L = t->cfg.n_layers;
  if (P == 0) return 0;
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return ERR_ARG;
  if (workspace_bytes < w.total) return ERR_WORKSPACE;
  if (!t->emd || !t->agent_type || !t->p_scene || !t->p_pos || !t->p_ori || !t->seg_agent || !t->seg_map || !t->w_a2p ||
      !t->w_m2p || !t->w_head || !t->dim_t16 || !t->fuse || !t->motion_pred)
    return ERR_ARG;
  if ((t->cfg.n_agent_tokens > 0 && (!t->x_agent || !t->agent_pos || !t->agent_ori)) ||
      (t->cfg.n_map_tokens > 0 && (!t->map_pos || !t->map_ori || !t->kv_map)))
    return ERR_ARG;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  int32_t* nbr_a = reinterpret_cast<int32_t*>(base + w.nbr_a);
  int32_t* deg_a = reinterpret_cast<int32_t*>(base + w.deg_a);
  int32_t* nbr_m = reinterpret_cast<int32_t*>(base + w.nbr_m);
  int32_t* deg_m = reinterpret_cast<int32_t*>(base + w.deg_m);
  float* z_a = reinterpret_cast<float*>(base + w.z_a);
  float* z_m = reinterpret_cast<float*>(base + w.z_m);
  float* kv_a = reinterpret_cast<float*>(base + w.kv_a);
  float* attn = reinterpret_cast<float*>(base + w.attn);
  const int cap = t->cfg.max_neigh;
  if (int e = prosim_build_radius_edges(t->p_pos, t->p_scene, P, t->agent_pos, t->seg_agent, t->agent_radius, cap, 0, nbr_a,
                                        deg_a, w.stride_a, stream)) return e;
  if (int e = prosim_build_radius_edges(t->p_pos, t->p_scene, P, t->map_pos, t->seg_map, t->map_radius, cap, 0, nbr_m, deg_m,
                                        w.stride_m, stream)) return e;
  if (int e = prosim_edge_pe(t->p_pos, t->p_ori, P, t->agent_pos, t->agent_ori, nbr_a, deg_a, w.stride_a, t->dim_t16, nullptr,
                             96, z_a, stream)) return e;
  if (int e = prosim_edge_pe(t->p_pos, t->p_ori, P, t->map_pos, t->map_ori, nbr_m, deg_m, w.stride_m, t->dim_t16, nullptr, 96,
                             z_m, stream)) return e;
  const size_t kv_a_stride = (size_t)t->cfg.n_agent_tokens * 256, kv_m_stride = (size_t)t->cfg.n_map_tokens * 256;
  if (t->cfg.n_agent_tokens > 0)
    if (int e = prosim_attn_kv(t->x_agent, t->cfg.n_agent_tokens, t->w_a2p, aw::SIZE, L, kv_a, kv_a_stride, stream)) return e;
  prosim_stack_side_t sa, sb;
  sa.w = t->w_a2p; sa.kv = kv_a; sa.kv_layer_stride = kv_a_stride;
  sa.graph = prosim_graph_t{z_a, nbr_a, deg_a, w.stride_a, w.stride_a, 96, 0};
  sb.w = t->w_m2p; sb.kv = t->kv_map; sb.kv_layer_stride = kv_m_stride;
  sb.graph = prosim_graph_t{z_m, nbr_m, deg_m, w.stride_m, w.stride_m, 96, 2};
  if (int e = prosim_attn_stack_fwd(t->emd, P, L, &sa, &sb, attn, w.attn_floats, t->fuse, stream)) return e;
  if (int e = prosim_policy_head_fwd(t->fuse, t->agent_type, P, t->w_head, t->noise, t->noise_std, t->motion_pred, stream))
    return e;
  if (t->traj) {
    if (!t->vel || !t->p_row) return ERR_ARG;
    if (int e = prosim_step_agent_traj(t->motion_pred, t->p_row, P, t->T, t->tidx, t->traj, t->vel, stream)) return e;
  }
  return 0;
}

int prosim_tc_gemm_test(const float* a, const float* w, float* c, int m, int split3, prosim_stream_t stream) {
  if (m <= 0 || !a || !w || !c) return ERR_ARG;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc::tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::TEST_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  tc::tc_gemm_test_kernel<<<(m + 127) / 128, 128, tc::TEST_SMEM, S(stream)>>>(a, w, c, m, split3);
  PROSIM_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
