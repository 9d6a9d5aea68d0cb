// Per-tick state kernels of the closed loop and the small dense heads.
//   init_traj_kernel        ProSim.init_agent_trajs            traj_sam.py:597-633
//   step_env_kernel         ProSim.step_env                    traj_sam.py:205-274 (+ geometry.py:19-58)
//   step_agent_traj_kernel  ProSim.step_agent_traj             traj_sam.py:276-349
//   policy_head_kernel      ActDecoder._compute_traj           policy/act_decoder.py:78-135 (+ layers/mlp.py:207-241)
//   mlp2_kernel             PromptEncoder / GoalConditionEncoder   prompt_encoder/base.py:30, condition_encoders.py:21-51
// State layout (HBM): traj [rows][T][4] = (x, y, sin, cos) in each agent's own t0 frame, vel [rows][T][2],
// T = 11 + rollout steps, preallocated once (the reference grows them with torch.cat every tick).
#pragma once
#include "common.cuh"
#include "gemm_tile.cuh"
#include "weights_layout.h"

namespace prosim {

constexpr int HIST = 11;
constexpr int STEP = 10;

__device__ __forceinline__ float2 rot2(float x, float y, float th) {
  // geometry.py:19-22  (x cos - y sin, y cos + x sin)
  float c = cosf(th), s = sinf(th);
  return make_float2(x * c - y * s, y * c + x * s);
}
__device__ __forceinline__ float nan0(float v) { return isnan(v) ? 0.f : v; }

// one thread per policy row
__global__ void init_traj_kernel(const float* __restrict__ obs_in, const float* __restrict__ obs_pos,
                                 const float* __restrict__ obs_head, const int* __restrict__ p_slot,
                                 const int* __restrict__ p_row, int P, int T, float* __restrict__ traj,
                                 float* __restrict__ vel, float* __restrict__ init_pos, float* __restrict__ init_heading) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int slot = p_slot[p], row = p_row[p];
  const float* src = obs_in + (size_t)slot * HIST * 24;
  for (int i = 0; i < HIST; ++i) {
    float4 t = make_float4(nan0(src[i * 24 + 0]), nan0(src[i * 24 + 1]), nan0(src[i * 24 + 2]), nan0(src[i * 24 + 3]));
    *reinterpret_cast<float4*>(traj + ((size_t)row * T + i) * 4) = t;
    *reinterpret_cast<float2*>(vel + ((size_t)row * T + i) * 2) = make_float2(nan0(src[i * 24 + 4]), nan0(src[i * 24 + 5]));
  }
  init_pos[row * 2 + 0] = obs_pos[slot * 2 + 0];
  init_pos[row * 2 + 1] = obs_pos[slot * 2 + 1];
  init_heading[row] = obs_head[slot];
}

// 16 lanes per policy row (lane i < 11 owns step i of the observation window).  tidx = number of valid steps in traj.
// When fut_in != nullptr (every tick but the first) the 11-step observation window of the agent is rebuilt in its
// current frame and written to fut_obs (input cols 0-7, position, heading, mask), exactly like the reference does in
// place.  (The first version ran one thread per row over the 11 steps: 4096 threads on the whole GPU, 45 us.)
__global__ void __launch_bounds__(128) step_env_kernel(const float* __restrict__ traj, const float* __restrict__ vel,
                                const float* __restrict__ init_pos, const float* __restrict__ init_heading,
                                const int* __restrict__ p_row, const int* __restrict__ p_slot, int P, int T, int tidx,
                                float* __restrict__ p_pos, float* __restrict__ p_ori, float* __restrict__ fut_in,
                                uint8_t* __restrict__ fut_mask, float* __restrict__ fut_pos,
                                float* __restrict__ fut_head) {
  const int p = blockIdx.x * 8 + (threadIdx.x >> 4), i = threadIdx.x & 15;
  if (p >= P) return;
  const int row = p_row[p];
  const float4* tr = reinterpret_cast<const float4*>(traj) + (size_t)row * T;
  const float2* vl = reinterpret_cast<const float2*>(vel) + (size_t)row * T;
  const float4 last = tr[tidx - 1];
  // quirk kept: world position adds the t0-frame offset without rotating it by init_heading (traj_sam.py:213)
  const float px = init_pos[row * 2 + 0] + last.x, py = init_pos[row * 2 + 1] + last.y;
  const float th_last = atan2f(last.z, last.w);
  const float heading = wrap_angle(th_last + init_heading[row]);
  if (i == 0) {
    p_pos[p * 2 + 0] = px;
    p_pos[p * 2 + 1] = py;
    p_ori[p] = heading;
  }
  if (fut_in == nullptr || i >= HIST) return;
  const int slot = p_slot[p];
  const float nth = -th_last;
  const float c = cosf(nth), s = sinf(nth);
  const float2 vp = vl[tidx - HIST + i - 1], v = vl[tidx - HIST + i];
  const float2 rv_prev = make_float2(vp.x * c - vp.y * s, vp.y * c + vp.x * s);
  const float2 rv = make_float2(v.x * c - v.y * s, v.y * c + v.x * s);
  const float4 w = tr[tidx - HIST + i];
  const float ox = w.x - last.x, oy = w.y - last.y;
  const float dth = wrap_angle(atan2f(w.z, w.w) - th_last);
  float* o = fut_in + ((size_t)slot * HIST + i) * 24;
  *reinterpret_cast<float4*>(o) = make_float4(ox * c - oy * s, oy * c + ox * s, sinf(dth), cosf(dth));
  *reinterpret_cast<float4*>(o + 4) = make_float4(rv.x, rv.y, (rv.x - rv_prev.x) / 0.1f, (rv.y - rv_prev.y) / 0.1f);
  unsigned long long* m = reinterpret_cast<unsigned long long*>(fut_mask + ((size_t)slot * HIST + i) * 24);
  m[0] = m[1] = m[2] = 0x0101010101010101ull;
  if (i == 0) {
    fut_pos[slot * 2 + 0] = px;
    fut_pos[slot * 2 + 1] = py;
    fut_head[slot] = heading;
  }
}

// compact gather of token poses: out_pos[i] = pos[rows[i]], out_ori[i] = head[rows[i]]
__global__ void gather_pose_kernel(const float* __restrict__ pos, const float* __restrict__ head,
                                   const int* __restrict__ rows, int n, float* __restrict__ out_pos,
                                   float* __restrict__ out_ori) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rows[i];
  out_pos[i * 2 + 0] = pos[r * 2 + 0];
  out_pos[i * 2 + 1] = pos[r * 2 + 1];
  out_ori[i] = head[r];
}

// one thread per policy row: append the 10 predicted steps (agent-frame deltas -> t0 frame)
__global__ void step_agent_traj_kernel(const float* __restrict__ motion_pred, const int* __restrict__ p_row, int P, int T,
                                       int tidx, float* __restrict__ traj, float* __restrict__ vel) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int row = p_row[p];
  float4* tr = reinterpret_cast<float4*>(traj) + (size_t)row * T;
  float2* vl = reinterpret_cast<float2*>(vel) + (size_t)row * T;
  const float4 last = tr[tidx - 1];
  const float th = atan2f(last.z, last.w);
  const float c = cosf(th), s = sinf(th);
  const float* mp = motion_pred + (size_t)p * STEP * 5;
  for (int i = 0; i < STEP; ++i) {
    const float x = mp[i * 5 + 0], y = mp[i * 5 + 1], h = mp[i * 5 + 2], vx = mp[i * 5 + 3], vy = mp[i * 5 + 4];
    const float nth = wrap_angle(th + h);
    tr[tidx + i] = make_float4((x * c - y * s) + last.x, (y * c + x * s) + last.y, sinf(nth), cosf(nth));
    vl[tidx + i] = make_float2(vx * c - vy * s, vy * c + vx * s);
  }
}

// Local (agent t0 frame) -> world: rollout/gpu_utils.py:255-266 (obtain_rollout_trajs_in_world) with
// rollout/utils.py:347-392 (batch_nd_transform_points_pt / angles_pt, angle_wrap).  tf = 3x3 row-major
// centre->world transform.  out[p][i] = (x_w, y_w, h_w) for the `steps` rolled-out steps starting at t0.
// One thread per (agent, step).
__global__ void to_world_kernel(const float* __restrict__ traj, const float* __restrict__ init_pos,
                                const float* __restrict__ init_heading, const int* __restrict__ p_row, int P, int T,
                                int t0, int steps, const float* __restrict__ tf, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * steps) return;
  const int p = idx / steps, i = idx % steps;
  const int row = p_row[p];
  const float4 w = reinterpret_cast<const float4*>(traj)[(size_t)row * T + t0 + i];
  const float ih = init_heading[row];
  const float2 xy = rot2(w.x, w.y, ih);
  const float xc = xy.x + init_pos[row * 2 + 0], yc = xy.y + init_pos[row * 2 + 1];
  const float hc = wrap_angle(atan2f(w.z, w.w) + ih);
  const float PI_F = 3.14159265358979323846f, TWO_PI_F = 6.28318530717958647692f;
  const float rot = atan2f(tf[3], tf[0]);
  float* o = out + (size_t)idx * 3;
  o[0] = (xc * tf[0] + yc * tf[1]) + tf[2];
  o[1] = (xc * tf[3] + yc * tf[4]) + tf[5];
  o[2] = py_mod((hc + rot) + PI_F, TWO_PI_F) - PI_F;       // angle_wrap: (a + pi) % 2pi - pi
}

// Linear(128->128)+LN+ReLU, Linear(128->64)+LN(64)+ReLU, Linear(64->n_out<=128) on a tile held in sIn;
// result left in sOut (columns >= n_out are zero).  Weight block layout: hw::MH0_W.. / hw::PM0_W.. pattern.
template <int RPT>
__device__ __forceinline__ void mlp3_tile(float* sIn, float* sTmp, float* sOut, const float* __restrict__ W) {
  constexpr int R = 2 * RPT;
  const int n = threadIdx.x & 127;
  float acc[RPT];
  acc_init(acc, __ldg(W + 16384 + n));
  gemm_tile_acc<RPT>(acc, sIn, LDS_PAD, D, W, D);
  acc_store_smem<RPT>(acc, sTmp, LDS_PAD, false);
  __syncthreads();
  ln_tile_inplace<D>(sTmp, LDS_PAD, R, W + 16384 + 128, W + 16384 + 256, true);
  __syncthreads();
  const float* W1 = W + 16384 + 384;
  acc_init(acc, __ldg(W1 + 16384 + n));
  gemm_tile_acc<RPT>(acc, sTmp, LDS_PAD, D, W1, D);
  acc_store_smem<RPT>(acc, sIn, LDS_PAD, false);
  __syncthreads();
  ln_tile_inplace<64>(sIn, LDS_PAD, R, W1 + 16384 + 128, W1 + 16384 + 256, true);
  __syncthreads();
  const float* W2 = W1 + 16384 + 384;
  acc_init(acc, __ldg(W2 + 64 * 128 + n));
  gemm_tile_acc<RPT>(acc, sIn, LDS_PAD, 64, W2, D);
  acc_store_smem<RPT>(acc, sOut, LDS_PAD, false);
  __syncthreads();
}

// feat: fused policy feature [P][128]; a_type: 1..3; motion_pred out [P][10][5] (xy cumsum, wrapped heading
// cumsum, velocity passthrough).  K = 1 so max over modes is the identity and CG's context == input.
// noise: optional [P][10][2] standard-normal draws (MODEL.POLICY.ACT_DECODER.RANDOM_NOISE_STD > 0).
template <int RPT>
__global__ void __launch_bounds__(256) policy_head_kernel(const float* __restrict__ feat, const int* __restrict__ a_type,
                                                          int P, const float* __restrict__ W,
                                                          const float* __restrict__ noise, float noise_std,
                                                          float* __restrict__ motion_pred) {
  constexpr int R = 2 * RPT;
  __shared__ __align__(16) float sS[R * LDS_PAD];
  __shared__ __align__(16) float sA[R * LDS_PAD];
  __shared__ __align__(16) float sB[R * LDS_PAD];
  const int row0 = blockIdx.x * R;
  const int n = threadIdx.x & 127, rg = threadIdx.x >> 7;
  for (int i = threadIdx.x; i < R * D; i += 256) {
    int r = i >> 7, c = i & 127;
    int t = row0 + r < P ? a_type[row0 + r] - 1 : 0;
    t = min(max(t, 0), 2);
    sA[r * LDS_PAD + c] = __ldg(W + hw::ANCHOR + t * D + c);
  }
  __syncthreads();
  float acc[RPT];
#pragma unroll 1
  for (int i = 0; i < 3; ++i) {
    const float* Wc = W + hw::CG_W + i * hw::CG_STRIDE;
    acc_init(acc, __ldg(Wc + 16384 + n));
    gemm_tile_acc<RPT>(acc, i == 0 ? sA : sS, LDS_PAD, D, Wc, D);
    acc_store_smem<RPT>(acc, sB, LDS_PAD, false);
    __syncthreads();
    ln_tile_inplace<D>(sB, LDS_PAD, R, Wc + 16384 + 128, Wc + 16384 + 256, true);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int lr = rg * RPT + r, row = row0 + lr;
      const float y = sB[lr * LDS_PAD + n];
      float sv;
      if (i == 0) {
        sv = y * (row < P ? feat[(size_t)row * D + n] : 0.f);
      } else {
        const float s_old = sS[lr * LDS_PAD + n];
        sv = (s_old * (float)i + y * s_old) / (float)(i + 1);
      }
      sS[lr * LDS_PAD + n] = sv;
    }
    __syncthreads();
  }
  mlp3_tile<RPT>(sS, sA, sB, W + hw::MH0_W);
  if (threadIdx.x < R && row0 + threadIdx.x < P) {
    const float* m = sB + threadIdx.x * LDS_PAD;
    float* o = motion_pred + (size_t)(row0 + threadIdx.x) * STEP * 5;
    // RANDOM_NOISE_STD > 0 (act_decoder.py:113-115): standard-normal draws [P][10][2] scaled by std are added to the
    // per-step displacements before the cumulative sum (a product then a sum in the reference: no fused multiply-add)
    const float* nz = noise ? noise + (size_t)(row0 + threadIdx.x) * STEP * 2 : nullptr;
    float cx = 0.f, cy = 0.f, ch = 0.f;
    for (int i = 0; i < STEP; ++i) {
      float dx = m[i * 5 + 0], dy = m[i * 5 + 1];
      if (nz) {
        dx = __fadd_rn(dx, __fmul_rn(nz[i * 2 + 0], noise_std));
        dy = __fadd_rn(dy, __fmul_rn(nz[i * 2 + 1], noise_std));
      }
      cx += dx;
      cy += dy;
      ch += m[i * 5 + 2];
      o[i * 5 + 0] = cx;
      o[i * 5 + 1] = cy;
      o[i * 5 + 2] = wrap_angle(ch);
      o[i * 5 + 3] = m[i * 5 + 3];
      o[i * 5 + 4] = m[i * 5 + 4];
    }
  }
}

// Throughput variant of policy_head_kernel for launches of >= 1024 rows: the six GEMMs go through the shared-memory weight
// stream of gemm_tile.cuh (gemm2 / WPipe: a weight element is fetched once per CTA, two chunks ahead, across GEMM
// boundaries) instead of per-thread LDG.  Same summation order per output (ascending k): bit-identical to policy_head_kernel.
template <int TR>
struct Head2Smem {
  static constexpr int M = 8 * TR;
  static constexpr size_t bytes = WPIPE_BYTES + 3 * (size_t)M * LDS_PAD * sizeof(float);
};

template <int TR>
__global__ void __launch_bounds__(256) policy_head2_kernel(const float* __restrict__ feat, const int* __restrict__ a_type,
                                                           int P, const float* __restrict__ W,
                                                           const float* __restrict__ noise, float noise_std,
                                                           float* __restrict__ motion_pred) {
  constexpr int NW = 8, M = NW * TR;
  extern __shared__ __align__(16) float smem[];
  WPipe p = wpipe_init<NW>(smem, [&](WSeg* sg) {
    for (int i = 0; i < 3; ++i) sg[i] = WSeg{W + hw::CG_W + i * hw::CG_STRIDE, D, D};
    sg[3] = WSeg{W + hw::MH0_W, D, D};
    sg[4] = WSeg{W + hw::MH1_W, D, D};
    sg[5] = WSeg{W + hw::MH2_W, D, 64};
    return 6;
  });
  float* sS = smem + WPIPE_BYTES / sizeof(float);
  float* sA = sS + M * LDS_PAD;
  float* sB = sA + M * LDS_PAD;
  const int row0 = blockIdx.x * M;
  for (int i = threadIdx.x; i < M * D; i += 256) {
    const int r = i >> 7, c = i & 127;
    int t = row0 + r < P ? a_type[row0 + r] - 1 : 0;
    t = min(max(t, 0), 2);
    sA[r * LDS_PAD + c] = __ldg(W + hw::ANCHOR + t * D + c);
  }
  __syncthreads();
  const TileCoord tc = tile_coord<TR>();
  float acc[TR][4];
#pragma unroll 1
  for (int i = 0; i < 3; ++i) {
    const float* Wc = W + hw::CG_W + i * hw::CG_STRIDE;
    acc2_init_bias<TR>(acc, Wc + 16384);
    gemm2<TR, NW>(acc, i == 0 ? sA : sS, LDS_PAD, p);
    acc2_store_smem<TR>(acc, sB, LDS_PAD, false);
    __syncthreads();
    ln_tile_inplace<D>(sB, LDS_PAD, M, Wc + 16384 + 128, Wc + 16384 + 256, true);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int lr = tc.row + r, row = row0 + lr;
      const float4 y = *reinterpret_cast<const float4*>(sB + lr * LDS_PAD + tc.col);
      float4 sv;
      if (i == 0) {
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < P) f = *reinterpret_cast<const float4*>(feat + (size_t)row * D + tc.col);
        sv = make_float4(y.x * f.x, y.y * f.y, y.z * f.z, y.w * f.w);
      } else {
        const float4 so = *reinterpret_cast<const float4*>(sS + lr * LDS_PAD + tc.col);
        const float fi = (float)i, fd = (float)(i + 1);
        sv = make_float4((so.x * fi + y.x * so.x) / fd, (so.y * fi + y.y * so.y) / fd, (so.z * fi + y.z * so.z) / fd,
                         (so.w * fi + y.w * so.w) / fd);
      }
      *reinterpret_cast<float4*>(sS + lr * LDS_PAD + tc.col) = sv;
    }
    __syncthreads();
  }
  // motion_head: Linear+LN+ReLU (128), Linear+LN+ReLU (64 real columns), Linear (50 real columns)
  acc2_init_bias<TR>(acc, W + hw::MH0_B);
  gemm2<TR, NW>(acc, sS, LDS_PAD, p);
  acc2_store_smem<TR>(acc, sA, LDS_PAD, false);
  __syncthreads();
  ln_tile_inplace<D>(sA, LDS_PAD, M, W + hw::MH0_G, W + hw::MH0_BB, true);
  __syncthreads();
  acc2_init_bias<TR>(acc, W + hw::MH1_B);
  gemm2<TR, NW>(acc, sA, LDS_PAD, p);
  acc2_store_smem<TR>(acc, sS, LDS_PAD, false);
  __syncthreads();
  ln_tile_inplace<64>(sS, LDS_PAD, M, W + hw::MH1_G, W + hw::MH1_BB, true);
  __syncthreads();
  acc2_init_bias<TR>(acc, W + hw::MH2_B);
  gemm2<TR, NW>(acc, sS, LDS_PAD, p);
  acc2_store_smem<TR>(acc, sB, LDS_PAD, false);
  __syncthreads();
  if (threadIdx.x < M && row0 + threadIdx.x < P) {
    const float* m = sB + threadIdx.x * LDS_PAD;
    float* o = motion_pred + (size_t)(row0 + threadIdx.x) * STEP * 5;
    const float* nz = noise ? noise + (size_t)(row0 + threadIdx.x) * STEP * 2 : nullptr;
    float cx = 0.f, cy = 0.f, ch = 0.f;
    for (int i = 0; i < STEP; ++i) {
      float dx = m[i * 5 + 0], dy = m[i * 5 + 1];
      if (nz) {
        dx = __fadd_rn(dx, __fmul_rn(nz[i * 2 + 0], noise_std));
        dy = __fadd_rn(dy, __fmul_rn(nz[i * 2 + 1], noise_std));
      }
      cx += dx;
      cy += dy;
      ch += m[i * 5 + 2];
      o[i * 5 + 0] = cx;
      o[i * 5 + 1] = cy;
      o[i * 5 + 2] = wrap_angle(ch);
      o[i * 5 + 3] = m[i * 5 + 3];
      o[i * 5 + 4] = m[i * 5 + 4];
    }
  }
}

// pred_mlp on the policy embeddings (act_decoder.py:129-131): out [P][2]
template <int RPT>
__global__ void __launch_bounds__(256) reconst_kernel(const float* __restrict__ emd, int P, const float* __restrict__ W,
                                                      float* __restrict__ out) {
  constexpr int R = 2 * RPT;
  __shared__ __align__(16) float sS[R * LDS_PAD];
  __shared__ __align__(16) float sA[R * LDS_PAD];
  __shared__ __align__(16) float sB[R * LDS_PAD];
  const int row0 = blockIdx.x * R;
  load_tile128(sS, LDS_PAD, emd, row0, P, R);
  __syncthreads();
  mlp3_tile<RPT>(sS, sA, sB, W + hw::PM0_W);
  if (threadIdx.x < R * 2) {
    int r = threadIdx.x >> 1, c = threadIdx.x & 1;
    if (row0 + r < P) out[(size_t)(row0 + r) * 2 + c] = sB[r * LDS_PAD + c];
  }
}

// Linear(K0<=8 -> 128) [+LN] + ReLU + Linear(128 -> 128) [+ fixed Fourier embedding of a scalar t]
// in: [N][ld_in] (first k0 columns used); tpe_t: optional [N] scalars, dim_t128: [128] denominators
template <int RPT>
__global__ void __launch_bounds__(256) mlp2_kernel(const float* __restrict__ in, int ld_in, int k0, int N, int use_ln,
                                                   const float* __restrict__ W, const float* __restrict__ tpe_t,
                                                   int tpe_ld, const float* __restrict__ dim_t128,
                                                   float* __restrict__ out) {
  constexpr int R = 2 * RPT;
  __shared__ __align__(16) float sA[R * LDS_PAD];
  __shared__ __align__(16) float sB[R * LDS_PAD];
  const int row0 = blockIdx.x * R;
  const int n = threadIdx.x & 127, rg = threadIdx.x >> 7;
  for (int i = threadIdx.x; i < R * 8; i += 256) {
    int r = i >> 3, c = i & 7;
    sA[r * LDS_PAD + c] = (row0 + r < N && c < k0) ? in[(size_t)(row0 + r) * ld_in + c] : 0.f;
  }
  __syncthreads();
  float acc[RPT];
  acc_init(acc, __ldg(W + mw::B0 + n));
  gemm_tile_acc<RPT>(acc, sA, LDS_PAD, 8, W + mw::W0, D);
  acc_store_smem<RPT>(acc, sB, LDS_PAD, !use_ln);
  __syncthreads();
  if (use_ln) {
    ln_tile_inplace<D>(sB, LDS_PAD, R, W + mw::G0, W + mw::BB0, true);
    __syncthreads();
  }
  acc_init(acc, __ldg(W + mw::B1 + n));
  gemm_tile_acc<RPT>(acc, sB, LDS_PAD, D, W + mw::W1, D);
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int row = row0 + rg * RPT + r;
    if (row >= N) continue;
    float v = acc[r];
    if (tpe_t != nullptr) {
      const float a = (tpe_t[(size_t)row * tpe_ld] * 6.28318530717958647692f) / dim_t128[n];
      v += (n & 1) ? cosf(a) : sinf(a);
    }
    out[(size_t)row * D + n] = v;
  }
}

// MotionTagEncoder for the unary action tags (condition_transformer/condition_encoders.py:76-145): one learned
// vector per tag plus FourierEmbeddingFix(64) of the start and of the end step, concatenated (:132-137).
// tags: int64 [n][3] = (tag id, start, end) as the dataset stores them; table: [16][128], row = tag id.
// Rows whose id is outside [0, n_tags) (the -1 padding) are written as zeros.  One thread per output float.
__global__ void __launch_bounds__(256) tag_embed_kernel(const long long* __restrict__ tags, int n, int n_tags,
                                                        const float* __restrict__ table,
                                                        const float* __restrict__ dim_t64, float* __restrict__ out) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int e = idx >> 7, c = idx & 127;
  if (e >= n) return;
  const long long id = tags[(size_t)e * 3];
  float v = 0.f;
  if (id >= 0 && id < n_tags) {
    const float t = (float)tags[(size_t)e * 3 + 1 + (c >> 6)];
    const float a = (t * 6.28318530717958647692f) / dim_t64[c & 63];
    v = table[(int)id * D + c] + ((c & 1) ? cosf(a) : sinf(a));
  }
  out[(size_t)e * D + c] = v;
}

// GNNConditionAttn._construct_cond_edge_matrix + _pool_edges restricted to unary conditions (self edges)
// (condition_transformer/condition_attns.py:114-189, COND_POOL_FUNC 'mean'): per policy row, the condition embeddings
// of the condition types present on it are summed in type order and divided by their number.
// slot: int32 [P][n_slots], row of emb for (policy row, condition type) or -1.  extra: [P][128]; has[p] = 1 if any.
__global__ void __launch_bounds__(256) cond_pool_kernel(const float* __restrict__ emb, const int* __restrict__ slot,
                                                        int P, int n_slots, float* __restrict__ extra,
                                                        int* __restrict__ has) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int p = idx >> 7, c = idx & 127;
  if (p >= P) return;
  float s = 0.f;
  int cnt = 0;
  for (int m = 0; m < n_slots; ++m) {
    const int e = slot[(size_t)p * n_slots + m];
    if (e >= 0) {
      s = __fadd_rn(s, emb[(size_t)e * D + c]);
      ++cnt;
    }
  }
  extra[(size_t)p * D + c] = cnt ? __fdiv_rn(s, (float)cnt) : 0.f;
  if (c == 0) has[p] = cnt > 0;
}

// MODEL.OBS_UPDATE.FUSION = 'mlp' (scene_encoder/attn_fusion.py:177-203): for every agent that was observed at the previous
// tick too, new token <- MLP([old token | new token]) (Linear 256 -> 128, LN, ReLU, Linear 128 -> 128), in place.
// idx_old / idx_new: token rows of the n agents in the old / new agent-token buffers.
template <int RPT>
__global__ void __launch_bounds__(256) obs_fuse_kernel(const float* __restrict__ x_old, const int* __restrict__ idx_old,
                                                       float* __restrict__ x_new, const int* __restrict__ idx_new, int n,
                                                       const float* __restrict__ W) {
  constexpr int R = 2 * RPT;
  constexpr int LDA = 2 * D + 4;
  __shared__ __align__(16) float sA[R * LDA];
  __shared__ __align__(16) float sB[R * LDS_PAD];
  const int row0 = blockIdx.x * R;
  const int col = threadIdx.x & 127, rg = threadIdx.x >> 7;
  for (int i = threadIdx.x; i < R * 64; i += 256) {
    const int r = i >> 6, c4 = i & 63;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < n) {
      const float* src = c4 < 32 ? x_old + (size_t)idx_old[row0 + r] * D + 4 * c4 : x_new + (size_t)idx_new[row0 + r] * D + 4 * (c4 - 32);
      v = *reinterpret_cast<const float4*>(src);
    }
    *reinterpret_cast<float4*>(sA + r * LDA + 4 * c4) = v;
  }
  __syncthreads();
  float acc[RPT];
  acc_init(acc, __ldg(W + fw::B0 + col));
  gemm_tile_acc<RPT>(acc, sA, LDA, 2 * D, W + fw::W0, D);
  acc_store_smem<RPT>(acc, sB, LDS_PAD, false);
  __syncthreads();
  ln_tile_inplace<D>(sB, LDS_PAD, R, W + fw::G0, W + fw::BB0, true);
  __syncthreads();
  acc_init(acc, __ldg(W + fw::B1 + col));
  gemm_tile_acc<RPT>(acc, sB, LDS_PAD, D, W + fw::W1, D);
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int row = row0 + rg * RPT + r;
    if (row < n) x_new[(size_t)idx_new[row] * D + col] = acc[r];
  }
}

}  // namespace prosim
