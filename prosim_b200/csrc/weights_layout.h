// Packed weight layout (float offsets) shared by the kernels; mirrored by prosim_b200/weights.py::pack_*.
// All dense weights are stored K-major ("transposed", [in][out]) so a CTA reads them coalesced.
#pragma once

namespace prosim {
namespace aw {  // one AttentionLayer (reference: prosim/models/layers/attention_layer.py:13-54), refactored:
//   r_hat = z * gamma_r + beta_r  (z = LayerNorm(r) without affine, layer independent)
//   k_e = K'_j + Wkr diag(gamma_r) z_e          K' = Wk LN_src(x_src) + Wkr beta_r
//   v_e = V'_j + Wvr diag(gamma_r) z_e          V' = Wv LN_src(x_src) + b_v + Wvr beta_r + b_vr
//   q.k_e = q.K'_j + (diag(gamma_r) Wkr_h^T q_h).z_e   -> Qhat [8][128] per destination row
//   sum_e a_e v_e = sum_e a_e V'_j + Wvr diag(gamma_r) (sum_e a_e z_e)   -> Rbar [8][128] per destination row
// WQT/BQ are pre-scaled by head_dim^-0.5 = 0.25 (exact in fp32).
constexpr int LN_SRC_G = 0;
constexpr int LN_SRC_B = LN_SRC_G + 128;
constexpr int LN_DST_G = LN_SRC_B + 128;
constexpr int LN_DST_B = LN_DST_G + 128;
constexpr int WQT = LN_DST_B + 128;      // [128 k][128 n]
constexpr int BQ = WQT + 16384;
constexpr int WKT = BQ + 128;            // [128][128]
constexpr int KB = WKT + 16384;          // Wkr beta_r
constexpr int WVT = KB + 128;            // [128][128]
constexpr int VB = WVT + 16384;          // b_v + Wvr beta_r + b_vr
constexpr int WKRG = VB + 128;           // [128 n=h*16+c][128 d] = gamma_r[d] * Wkr[n][d]
constexpr int WVRGT = WKRG + 16384;      // [128 d][128 c] = gamma_r[d] * Wvr[c][d]
constexpr int WVRG96T = WVRGT + 16384;   // [96 d][128 c]: WVRGT with rows 96..127 added onto rows 64..95, for 96-wide z
constexpr int WST = WVRG96T + 12288;     // [128][128]
constexpr int BS = WST + 16384;
constexpr int WGAT = BS + 128;           // to_g columns 0..127 (agg part), [128 k][128 n]
constexpr int WGXT = WGAT + 16384;       // to_g columns 128..255 (x_dst part)
constexpr int BG = WGXT + 16384;
constexpr int WOT = BG + 128;
constexpr int BO = WOT + 16384;
constexpr int LN_POST_G = BO + 128;
constexpr int LN_POST_B = LN_POST_G + 128;
constexpr int LN_FFPRE_G = LN_POST_B + 128;
constexpr int LN_FFPRE_B = LN_FFPRE_G + 128;
constexpr int W1T = LN_FFPRE_B + 128;    // [128 k][512 n]
constexpr int B1 = W1T + 65536;
constexpr int W2T = B1 + 512;            // [512 k][128 n]
constexpr int B2 = W2T + 65536;
constexpr int LN_FFPOST_G = B2 + 128;
constexpr int LN_FFPOST_B = LN_FFPOST_G + 128;
constexpr int FP32_SIZE = LN_FFPOST_B + 128;
// ---- tensor-core operand copies of the dense weights (tc_post.cuh): every [N][K] B-operand chunk is stored as
// tf32 "hi" then tf32 "lo" (hi = rna_tf32(w), lo = rna_tf32(w - hi)), each in the tcgen05 K-major no-swizzle
// core-matrix order  idx(n, k) = (n / 8) * (8 K) + (k / 4) * 32 + (n % 8) * 4 + k % 4,  chunks in the order the
// kernel's MMA warp consumes them, so the producer warp streams the block with plain bulk copies.
constexpr int TC_VR96 = FP32_SIZE;             // 8 heads x [16 c][96 d]   (WVRG96T columns of the head)
constexpr int TC_VR128 = TC_VR96 + 8 * 3072;   // 8 heads x [16 c][128 d]  (WVRGT)
constexpr int TC_GA = TC_VR128 + 8 * 4096;     // 4 k-chunks x [128 n][32 k]
constexpr int TC_O = TC_GA + 4 * 8192;         // 4 k-chunks x [128 n][32 k]
constexpr int TC_FF = TC_O + 4 * 8192;         // up_0, up_1, then (up_{j+2}, down_j) for j = 0..13, down_14, down_15:
                                               //   up_j = W1 rows 32j..32j+31 [32 n][128 k], down_j = W2[:, 32j..] [128 n][32 k]
constexpr int TC_Q = TC_FF + 32 * 8192;        // 4 k-chunks each: to_q (pre-scaled), to_s, to_g[x part]
constexpr int TC_S = TC_Q + 4 * 8192;
constexpr int TC_GX = TC_S + 4 * 8192;
constexpr int TC_KRG = TC_GX + 4 * 8192;       // 8 heads x [128 d][16 c]  (WKRG rows of the head, transposed)
constexpr int TC_KV = TC_KRG + 8 * 4096;       // kv_tc.cuh: 4 k-chunks x ([128 n][32 k] of to_k, then of to_v), interleaved
constexpr int TC_FF2 = TC_KV + 8 * 8192;       // post_sw.cuh (weights on the M side): FFN in hidden tiles of 128, 4 k-chunks [128 m][32 k] each,
                                               //   in consumption order up_0, up_1, down_0, up_2, down_1, up_3, down_2, down_3
                                               //   up_t = W1 rows 128t..128t+127, down_t = W2[:, 128t..128t+127]
constexpr int SIZE = TC_FF2 + 32 * 8192;
}  // namespace aw

namespace pw {  // PointNet polyline encoder (scene_encoder/pointnet_encoder.py:13-62)
// pre_mlps: up to 3 layers; layer 0 is [KIN_PAD][128] (input dim padded to a multiple of 4 with zero rows).
// Each layer block: Wt, bias[128], ln_g[128], ln_b[128]  (ln unused on the last pre layer)
constexpr int PRE0_W = 0;                    // [24][128]  (obs: 24 inputs; map: 11 padded to 12, rest zero)
constexpr int PRE0_B = PRE0_W + 24 * 128;
constexpr int PRE0_G = PRE0_B + 128;
constexpr int PRE0_BB = PRE0_G + 128;
constexpr int PRE1_W = PRE0_BB + 128;        // [128][128] (map only)
constexpr int PRE1_B = PRE1_W + 16384;
constexpr int PRE1_G = PRE1_B + 128;
constexpr int PRE1_BB = PRE1_G + 128;
constexpr int PRE2_W = PRE1_BB + 128;        // [128][128] (map only)
constexpr int PRE2_B = PRE2_W + 16384;
constexpr int MLP0_WA = PRE2_B + 128;        // mlps.0 columns 0..127 (point feature part)  [128][128]
constexpr int MLP0_WB = MLP0_WA + 16384;     // mlps.0 columns 128..255 (pooled part)       [128][128]
constexpr int MLP0_B = MLP0_WB + 16384;
constexpr int MLP0_G = MLP0_B + 128;
constexpr int MLP0_BB = MLP0_G + 128;
constexpr int MLP1_W = MLP0_BB + 128;        // [128][128]
constexpr int MLP1_B = MLP1_W + 16384;
constexpr int OUT0_W = MLP1_B + 128;         // [128][128]
constexpr int OUT0_B = OUT0_W + 16384;
constexpr int OUT1_W = OUT0_B + 128;         // [128][128]
constexpr int OUT1_B = OUT1_W + 16384;
constexpr int SIZE = OUT1_B + 128;
}  // namespace pw

namespace hw {  // policy head (policy/act_decoder.py:50-135): anchors, CG_decode x3, motion_head, pred_mlp
constexpr int ANCHOR = 0;                    // [3][128]
constexpr int CG_W = ANCHOR + 3 * 128;       // 3 x { Wt[128][128], b[128], ln_g[128], ln_b[128] }
constexpr int CG_STRIDE = 16384 + 3 * 128;
constexpr int MH0_W = CG_W + 3 * CG_STRIDE;  // motion_head: [128][128] b g bb
constexpr int MH0_B = MH0_W + 16384;
constexpr int MH0_G = MH0_B + 128;
constexpr int MH0_BB = MH0_G + 128;
constexpr int MH1_W = MH0_BB + 128;          // [128][128] (64 real columns, zero padded) b g bb (64 real)
constexpr int MH1_B = MH1_W + 16384;
constexpr int MH1_G = MH1_B + 128;
constexpr int MH1_BB = MH1_G + 128;
constexpr int MH2_W = MH1_BB + 128;          // [64][128] (50 real columns)
constexpr int MH2_B = MH2_W + 64 * 128;
constexpr int PM0_W = MH2_B + 128;           // pred_mlp, same shape pattern, 2 real output columns
constexpr int PM0_B = PM0_W + 16384;
constexpr int PM0_G = PM0_B + 128;
constexpr int PM0_BB = PM0_G + 128;
constexpr int PM1_W = PM0_BB + 128;
constexpr int PM1_B = PM1_W + 16384;
constexpr int PM1_G = PM1_B + 128;
constexpr int PM1_BB = PM1_G + 128;
constexpr int PM2_W = PM1_BB + 128;
constexpr int PM2_B = PM2_W + 64 * 128;
constexpr int SIZE = PM2_B + 128;
}  // namespace hw

namespace mw {  // two-layer MLP  Linear(K0->128) [+LN] ReLU Linear(128->128)   (prompt / goal encoders)
constexpr int W0 = 0;                        // [8][128]  (7 or 2 inputs, zero padded to 8)
constexpr int B0 = W0 + 8 * 128;
constexpr int G0 = B0 + 128;
constexpr int BB0 = G0 + 128;
constexpr int W1 = BB0 + 128;                // [128][128]
constexpr int B1 = W1 + 16384;
constexpr int SIZE = B1 + 128;
}  // namespace mw

namespace fw {  // obs_update_mlp (scene_encoder/attn_fusion.py:18-19, 196): Linear(256 -> 128) LN ReLU Linear(128 -> 128)
constexpr int W0 = 0;                        // [256 k][128 n]: rows 0..127 multiply the OLD token, 128..255 the NEW one
constexpr int B0 = W0 + 256 * 128;
constexpr int G0 = B0 + 128;
constexpr int BB0 = G0 + 128;
constexpr int W1 = BB0 + 128;                // [128][128]
constexpr int B1 = W1 + 16384;
constexpr int SIZE = B1 + 128;
}  // namespace fw
}  // namespace prosim
