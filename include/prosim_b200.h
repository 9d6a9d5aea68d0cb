/* prosim_b200 -- C ABI of the B200-native closed-loop rollout path of ProSim.
 *
 * The reference (Ariostgx/ProSim @ 78a398c) is pure Python/PyTorch and has NO FFI of its own: its
 * "plugin" boundary is the class registry (prosim/core/registry.py:25-136) behind which
 * ProSim.forward (prosim/models/traj_sam.py:59-175) runs.  This header is therefore the boundary a
 * maintainer would bind from Python (ctypes) to replace, one for one, the native
 * work the reference delegates to ATen / torch_cluster / torch_geometric on this path.  Each entry
 * point cites the reference code it replaces.
 *
 * Conventions: every pointer is a DEVICE pointer owned by the caller (PyTorch's allocator on the Python
 * side); nothing is allocated, freed or retained; calls only ENQUEUE work on `stream` and return
 * 0 on success, a positive cudaError_t on a launch failure, or a negative value for a bad argument.
 * The data-path functions keep no state besides one-time per-device kernel attribute setup and a per-host-thread cache of
 * encoded TMA descriptors.  The measurement / A-B switches (prosim_set_tensor_core, prosim_set_stack_split,
 * prosim_profile_enable, prosim_launch_count) ARE process-global by design: set them from one thread, not mid-forward.
 * Index arrays are int32, masks are uint8 (torch.bool), floats are IEEE fp32.
 */
#ifndef PROSIM_B200_H
#define PROSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* prosim_stream_t; /* cudaStream_t */

/* Fixed-stride neighbour lists + the per-edge normalised relative PE that go with them. */
typedef struct {
  const float* z;      /* [n_dst*stride][zd] LayerNorm(rel PE) without affine, row = dst*stride + j  */
  const int32_t* nbr;  /* [n_dst*stride] source row of edge j of destination dst (ascending)          */
  const int32_t* deg;  /* [n_dst] number of valid edges per destination                               */
  int32_t stride;      /* row stride of nbr / z                                                        */
  int32_t max_deg;     /* upper bound of deg                                                           */
  int32_t zd;          /* width of a z row: 128, or 96 for pure rel-PE edges (features 96..127 duplicate
                          64..95 and are dropped; prosim_edge_pe writes either)                         */
  int32_t warps_per_row; /* hint: 32-edge tiles expected per destination (1,2,3,6); 0 = derive from max_deg */
} prosim_graph_t;

/* One side of an alternating attention stack (prosim_attn_stack_fwd). */
typedef struct {
  const float* w;          /* n_layers packed AttentionLayers, consecutive (prosim_attn_layer_floats() each) */
  const float* kv;         /* fixed sources: precomputed [n_layers][n_src][256] K'|V' ; NULL => sources are the
                              destination rows themselves (non-bipartite layer), K'|V' recomputed every layer   */
  size_t kv_layer_stride;  /* floats between layers in kv                                                       */
  prosim_graph_t graph;
} prosim_stack_side_t;

int prosim_abi_version(void);
/* Kernel selection: 1 (default) = tcgen05 / TMEM 3xTF32 kernels for the node-side GEMMs of the AttentionLayer
 * (csrc/post_sw.cuh: 16 or 32 destination rows per CTA for launches of up to 9472 rows, incl. its "pre-only" mode for the
 * first layer of a stack; csrc/tc_post.cuh, 128 rows per CTA, above), the K'|V' projections (csrc/kv_tc.cuh) and the
 * PointNet encoders (csrc/pointnet_tc.cuh), and the one-launch edge kernel (csrc/edge_row.cuh) for launches of <= 592 rows;
 * 0 = fp32 FFMA kernels and the three-kernel edge path everywhere (A/B measurement and parity cross-checks).  Other values
 * select by bit (1 node kernels, 2 K'|V', 4 PointNet, 8 allow the 16/32-row node kernel, 16 the one-launch edge kernel) for
 * fault isolation.  PROCESS-GLOBAL switch: not part of the re-entrant data path, must not be flipped while another thread
 * enqueues work.  Environment: PROSIM_TC_MASK = initial mask (fault isolation from a fresh process); PROSIM_NO_PDL=1 (read
 * once, at the first launch) turns programmatic dependent launch off for the kernels that use it. */
int prosim_set_tensor_core(int on);
/* prosim_attn_stack_fwd with fixed sources on both sides runs as up to `parts` (1..4, default 1) independent row chains of
 * >= 1024 rows on internal side streams (forked from / joined to the caller's stream with events; capturable in a CUDA
 * graph).  Results are bit-identical for every setting.  Measured on B200 (bench workload): 31.4 ms per forward at 1 part,
 * 30.9 at 2, 30.6 at 3, 31.4 at 4.  The chains drift into lockstep (both in their edge phase, then both in the node
 * kernel: gpurun_out/split_timeline2.txt), so the gain is the node kernels of the chains sharing the chip, not an
 * edge-phase / node-phase overlap; starting the chains half a layer apart and reserving SMs for the sibling's node kernel
 * was tried and did not hold the stagger (the edge phase of half the rows is longer than a node kernel).  Off by default:
 * 2.6 % is not worth per-launch timings that overlap (the roofline accounting of bench.py times single launches). */
int prosim_set_stack_split(int parts);
/* Measurement support: SM-clock timestamps of the phases of CTA 0 of the last tcgen05 node-kernel launch
 * ([0..15] epilogue thread, [16..31] MMA thread; csrc/tc_post.cuh TCP_MARK). */
int prosim_tc_debug_read(long long* out32);
/* Test support: fills the shared memory (226 KB) and all 512 tensor-memory columns of every SM with `pattern` (e.g. NaN).
 * Neither is cleared between kernels or processes; a forward whose result changes after a scrub consumes on-chip state it
 * never wrote (tools/poison_check.py). */
int prosim_debug_scrub(float pattern, void* stream);

/* Launch accounting and optional per-kernel CUDA-event timing (measurement support for bench.py; not part of
 * the data path).  kernel_class < 0 in prosim_launch_count = all classes; prosim_profile_enable(-1) disables. */
enum {
  PROSIM_K_POINTNET = 0, PROSIM_K_RADIUS = 1, PROSIM_K_KNN = 2, PROSIM_K_EDGE_PE = 3, PROSIM_K_ATTN_KV = 4,
  PROSIM_K_ATTN_DSTPRE = 5, PROSIM_K_ATTN_EDGE = 6, PROSIM_K_ATTN_POST = 7, PROSIM_K_HEAD = 8, PROSIM_K_MLP2 = 9,
  PROSIM_K_STATE = 10, PROSIM_K_EDGE_QK = 11, PROSIM_K_EDGE_AV = 12,
  PROSIM_K_ATTN_POST_SW = 13 /* the launches of PROSIM_K_ATTN_POST that took the 32-row tcgen05 kernel (csrc/post_sw.cuh) */
};
long long prosim_launch_count(int kernel_class);
int prosim_profile_enable(int kernel_class);
int prosim_profile_read(double* total_ms, int* count);
/* sizes (in floats) of the packed weight blocks -- cross-checked by the Python packer */
int prosim_attn_layer_floats(void);
int prosim_pointnet_floats(void);
int prosim_head_floats(void);
int prosim_mlp2_floats(void);
/* scratch floats needed by prosim_attn_layer_fwd / prosim_attn_stack_fwd for n_dst destination rows, n_src self-source
 * rows and neighbour lists of row stride <= max_stride */
size_t prosim_attn_workspace_floats(int n_dst, int n_src, int max_stride);

/* PointNet polyline encoder: prosim/models/scene_encoder/pointnet_encoder.py:24-62.
 * kind 0 = agent history (11 points x 24 features, mask uint8 [.,11,24]; obs_encoder.py:75-86)
 * kind 1 = map polyline  (19 vectors x 11 features, mask uint8 [.,19];   map_encoder.py:67-88)
 * kind 2 / 3 = drag-point condition (16 / 8 points x 2 features; condition_transformer/condition_encoders.py:147-191);
 *          mask may be NULL: a point is then valid when neither coordinate is NaN (:178)
 * rows[n_poly] selects the polylines to encode out of x; out is compact [n_poly][128].
 * w = fp32 block (prosim_pointnet_floats), w_tc (nullable) = the same matrices as tcgen05 operands (K = 32 chunks
 * pre-split into tf32 hi / lo, weights.py::pack_pointnet_tc): with w_tc and prosim_set_tensor_core(1) the GEMM chain runs
 * on the tensor cores (csrc/pointnet_tc.cuh, 3xTF32), otherwise on fp32 FFMA (csrc/pointnet.cuh). */
int prosim_pointnet_fwd(int kind, const float* x, const uint8_t* mask, const int32_t* rows, int n_poly,
                        const float* w, const float* w_tc, float* out, prosim_stream_t stream);

/* torch_cluster.radius / radius_graph (call sites decoder/sym_coord.py:86,94; policy/act_decoder.py:250,259).
 * seg[B][4] = {start0,len0,start1,len1}: source ranges of each scene.  drop_self=1 gives radius_graph(loop=False)
 * semantics (cap+1 candidates, then the query itself removed; query index == source index). */
int prosim_build_radius_edges(const float* qpos, const int32_t* qscene, int n_q, const float* spos,
                              const int32_t* seg, float r, int cap, int drop_self, int32_t* nbr, int32_t* deg,
                              int stride, prosim_stream_t stream);
/* torch_cluster.knn_graph(loop=True) (call sites scene_encoder/attn_fusion.py:107,109). nmax >= largest scene. */
int prosim_build_knn_edges(const float* qpos, const int32_t* qscene, int n_q, const float* spos, const int32_t* seg,
                           int k, int nmax, int32_t* nbr, int32_t* deg, int stride, prosim_stream_t stream);
/* _get_rel_pe + FourierEmbeddingFix + attn_prenorm_r statistics (act_decoder.py:203-221,
 * layers/fourier_embedding.py:56-79, attention_layer.py:68). extra (nullable): per-edge vector added before
 * the normalisation (condition edges, condition_transformer/condition_attns.py:211-216).
 * Also zero-fills z up to the next multiple of 8 entries past each list (at least 8 entries per row, never past the
 * row's stride): prosim_attn_layer_fwd / prosim_attn_stack_fwd read whole groups of 8 entries (with weight 0 beyond a
 * list) and need finite values there -- a caller that fills z itself must guarantee the same. */
int prosim_edge_pe(const float* dpos, const float* dori, int n_dst, const float* spos, const float* sori,
                   const int32_t* nbr, const int32_t* deg, int stride, const float* dim_t16, const float* extra,
                   int zd, float* z, prosim_stream_t stream);

/* AttentionLayer pieces (prosim/models/layers/attention_layer.py:56-118) */
int prosim_attn_kv(const float* x_src, int n_src, const float* w, size_t w_layer_stride, int n_layers, float* kv,
                   size_t kv_layer_stride, prosim_stream_t stream);
/* one full layer: x_src == x_dst (same pointer) selects the non-bipartite form */
int prosim_attn_layer_fwd(const float* x_src, int n_src, const float* x_dst, int n_dst, const prosim_graph_t* g,
                          const float* w, float* workspace, size_t workspace_floats, float* out,
                          prosim_stream_t stream);
/* n_layers x (layer A [, layer B]) on the same destination rows, each post kernel fused with the next layer's
 * destination-side projections: the policy's a2p/m2p loop (act_decoder.py:270-277), the generator's p2p/s2p loop
 * (sym_coord.py:100-103) and the goal-condition GNN (condition_attns.py:222-224; side_b == NULL). */
int prosim_attn_stack_fwd(const float* x, int n_dst, int n_layers, const prosim_stack_side_t* side_a,
                          const prosim_stack_side_t* side_b, float* workspace, size_t workspace_floats, float* out,
                          prosim_stream_t stream);

/* ActDecoder._compute_traj (policy/act_decoder.py:78-135): motion_pred [P][10][5].
 * noise (may be NULL): [P][10][2] standard-normal draws; noise * noise_std is added to the per-step (dx, dy) before the
 * cumulative sum (RANDOM_NOISE_STD > 0, act_decoder.py:113-115).  The draws come from the caller's generator. */
int prosim_policy_head_fwd(const float* feat, const int32_t* agent_type, int P, const float* w, const float* noise,
                           float noise_std, float* motion_pred, prosim_stream_t stream);
/* pred_mlp(policy emd) (act_decoder.py:129-131): out [P][2] */
int prosim_reconst_fwd(const float* emd, int P, const float* w, float* out, prosim_stream_t stream);
/* PromptEncoder (prompt_encoder/base.py:30,37-50) / GoalConditionEncoder (condition_encoders.py:21-51) */
int prosim_mlp2_fwd(const float* in, int ld_in, int k0, int n, int use_ln, const float* w, const float* tpe_t,
                    int tpe_ld, const float* dim_t128, float* out, prosim_stream_t stream);
/* MODEL.OBS_UPDATE.FUSION = 'mlp' (scene_encoder/attn_fusion.py:177-203): x_new[idx_new[i]] <- obs_update_mlp([x_old[idx_old[i]] |
 * x_new[idx_new[i]]]) for the n agents observed at the previous tick too, in place (w: prosim_obs_fuse_floats() floats). */
int prosim_obs_fuse_floats(void);
int prosim_obs_fuse_fwd(const float* x_old, const int32_t* idx_old, float* x_new, const int32_t* idx_new, int n,
                        const float* w, prosim_stream_t stream);
/* MotionTagEncoder for unary action tags (condition_transformer/condition_encoders.py:76-145): tags int64 [n][3] =
 * (tag id, start step, end step); table [16][128] indexed by tag id; out [n][128] = tag vector + FourierEmbeddingFix(64)
 * of start | end.  Rows with an id outside [0, n_tags) become zeros. */
int prosim_tag_embed_fwd(const int64_t* tags, int n, int n_tags, const float* table, const float* dim_t64, float* out,
                         prosim_stream_t stream);
/* GNNConditionAttn edge matrix + mean pooling for unary conditions (condition_transformer/condition_attns.py:114-189):
 * slot int32 [P][n_slots] = row of emb per (policy row, condition type) or -1; extra [P][128], has int32 [P]. */
int prosim_cond_pool_fwd(const float* emb, const int32_t* slot, int P, int n_slots, float* extra, int32_t* has,
                         prosim_stream_t stream);

/* ProSim.init_agent_trajs (traj_sam.py:597-633) */
int prosim_init_traj(const float* obs_in, const float* obs_pos, const float* obs_head, const int32_t* p_slot,
                     const int32_t* p_row, int P, int T, float* traj, float* vel, float* init_pos,
                     float* init_heading, prosim_stream_t stream);
/* ProSim.step_env (traj_sam.py:205-274); fut_* may be NULL on the first tick */
int prosim_step_env(const float* traj, const float* vel, const float* init_pos, const float* init_heading,
                    const int32_t* p_row, const int32_t* p_slot, int P, int T, int tidx, float* p_pos, float* p_ori,
                    float* fut_in, uint8_t* fut_mask, float* fut_pos, float* fut_head, prosim_stream_t stream);
int prosim_gather_pose(const float* pos, const float* head, const int32_t* rows, int n, float* out_pos,
                       float* out_ori, prosim_stream_t stream);
/* ProSim.step_agent_traj (traj_sam.py:276-349), TOP_K = 1 */
int prosim_step_agent_traj(const float* motion_pred, const int32_t* p_row, int P, int T, int tidx, float* traj,
                           float* vel, prosim_stream_t stream);

/* obtain_rollout_trajs_in_world (rollout/gpu_utils.py:230-266) + batch_nd_transform_points_pt / angles_pt
 * (rollout/utils.py:347-392): agent-t0-frame trajectories -> world (x, y, heading); tf = 3x3 row-major, out [P][steps][3] */
int prosim_rollout_to_world(const float* traj, const float* init_pos, const float* init_heading, const int32_t* p_row,
                            int P, int T, int t0, int steps, const float* tf, float* out, prosim_stream_t stream);

/* ---- one whole policy tick in ONE call (reference loop body traj_sam.py:159-170 ->
 * policy/temporal_ar.py:75-92 -> act_decoder.py:239-279 attn_fuse + :78-135 _compute_traj [-> traj_sam.py:276-349]).
 * Enqueues, on `stream`: the two radius graphs (policy rows -> agent tokens, policy rows -> map tokens), their relative
 * PE, the agent-side K'|V' of all layers, n_layers x (a2p, m2p) attention layers, the MCG head, and -- if `traj` is given
 * -- the state update.  Same kernels and the same bits as the per-step entry points above; what it saves is ~10 host
 * round trips through the binding per tick.  All scratch (neighbour lists, z, K'|V' of the agents, the attention
 * workspace) is carved from ONE caller-provided workspace of prosim_workspace_bytes(&cfg) bytes (256-byte aligned). */
typedef struct {
  int32_t n_policy_rows;        /* P                                                                                */
  int32_t n_agent_tokens;       /* agent tokens of this tick (all scenes)                                           */
  int32_t n_map_tokens;         /* map tokens (all scenes)                                                          */
  int32_t max_agents_per_scene; /* bounds the agent neighbour-list stride: min(max_neigh, this)                     */
  int32_t max_map_per_scene;    /* bounds the map neighbour-list stride                                             */
  int32_t max_neigh;            /* MODEL.POLICY.ACT_DECODER.ATTN.MAX_NUM_NEIGH (768)                                 */
  int32_t n_layers;             /* MODEL.POLICY.ACT_DECODER.ATTN.NUM_LAYER (6)                                       */
} prosim_cfg_t;
size_t prosim_workspace_bytes(const prosim_cfg_t* cfg);

typedef struct {
  prosim_cfg_t cfg;
  /* policy rows (one per controlled agent) */
  const float* emd;             /* [P][128] policy tokens (SymCoordDecoder output rows)                             */
  const int32_t* agent_type;    /* [P] in {1,2,3}                                                                    */
  const int32_t* p_scene;       /* [P] scene of each row                                                             */
  const float* p_pos;           /* [P][2] current world pose of each row (prosim_step_env output)                   */
  const float* p_ori;           /* [P]                                                                               */
  /* agent tokens of this tick: scene_tokens rows of type 1 (attn_fusion.py:205-236) */
  const float* x_agent;         /* [n_agent_tokens][128]                                                             */
  const float* agent_pos;       /* [n_agent_tokens][2]                                                               */
  const float* agent_ori;       /* [n_agent_tokens]                                                                  */
  const int32_t* seg_agent;     /* [B][4] {start, len, 0, 0} token range of each scene                               */
  /* map tokens: never change during a rollout */
  const float* map_pos;         /* [n_map_tokens][2]                                                                 */
  const float* map_ori;         /* [n_map_tokens]                                                                    */
  const int32_t* seg_map;       /* [B][4]                                                                            */
  const float* kv_map;          /* [n_layers][n_map_tokens][256] prosim_attn_kv(map tokens, m2p weights), once per scene */
  /* weights (packed arena sections) */
  const float* w_a2p;           /* n_layers x prosim_attn_layer_floats()                                             */
  const float* w_m2p;
  const float* w_head;          /* prosim_head_floats()                                                              */
  const float* dim_t16;         /* FourierEmbeddingFix denominators                                                  */
  float agent_radius, map_radius;
  const float* noise;           /* nullable [P][10][2] standard-normal draws                                        */
  float noise_std;
  /* outputs */
  float* fuse;                  /* [P][128] attn_fuse output (kept for callers that want the features)              */
  float* motion_pred;           /* [P][10][5]                                                                        */
  /* optional state update (prosim_step_agent_traj); traj == NULL skips it */
  const int32_t* p_row;
  int32_t T, tidx;
  float* traj;
  float* vel;
} prosim_tick_t;
int prosim_policy_tick(const prosim_tick_t* t, void* workspace, size_t workspace_bytes, prosim_stream_t stream);

/* tcgen05 / TMEM building block under validation (csrc/tc_gemm.cuh): c[m][128] = a[m][128] * w[128][128]^T with
 * kind::tf32 tensor-core MMAs, accumulators in tensor memory; split3 = 1 selects the fp32-class 3xTF32 scheme. */
int prosim_tc_gemm_test(const float* a, const float* w, float* c, int m, int split3, prosim_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
