/* libpilegnn -- C ABI of the B200-native particle-GNN rollout path.
 *
 * The reference (WangYixuan12/dyn-res-pile-manip) has no FFI layer: its boundary for this path is the
 * Python class API.  These entry points are what a maintainer binds underneath that API
 * (ctypes stub in INTEGRATION.md); each one names the reference code it replaces.
 *
 * Conventions: every pointer is a DEVICE pointer unless stated; float32 / int32, row-major, contiguous;
 * `stream` is a cudaStream_t passed as void*; nothing allocates, nothing synchronises (except the
 * measurement hook), no global state except one-time kernel attribute setup and the pile_set_tensor_cores
 * switch.  Return value: 0 = ok, otherwise a cudaError_t value
 * (cudaErrorInvalidValue for rejected arguments).  All launches are CUDA-graph capturable.
 */
#ifndef PILE_GNN_H
#define PILE_GNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PILE_ABI_VERSION 4

int pile_abi_version(void);
int pile_nf_effect(void);        /* hidden width the planner engines are compiled for (config train.particle.nf_effect
                                  * = 64); other widths run on the general-width engine, pile_general_* below */
int pile_max_relations(void);    /* 10, model/gnn_dyn.py:231 */
const char* pile_error_string(int code);

/* GEMM engine: 0 = FP32 CUDA-core tiles (the parity anchor), 1 = tcgen05/TMEM tensor-core tiles with bf16
 * hi/lo split operands (3 passes, fp32 accumulate) and shared-memory activation tiles, 2 = as 1 with the
 * relation encoder's activation operand kept in tensor memory.  Process-wide; returns the previous setting. */
int pile_set_tensor_cores(int enable);
int pile_get_tensor_cores(void);

/* measurement hook: when device_buf != NULL, one warp of the tcgen05 relation-encoder kernel records
 * (tag << 56 | clock64) stamps of its per-layer phases into device_buf[0..capacity). NULL disables. */
int pile_debug_set_trace(long long* device_buf, int capacity, int which /*0 relation encoder, 1 particle kernels*/);

/* ---- packed weights ---------------------------------------------------------------------------
 * The host packs the 18 checkpoint tensors into one float buffer; slots are listed in
 * csrc/common.cuh (enum WSlot): transposed [in][out] blocks for the forward, [out][in] for the dgrad. */
int pile_wpack_num_slots(void);
long long pile_wpack_slot_offset(int slot);   /* in floats */
long long pile_wpack_slot_size(int slot);
long long pile_wpack_total(void);

/* ---- pusher model: replaces PlannerGD.world2cam + gen_s_delta (planners.py:192-257) and, for the real robot
 * (env.is_real), gen_s_delta_irl (planners.py:259-300; dispatch at :345-348) -----------------------------------
 * action[b] = (sx, sy, ex, ey) at action + b*act_stride.  The frame of the push is described by a HOST struct:
 *   kind 0 (simulator): end points = cam_m12 * (a_x, 0, -a_y, 1) / global_scale with cam_m12 = rows 0..2 of the
 *          4x4 world->camera matrix the reference rebuilds per call (planners.py:197-203); half width 0.8/24;
 *   kind 1 (real robot): end points = (a_x / s2r_scale, -a_y / s2r_scale, 0.88), particles shifted by the
 *          workspace centre (wkspc_center_x, wkspc_center_y, 0) before the projection; half width 0.048. */
#define PILE_PUSHER_SIM 0
#define PILE_PUSHER_REAL 1
typedef struct pile_pusher {
  int kind;
  float cam_m12[12];
  float global_scale;
  float s2r_scale;
  float wkspc_center_x, wkspc_center_y;
} pile_pusher;
int pile_gen_s_delta(const float* s_cur, const float* action, int act_stride, const pile_pusher* pusher /*HOST*/,
                     int B, int N, float* s_delta, void* stream);

/* ---- relation construction: replaces model/gnn_dyn.py:221-251 ------------------------------------
 * rowptr [B, N+1] (offsets local to the sample), col/row [B, 10*N] sender / receiver index; relations of
 * a sample are sorted by (receiver, sender) = torch.nonzero() order.  trowptr/trecv/tedge (nullable) are
 * the sender-major transpose used by the backward.  particle_nums (nullable) [B] = padding mask. */
int pile_build_relations(const float* s_cur, const float* s_delta, const int* particle_nums, int B, int N,
                         float adj_thresh, int* rowptr, int* col, int* row, int* trowptr, int* trecv,
                         int* tedge, void* stream);

/* ---- one model step: replaces PropNetDiffDenModel.predict_one_step (model/gnn_dyn.py:209-254) ----
 * scratch: pile_step_scratch_bytes(B,N) bytes; tape (nullable): pile_tape_step_bytes(B,N) bytes that
 * receive the relation lists + ReLU sign bits needed by pile_step_backward. */
long long pile_step_scratch_bytes(int B, int N);
long long pile_tape_step_bytes(int B, int N);
int pile_predict_step(const float* wpack, const float* attr, const float* dens, const int* particle_nums,
                      const float* s_cur, const float* s_delta, float adj_thresh, int B, int N, void* scratch,
                      void* tape, float* s_pred, void* stream);
/* forward on caller-provided relation lists -- the "Rr/Rs-equivalent" entry of PropModuleDiffDen.forward
 * (model/gnn_dyn.py:147): relations of a sample must be grouped by receiver (CSR), at most pile_max_relations()
 * per receiver (what the reference's own builder emits, gnn_dyn.py:231); the caller checks this. */
int pile_forward_relations(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                           const float* s_delta, const int* rowptr, const int* col, const int* row, int B, int N,
                           void* scratch, void* tape, float* s_pred, void* stream);
/* backward of one recorded step (dgrad only): g_pred [B,N,3] -> g_s_cur, g_s_delta [B,N,3] (overwritten) */
int pile_step_backward(const float* wpack, const float* dens, const void* tape, int B, int N, const float* g_pred,
                       float* g_s_cur, float* g_s_delta, void* bwd_scratch, void* stream);
/* backward of pile_gen_s_delta: g_s_cur += d/ds_cur, g_action[b] (at g_action + b*g_act_stride) = d/daction */
int pile_gen_s_delta_backward(const float* s_cur, const float* action, int act_stride, const pile_pusher* pusher,
                              int B, int N, const float* g_s_delta, float* g_s_cur, float* g_action,
                              int g_act_stride, void* stream);
/* read the relation lists of a step back out of scratch (tape == NULL run) or tape */
int pile_relations_view(void* scratch_or_tape, int is_tape, int B, int N, int** rowptr, int** col, int** row);

/* ---- horizon rollout: replaces PlannerGD.ptcl_model_rollout (planners.py:302-370) ---------------
 * attr [B,N], dens [B], s0 [B,N,3], actions [B,T,4] -> states [B,T,N,3]; B is the already tiled
 * sample*state-variant batch (flat index = sample*n_batch + b).  tape (nullable): T * tape_step_bytes. */
int pile_rollout_forward(const float* wpack, const float* attr, const float* dens, const float* s0,
                         const float* actions, const pile_pusher* pusher, float adj_thresh, int B, int N, int T,
                         void* scratch, void* tape, float* states, void* stream);

/* measurement hook for bench.py's roofline line: runs one rollout step `reps` times with CUDA events
 * around each of its 9 kernels (relation search, particle encoder, relation encoder, then per propagation step the
 * receiver-segmented sum k_edge_agg and the particle update; the FP32 engine fuses the last two: its slots 3/5/7
 * read 0) on `stream`, SYNCHRONISES, and writes the mean milliseconds per kernel to ms_out (HOST, 9 floats). */
int pile_profile_step(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                      const float* action, int act_stride, const pile_pusher* pusher, float adj_thresh, int B,
                      int N, void* scratch, float* s_out, int reps, float* ms_out, void* stream);

/* backward of the rollout w.r.t. the actions (dgrad only; relation sets and the hard along-push mask
 * carry no gradient, as in autograd: planners.py:742-745).  g_states [B,T,N,3] = dL/dstates (consumed,
 * overwritten), g_actions [B,T,4] out.  bwd_scratch: pile_bwd_scratch_bytes(B,N). */
long long pile_bwd_scratch_bytes(int B, int N);
int pile_rollout_backward(const float* wpack, const float* dens, const float* s0, const float* actions,
                          const pile_pusher* pusher, int B, int N, int T, const void* tape, const float* states,
                          float* g_states, void* bwd_scratch, float* g_actions, void* stream);

/* ---- reward: replaces config_reward_ptcl via ptcl_evaluate_traj (env/flex_rewards.py:156-214) ----
 * states: n_states blocks of [N,3], state_stride floats apart; goal_img [Hh,Ww] is the SHAPED image
 * (goal - DT - min, computed once per goal on the host with cv2); goal_coor [M,2] = (col,row).
 * argmin (nullable) [n_states, M] int32 receives the nearest-particle index for the backward. */
int pile_reward(const float* states, long long n_states, long long state_stride, int N, const float* goal_img,
                int Hh, int Ww, const float* goal_coor, int M, const float* cam_params4 /*HOST fx,fy,cx,cy*/,
                float off_x, float off_y, int normalize, float* reward, int* argmin, void* stream);
int pile_reward_backward(const float* states, long long n_states, long long state_stride, int N,
                         const float* goal_img, int Hh, int Ww, const float* goal_coor, int M,
                         const float* cam_params4, float off_x, float off_y, int normalize,
                         const float* g_reward, const int* argmin, float* g_states, long long g_stride,
                         int accumulate, void* stream);

/* ---- action update: replaces optimizer.step() + the clamp_ calls of the GD planner (planners.py:674,
 * 742-764).  actions/grad/exp_avg/exp_avg_sq: n floats laid out [..., 4] = (sx, sy, ex, ey); torch.optim.Adam's
 * update for 1-based `step`, then clamp component c to [lo4[c], hi4[c]] (HOST pointers). */
int pile_adam_clamp(float* actions, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, int step,
                    float lr, float beta1, float beta2, float eps, const float* lo4, const float* hi4, void* stream);
/* Same update with the 1-based step number read from DEVICE memory (step = *iter_dev + 1), so that one captured
 * CUDA graph of a planner iteration can be replayed n_iter times; pile_counter_add advances the counter on the
 * stream (planners.py:682 `for i in range(n_iter)`). */
int pile_adam_clamp_dev(float* actions, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                        const int* iter_dev, float lr, float beta1, float beta2, float eps, const float* lo4,
                        const float* hi4, void* stream);
int pile_counter_add(int* counter_dev, int delta, void* stream);
/* Per-iteration bookkeeping of the GD planner on the device (planners.py:721-740): reward [n_sample*n_batch] with
 * row = sample*n_batch + b, actions [rows, T, 4] BEFORE the update.  For every state variant b: the best sample of
 * this iteration (lowest index on ties); if it beats max_reward[b] (strictly) it replaces max_reward[b],
 * max_idx[b] and best_actions[b, T, 4].  Statistics: for every scene k (state variants k*stat_every ..., one scene =
 * stat_every variants; a single planner call has stat_every = n_batch) rew_mean[k*stat_stride + it] /
 * rew_std[...] (it = *iter_dev) receive the mean and the unbiased standard deviation of reward[:, k*stat_every]
 * over the samples (planners.py:737-738 reads reward_seqs[:, 0]). */
int pile_gd_track(const float* reward, const float* actions, int n_sample, int n_batch, int T, float* max_reward,
                  int* max_idx, float* best_actions, float* rew_mean, float* rew_std, const int* iter_dev,
                  int stat_every, int stat_stride, void* stream);

/* ---- farthest-point sampling: replaces utils.fps_np (utils.py:451-466) as used for the goal pixels
 * (planners.py:620-624).  pts [n_sets, n, dim] (dim <= 3); per set: start at init_idx, take the farthest
 * point `count` times (first index on ties).  gap_workspace [n_sets, n] floats; out_idx [n_sets, count],
 * out_pts [n_sets, count, dim], out_radius (nullable) [n_sets] = covering radius after all picks (the
 * dist.max() that fps_np returns). */
int pile_fps(const float* pts, int n_sets, int n, int dim, int count, int init_idx, float* gap_workspace,
             int* out_idx, float* out_pts, float* out_radius, void* stream);

/* Same sampler with one start index per set (device int[n_sets]); shared_cloud != 0: every set samples the same
 * [n, dim] cloud; squared != 0: compare squared distances (dgl.geometry.farthest_point_sampler, used by utils.fps,
 * utils.py:423-437) instead of norms (fps_np).  out_radius then holds the squared gap. */
int pile_fps_sets(const float* pts, int shared_cloud, int n_sets, int n, int dim, int count, const int* init_idx,
                  int squared, float* gap_workspace, int* out_idx, float* out_pts, float* out_radius, void* stream);

/* ---- observation -> particles: the planner-side work of an MPC step before the rollout
 * (env/flex_env.py:910-951 obs2ptcl_fixed_num_batch, called with batch_size 30 at :1028 and :1086) -------------
 * pile_depth_to_points: utils.depth2fgpcd (utils.py:491-506) with mask = (0 < depth < max_depth) as at
 * flex_env.py:927.  depth [H,W] float32 (already divided by global_scale), cam4 = HOST doubles (fx, fy, cx, cy);
 * out_pts [capacity,3] float64 in row-major pixel order, *n_out (device) = number of foreground pixels,
 * counts_workspace: pile_depth_counts_len(H, W) ints. */
int pile_depth_counts_len(int H, int W);
int pile_depth_to_points(const float* depth, int H, int W, const double* cam4, float max_depth, double* out_pts,
                         int capacity, int* n_out, int* counts_workspace, void* stream);
/* pile_voxel_downsample: utils.downsample_pcd (utils.py:533-544) = open3d PointCloud::VoxelDownSample: voxel index
 * floor((p - (min_bound - voxel/2)) / voxel), one output point per occupied voxel = mean of its points (summed in
 * input order).  Output order: ascending (ix, iy, iz) (open3d's hash-map order is unspecified).  pts [n,3] float64,
 * out_pts [n,3] capacity, *m_out (device) = number of voxels, workspace pile_voxel_downsample_bytes(n) bytes. */
long long pile_voxel_downsample_bytes(int n);
int pile_voxel_downsample(const double* pts, int n, double voxel_size, double* out_pts, int* m_out, void* workspace,
                          void* stream);
/* pile_cover_radius: particle_r of utils.fps (utils.py:435-437): max over cloud points of the distance to the nearest
 * pick.  cloud [m,3] float64, picks [n_sets,count,3] float32, radius [n_sets] float64. */
int pile_cover_radius(const double* cloud, int m, const float* picks, int n_sets, int count, double* radius,
                      void* stream);
/* pile_recenter: utils.recenter (utils.py:468-477) with r = min(r_cap, r_scale * radius[set]) (flex_env.py:949):
 * out[set][k] = mean of the cloud points closer than r to pick k, float32 like the reference's zeros_like(picks). */
int pile_recenter(const double* cloud, int m, const float* picks, int n_sets, int count, const double* radius,
                  double r_cap, double r_scale, float* out, void* stream);

/* ---- training step: replaces predict_one_step under autograd as driven by train/train_gnn_dyn.py:150-199 (Adam on
 * all 18 tensors).  pile_train_forward = the same model step (relation search incl. the particle_nums padding mask,
 * model/gnn_dyn.py:238-241) that leaves every layer input in train_tape (pile_train_tape_bytes(B, N) bytes);
 * pile_train_backward: g_pred [B,N,3] dense -> g_s_cur, g_s_delta [B,N,3] (overwritten) and the weight gradients
 * ACCUMULATED (+=) into `grads`: the 18 tensors of the model's state_dict, in its order and natural [out][in]
 * shapes, flattened into one float buffer; pile_train_grad_offset(i) = offset of tensor i (i = 18: total floats).
 * scratch: pile_train_scratch_bytes(B, N).  Sums run in fixed orders: deterministic. */
long long pile_train_tape_bytes(int B, int N);
long long pile_train_scratch_bytes(int B, int N);
long long pile_train_grad_offset(int tensor_index);
int pile_train_forward(const float* wpack, const float* attr, const float* dens, const int* particle_nums,
                       const float* s_cur, const float* s_delta, float adj_thresh, int B, int N, void* train_tape,
                       float* s_pred, void* stream);
/* the same step on caller-provided relation lists (receiver-grouped, any number of relations per receiver, at most
 * 10*N per sample): the training / general form of the Rr, Rs entry of PropModuleDiffDen.forward (gnn_dyn.py:147) */
int pile_train_forward_relations(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                                 const float* s_delta, const int* rowptr, const int* col, const int* row, int B, int N,
                                 void* train_tape, float* s_pred, void* stream);
int pile_train_backward(const float* wpack, const float* dens, void* train_tape, int B, int N, const float* g_pred,
                        float* g_s_cur, float* g_s_delta, float* grads, void* scratch, void* stream);
int pile_train_relations_view(void* train_tape, int B, int N, int** rowptr, int** col, int** row);

/* ---- general-width engine: the same model step for ANY hidden width nf_effect <= 256 (model/gnn_dyn.py:119 reads it
 * from the config; the engines above are compiled for 64).  Feature arrays and weight matrices are zero-padded to a
 * multiple of 64, the layers are block GEMMs on the CUDA cores.  wpack: the forward ([in][out]) and backward
 * ([out][in]) images of the 9 layers, padded, in the order of the non-tensor-core weight slots;
 * pile_general_wpack_slot_offset(slot, nf_effect) = float offset of slot 0..34, slot 35 = total floats.
 * pile_general_forward leaves every layer input in `tape` (pile_general_tape_bytes); pile_general_backward returns
 * d/ds_cur, d/ds_delta (overwritten) and, when grads != NULL, ACCUMULATES the 18 weight gradients into `grads`
 * (state_dict order and shapes, pile_general_grad_offset(i, nf_effect), i = 18: total) -- grads == NULL is the
 * planner's input-gradient-only case (planners.py:674).  Replaces predict_one_step / forward / their autograd for
 * checkpoints with nf_effect != 64 in the planner, the MPC loop and the training loop alike. */
long long pile_general_wpack_slot_offset(int slot, int nf_effect);
long long pile_general_tape_bytes(int B, int N, int nf_effect);
long long pile_general_scratch_bytes(int B, int N, int nf_effect);
long long pile_general_grad_offset(int tensor_index, int nf_effect);
int pile_general_forward(const float* wpack, int nf_effect, const float* attr, const float* dens,
                         const int* particle_nums, const float* s_cur, const float* s_delta, float adj_thresh, int B,
                         int N, void* tape, float* s_pred, void* stream);
/* The same step when NO backward pass follows (torch.no_grad(): the MPPI planner's rollouts, planners.py:302-370 under
 * `enable_grad=False`, plain predict_one_step calls, gnn_dyn.py:200-254): the relation propagator runs in its regrouped
 * form, Linear([r3, eff_r, eff_s, d]) = (W_e r3 + w_d d + b) + (W_r eff)[recv] + (W_s eff)[send] with the first term once per
 * step and the other two once per particle (2.7 x fewer FLOPs at E = 10 N).  `tape` has the size of pile_general_tape_bytes
 * and afterwards holds the relation lists (pile_general_relations_view) but not the layer inputs pile_general_backward
 * needs: do not call it on this tape. */
int pile_general_forward_inference(const float* wpack, int nf_effect, const float* attr, const float* dens,
                                   const int* particle_nums, const float* s_cur, const float* s_delta, float adj_thresh,
                                   int B, int N, void* tape, float* s_pred, void* stream);
int pile_general_forward_relations(const float* wpack, int nf_effect, const float* attr, const float* dens,
                                   const float* s_cur, const float* s_delta, const int* rowptr, const int* col,
                                   const int* row, int B, int N, void* tape, float* s_pred, void* stream);
int pile_general_backward(const float* wpack, int nf_effect, const float* dens, void* tape, int B, int N,
                          const float* g_pred, float* g_s_cur, float* g_s_delta, float* grads, void* scratch,
                          void* stream);
int pile_general_relations_view(void* tape, int B, int N, int nf_effect, int** rowptr, int** col, int** row);

/* ---- resolution regressor: replaces MPCResRgrNoPool.forward (model/res_regressor.py:106-144), the network that
 * picks the particle count once per MPC step (env/flex_env.py:981-998, 1080-1090).  params: all 20 tensors of the
 * module's state_dict in its own order (model.0.weight, model.0.bias, model.2.weight, ... model.19.bias) flattened
 * into one float buffer; pile_rgr_param_offset(2*l) / (2*l+1) = offset of layer l's weight / bias (l = 0..9),
 * pile_rgr_param_offset(20) = total floats (114 193 217).  x [B, 6, H, W] (H = W = 224: five stride-2 convolutions
 * must end at 7 x 7), y [B, 1].  workspace: pile_rgr_workspace_bytes(B, H, W) bytes (-1: unsupported size). */
long long pile_rgr_param_offset(int tensor_index);
long long pile_rgr_workspace_bytes(int B, int H, int W);
int pile_rgr_forward(const float* params, const float* x, int B, int H, int W, void* workspace, float* y,
                     void* stream);

/* ---- MPPI weighting: replaces PlannerGD.optimize_action (planners.py:549-561) -------------------
 * partials: [pile_mppi_num_chunks(S)][2 + 4T] = (max z, sum exp(z-max), sum exp(z-max)*act) with
 * z = reward_weight*reward; combine merges P such records (chunks and/or ranks) into one record
 * out[2+4T]; the optimised action sequence is out[2:] / out[1]. */
int pile_mppi_num_chunks(int S);
int pile_mppi_partials(const float* reward, const float* acts, int S, int T, float reward_weight,
                       float* partials, void* stream);
int pile_mppi_combine(const float* partials, int P, int T, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PILE_GNN_H */
