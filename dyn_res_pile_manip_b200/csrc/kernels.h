// Internal launch interface between the translation units of libpilegnn (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "common.cuh"

namespace pile {

// per-step relation structure of all samples
struct Csr {
  int* rowptr;   // [B, N+1]
  int* col;      // [B, KMAX*N] sender
  int* row;      // [B, KMAX*N] receiver
  int* trowptr;  // [B, N+1]    sender-major transpose (nullable)
  int* trecv;    // [B, KMAX*N]
  int* tedge;    // [B, KMAX*N]
};

// ReLU sign bits kept for the dgrad-only backward: one byte per (row, 8-column block)
struct Masks {
  uint8_t* pe0;        // [B*N, 8]
  uint8_t* pe1;
  uint8_t* eff[PSTEP]; // [B*N, 8]
  uint8_t* q;
  uint8_t* re0;        // [B*KMAX*N, 8]
  uint8_t* re1;
  uint8_t* re2;
  uint8_t* edge[PSTEP];
};

// scratch of one forward step (reused across steps)
struct StepScratch {
  float* s_delta;  // [B, N, 3]
  float* Ce;       // [B*KMAX*N, H]
  float* efeat;    // [B*KMAX*N + TILE, 8] relation input features (tensor-core path)
  float* agg;      // [B*N, H] receiver-aggregated relation effects (tensor-core path)
  float* Cp;       // [B*N, H]
  float* eff;      // [B*N, H]
  float* Pr[2];    // [B*N, H] ping-pong
  float* Ps[2];
};

int launch_gen_s_delta(const float* s_cur, long long s_stride, const float* action, int act_stride,
                       const PushCam& cam, int B, int N, float* s_delta, cudaStream_t st);
size_t nbr_smem_bytes(int N);
// s_delta either given (s_delta_in) or computed from `action` and written to s_delta_out
int launch_nbr_search(const float* s_cur, long long s_stride, const float* s_delta_in, const float* action,
                      int act_stride, const PushCam& cam, float* s_delta_out, const int* particle_nums, int B,
                      int N, float thr, const Csr& csr, cudaStream_t st,
                      // optional (tensor engine): also write the relation encoder's input rows to efeat
                      const float* attr = nullptr, const float* dens = nullptr, float* efeat = nullptr);

// forward of the propagation network on a prepared CSR; s_cur / s_out are [B, N, 3] with a per-sample
// stride in floats (so slices of [B, T, N, 3] work in place)
int launch_forward(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                   long long s_cur_stride, const float* s_delta, const Csr& csr, const StepScratch& ws,
                   const Masks* masks, float* s_out, long long s_out_stride, int B, int N, cudaStream_t st,
                   cudaEvent_t* ev = nullptr,    // ev: PROFILE_SLOTS + 1 events: before nbr-search's successors (node
                                                 // encode, relation encode, then per propagation step the segmented
                                                 // sum and the particle update) and after the last kernel
                   bool efeat_ready = false);    // ws.efeat already written by launch_nbr_search

// process-wide switch: 0 = FP32 CUDA-core tiles, 1 = tcgen05 tiles with shared-memory activations,
// 2 = 1 + relation encoder with the activation operand in tensor memory
extern std::atomic<int> g_use_tensor_cores;
int launch_edge_encode_tc(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                          long long s_stride, const Csr& csr, const Masks* mk, float* efeat, float* Ce, int B, int N,
                          cudaStream_t st, bool efeat_ready = false);

int launch_edge_encode_tmem(const float* wpack, const float* efeat, const Csr& csr, const Masks* mk, float* Ce,
                            int B, int N, cudaStream_t st);
int launch_bwd_edge_tc(const float* wpack, const Csr& csr, const Masks& mk, const float* ga0, const float* ga1,
                       const float* ga2, float* gx, int B, int N, cudaStream_t st);
int launch_bwd_head_tc(const float* wpack, const float* g_pred, long long g_stride, const uint8_t* m_q,
                       const uint8_t* m_eff2, float* gz, float* gcp, float* gagg2, int B, int N, cudaStream_t st);
int launch_bwd_prop_tc(const float* wpack, bool first, const float* gpr, const float* gps, const uint8_t* m_next,
                       const uint8_t* m_pe0, float* gz, float* gcp, float* gagg_out, float* g_s_delta, int B, int N,
                       cudaStream_t st);
int set_edge_trace(long long* buf, int cap);
int set_node_trace(long long* buf, int cap);
int set_edge_tmem_trace(long long* buf, int cap);
int launch_node_encode_tc(const float* wpack, const float* attr, const float* dens, const float* s_delta,
                          const Masks* mk, const StepScratch& ws, int B, int N, cudaStream_t st);
// propagation step p on the tensor-core path: k_edge_agg + k_node_update_tc (p == PSTEP-1: + predictor)
int launch_propagate_tc(const float* wpack, const Csr& csr, const StepScratch& ws, const Masks* mk, int p,
                        const float* s_cur, long long s_stride, float* s_out, long long o_stride, int B, int N,
                        cudaStream_t st, cudaEvent_t mid = nullptr /* recorded between the two kernels */);
// kernels timed by pile_profile_step: relation search, particle encoder, relation encoder, 3 x (k_edge_agg, particle update)
constexpr int PROFILE_SLOTS = 3 + 2 * PSTEP;

int launch_reward(const float* states, long long n_states, long long state_stride, int N, const float* goal_img,
                  int Hh, int Ww, const float* goal_coor, int M, float fx, float fy, float cx, float cy,
                  float off_x, float off_y, int normalize, float* reward, int* argmin_out, cudaStream_t st);
int launch_reward_bwd(const float* states, long long n_states, long long state_stride, int N, const float* goal_img,
                      int Hh, int Ww, const float* goal_coor, int M, float fx, float fy, float cx, float cy,
                      float off_x, float off_y, int normalize, const float* g_reward, const int* argmin_in,
                      float* g_states, long long g_stride, int accumulate, cudaStream_t st);
int launch_fps(const float* pts, int n_sets, int n, int dim, int count, int init_idx, float* gap_ws, int* out_idx,
               float* out_pts, float* out_radius, cudaStream_t st);
int launch_fps_sets(const float* pts, int shared_cloud, int n_sets, int n, int dim, int count, const int* init_idx,
                    int squared, float* gap_ws, int* out_idx, float* out_pts, float* out_radius, cudaStream_t st);

// observation -> particles (obs.cu)
int launch_depth_to_points(const float* depth, int H, int W, const double* cam4, float max_depth, double* out_pts,
                           int cap, int* n_out, int* ws_counts, cudaStream_t st);
size_t voxel_downsample_bytes(int n);
int launch_voxel_downsample(const double* pts, int n, double voxel, double* out_pts, int* m_out, void* ws,
                            cudaStream_t st);
int launch_cover_radius(const double* pcd, int m, const float* picks, int S, int N, double* radius, cudaStream_t st);
int launch_recenter(const double* pcd, int m, const float* picks, int S, int N, const double* radius, double r_cap,
                    double r_scale, float* out, cudaStream_t st);

int launch_adam_clamp(float* p, const float* g, float* m, float* v, long long n, float b1, float b2, float step_size,
                      float bc2_sqrt, float eps, const float* lo4, const float* hi4, cudaStream_t st);
int launch_adam_clamp_dev(float* p, const float* g, float* m, float* v, long long n, const int* iter_dev, float lr,
                          float b1, float b2, float eps, const float* lo4, const float* hi4, cudaStream_t st);
int launch_counter_add(int* counter, int delta, cudaStream_t st);
int launch_gd_track(const float* reward, const float* acts, int n_sample, int n_batch, int T, float* max_reward,
                    int* max_idx, float* best_actions, float* rew_mean, float* rew_std, const int* iter_dev,
                    int stat_every, int stat_stride, cudaStream_t st);
int mppi_num_chunks(int S);
int launch_mppi_partials(const float* reward, const float* acts, int S, int T, float weight, float* part,
                         cudaStream_t st);
int launch_mppi_combine(const float* part, int P, int T, float* out, cudaStream_t st);

// training path (train.cu): forward that keeps every layer input, backward with weight gradients
long long train_tape_bytes(int B, int N);
long long train_bwd_scratch_bytes(int B, int N);
long long train_grad_offset(int tensor_index);
int train_relations_view(void* tape, int B, int N, int** rowptr, int** col, int** row);
int launch_train_forward(const float* wpack, const float* attr, const float* dens, const int* particle_nums,
                         const float* s_cur, const float* s_delta, float adj_thresh, int B, int N, void* tape,
                         float* s_pred, cudaStream_t st);
int launch_train_forward_relations(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                                   const float* s_delta, const int* rowptr, const int* col, const int* row, int B,
                                   int N, void* tape, float* s_pred, cudaStream_t st);
int launch_train_backward(const float* wpack, const float* dens, void* tape, int B, int N, const float* g_pred,
                          float* g_s_cur, float* g_s_delta, float* grads, void* scratch, cudaStream_t st);

// general-width engine (general.cu): any nf_effect <= 256, forward with kept activations + full backward
long long general_wpack_slot_offset(int slot, int H);
long long general_tape_bytes(int B, int N, int H);
long long general_bwd_scratch_bytes(int B, int N, int H);
long long general_grad_offset(int tensor_index, int H);
int general_relations_view(void* tape, int B, int N, int H, int** rowptr, int** col, int** row);
int launch_general_forward(const float* wpack, int H, const float* attr, const float* dens, const int* particle_nums,
                           const float* s_cur, const float* s_delta, float adj_thresh, int B, int N, void* tape,
                           float* s_pred, cudaStream_t st, bool hoisted = false);
// tensor-core form of the wide relation-side layers of the general-width engine (general_tc.cu), inference only
namespace general {
size_t tc_image_floats(int Hp);
int launch_tc_images(const float* wpack, const long long* slot_off, int n, int Hp, float* img0, cudaStream_t st);
int launch_lin_tc_edge(const float* x, const float* img, const float* bias, const float* wd, const float* dens, int relu,
                       float* y, const int* rowptr, int B, int N, int Hp, cudaStream_t st, const float* x8 = nullptr,
                       const float* w8 = nullptr, const float* b8 = nullptr);
int launch_lin_tc_node(const float* x0, const float* img0, const float* x1, const float* img1, const float* bias,
                       const float* wd, const float* dens, const float* res, int relu, float* y, int B, int N, int Hp,
                       cudaStream_t st);
}  // namespace general
int launch_general_forward_relations(const float* wpack, int H, const float* attr, const float* dens, const float* s_cur,
                                     const float* s_delta, const int* rowptr, const int* col, const int* row, int B,
                                     int N, void* tape, float* s_pred, cudaStream_t st);
int launch_general_backward(const float* wpack, int H, const float* dens, void* tape, int B, int N, const float* g_pred,
                            float* g_s_cur, float* g_s_delta, float* grads, void* scratch, cudaStream_t st);

// resolution regressor (rgr.cu)
long long rgr_param_offset(int idx);
long long rgr_workspace_bytes(int B, int H, int W);
int launch_rgr_forward(const float* params, const float* x, int B, int H, int W, void* ws, float* y, cudaStream_t st);

size_t bwd_scratch_bytes(int B, int N);
// backward of one model step: g_pred [B,N,3] (strided) -> g_s_cur [B,N,3] dense (overwritten, includes the
// residual), g_s_delta [B,N,3] dense (overwritten)
int launch_step_backward(const float* wpack, const Csr& csr, const Masks& mk, const float* g_pred,
                         long long g_stride, float* g_s_cur, float* g_s_delta, void* scratch, int B, int N,
                         cudaStream_t st);
int launch_gen_s_delta_bwd(const float* s_cur, long long s_stride, const float* action, int act_stride,
                           const PushCam& cam, int B, int N, const float* g_sd, float* g_s_cur, long long g_stride,
                           float* g_action, int g_act_stride, cudaStream_t st);
int launch_transpose_relations(const Csr& csr, int B, int N, cudaStream_t st);
int launch_add_strided(float* dst, long long d_stride, const float* src, long long s_stride, int B, int per,
                       cudaStream_t st);

}  // namespace pile
