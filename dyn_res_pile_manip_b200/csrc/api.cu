// extern "C" surface of libpilegnn (declared in include/pile_gnn.h).
#include "../../include/pile_gnn.h"
#include "common.cuh"
#include "kernels.h"
#include <math.h>

using namespace pile;

namespace {

constexpr size_t ALIGN = 256;
inline size_t up(size_t x) { return (x + ALIGN - 1) / ALIGN * ALIGN; }

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += up(count * sizeof(T));
    return p;
  }
};

struct ScratchView {
  StepScratch ws;
  Csr csr;       // relation lists when no tape is recorded
  size_t bytes;
};

ScratchView carve_scratch(void* p, int B, int N) {
  Carver c(p);
  const size_t R = (size_t)B * N, E = (size_t)B * KMAX * N;
  // the tensor-core path stores features tile-blocked (tc_tile.cuh): round rows up to whole 128-row tiles,
  // relation tiles are per sample
  const size_t Rp = (R + TILE - 1) / TILE * TILE;
  const size_t Ep = (size_t)B * ((KMAX * N + TILE - 1) / TILE) * TILE;
  ScratchView v;
  v.ws.s_delta = c.take<float>(R * 3);
  v.ws.Ce = c.take<float>(Ep * H);
  v.ws.efeat = c.take<float>((E + TILE) * 8);
  v.ws.agg = c.take<float>(Rp * H);
  v.ws.Cp = c.take<float>(Rp * H);
  v.ws.eff = c.take<float>(Rp * H);
  v.ws.Pr[0] = c.take<float>(Rp * H);
  v.ws.Pr[1] = c.take<float>(Rp * H);
  v.ws.Ps[0] = c.take<float>(Rp * H);
  v.ws.Ps[1] = c.take<float>(Rp * H);
  v.csr.rowptr = c.take<int>((size_t)B * (N + 1));
  v.csr.col = c.take<int>(E);
  v.csr.row = c.take<int>(E);
  v.csr.trowptr = nullptr;
  v.csr.trecv = nullptr;
  v.csr.tedge = nullptr;
  v.bytes = c.off;
  return v;
}

struct TapeView {
  Csr csr;
  Masks mk;
  size_t bytes;
};

TapeView carve_tape(void* p, int B, int N) {
  Carver c(p);
  const size_t R = (size_t)B * N, E = (size_t)B * KMAX * N;
  TapeView v;
  v.csr.rowptr = c.take<int>((size_t)B * (N + 1));
  v.csr.col = c.take<int>(E);
  v.csr.row = c.take<int>(E);
  v.csr.trowptr = c.take<int>((size_t)B * (N + 1));
  v.csr.trecv = c.take<int>(E);
  v.csr.tedge = c.take<int>(E);
  v.mk.pe0 = c.take<uint8_t>(R * 8);
  v.mk.pe1 = c.take<uint8_t>(R * 8);
  for (int p2 = 0; p2 < PSTEP; ++p2) v.mk.eff[p2] = c.take<uint8_t>(R * 8);
  v.mk.q = c.take<uint8_t>(R * 8);
  v.mk.re0 = c.take<uint8_t>(E * 8);
  v.mk.re1 = c.take<uint8_t>(E * 8);
  v.mk.re2 = c.take<uint8_t>(E * 8);
  for (int p2 = 0; p2 < PSTEP; ++p2) v.mk.edge[p2] = c.take<uint8_t>(E * 8);
  v.bytes = c.off;
  return v;
}

PushCam make_cam(const pile_pusher* p) {
  PushCam c{};
  for (int i = 0; i < 12; ++i) c.m[i] = p->cam_m12[i];
  c.global_scale = p->global_scale;
  c.decay = 0.01f;             // planners.py:251 / :294
  c.kind = p->kind;
  if (p->kind == PILE_PUSHER_REAL) {
    c.pusher_w = 0.048f;       // planners.py:279
    c.s2r_scale = p->s2r_scale;
    c.shift_x = p->wkspc_center_x;   // planners.py:271-272
    c.shift_y = p->wkspc_center_y;
    c.height = 0.88f;          // planners.py:276
  } else {
    c.pusher_w = 0.8f / 24.0f; // planners.py:225
    c.s2r_scale = 1.f;
    c.shift_x = c.shift_y = 0.f;
    c.height = 0.f;
  }
  return c;
}

inline bool bad_pusher(const pile_pusher* p) {
  if (!p) return true;
  if (p->kind == PILE_PUSHER_SIM) return !(p->global_scale != 0.f);
  if (p->kind == PILE_PUSHER_REAL) return !(p->s2r_scale != 0.f);
  return true;
}

inline bool bad_dims(int B, int N) { return B <= 0 || N <= 0 || nbr_smem_bytes(N) > 200 * 1024; }

}  // namespace

extern "C" {

int pile_abi_version(void) { return PILE_ABI_VERSION; }
int pile_nf_effect(void) { return H; }
int pile_max_relations(void) { return KMAX; }
const char* pile_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

int pile_set_tensor_cores(int enable) {
  return g_use_tensor_cores.exchange(enable < 0 ? 0 : (enable > 2 ? 2 : enable));
}
int pile_get_tensor_cores(void) { return g_use_tensor_cores.load(); }

int pile_debug_set_trace(long long* device_buf, int capacity, int which) {
  if (which == 2) return set_edge_tmem_trace(device_buf, capacity);
  return which == 0 ? set_edge_trace(device_buf, capacity) : set_node_trace(device_buf, capacity);
}

int pile_wpack_num_slots(void) { return W_NUM; }
long long pile_wpack_slot_offset(int slot) { return (slot < 0 || slot > W_NUM) ? -1 : wslot_offset(slot); }
long long pile_wpack_slot_size(int slot) { return (slot < 0 || slot >= W_NUM) ? -1 : wslot_size(slot); }
long long pile_wpack_total(void) { return wslot_offset(W_NUM); }

int pile_gen_s_delta(const float* s_cur, const float* action, int act_stride, const pile_pusher* pusher, int B,
                     int N, float* s_delta, void* stream) {
  if (B <= 0 || N <= 0 || !s_cur || !action || bad_pusher(pusher) || !s_delta) return (int)cudaErrorInvalidValue;
  return launch_gen_s_delta(s_cur, (long long)N * 3, action, act_stride, make_cam(pusher), B, N, s_delta,
                            (cudaStream_t)stream);
}

int pile_build_relations(const float* s_cur, const float* s_delta, const int* particle_nums, int B, int N,
                         float adj_thresh, int* rowptr, int* col, int* row, int* trowptr, int* trecv,
                         int* tedge, void* stream) {
  if (bad_dims(B, N) || !s_cur || !s_delta || !rowptr || !col || !row) return (int)cudaErrorInvalidValue;
  if (trowptr && (!trecv || !tedge)) return (int)cudaErrorInvalidValue;
  PushCam none{};
  Csr csr{rowptr, col, row, trowptr, trecv, tedge};
  return launch_nbr_search(s_cur, (long long)N * 3, s_delta, nullptr, 0, none, nullptr, particle_nums, B, N,
                           adj_thresh * adj_thresh, csr, (cudaStream_t)stream);
}

long long pile_step_scratch_bytes(int B, int N) {
  if (bad_dims(B, N)) return -1;
  return (long long)carve_scratch(nullptr, B, N).bytes;
}

long long pile_tape_step_bytes(int B, int N) {
  if (bad_dims(B, N)) return -1;
  return (long long)carve_tape(nullptr, B, N).bytes;
}

int pile_relations_view(void* p, int is_tape, int B, int N, int** rowptr, int** col, int** row) {
  if (!p || bad_dims(B, N)) return (int)cudaErrorInvalidValue;
  const Csr c = is_tape ? carve_tape(p, B, N).csr : carve_scratch(p, B, N).csr;
  *rowptr = c.rowptr; *col = c.col; *row = c.row;
  return 0;
}

int pile_predict_step(const float* wpack, const float* attr, const float* dens, const int* particle_nums,
                      const float* s_cur, const float* s_delta, float adj_thresh, int B, int N, void* scratch,
                      void* tape, float* s_pred, void* stream) {
  if (bad_dims(B, N) || !wpack || !attr || !dens || !s_cur || !s_delta || !scratch || !s_pred)
    return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  ScratchView sv = carve_scratch(scratch, B, N);
  TapeView tv;
  Csr csr = sv.csr;
  if (tape) { tv = carve_tape(tape, B, N); csr = tv.csr; }
  PushCam none{};
  const bool tc = g_use_tensor_cores != 0;      // the tensor engine takes its relation rows from the search kernel
  int e = launch_nbr_search(s_cur, (long long)N * 3, s_delta, nullptr, 0, none, nullptr, particle_nums, B, N,
                            adj_thresh * adj_thresh, csr, st, tc ? attr : nullptr, dens, tc ? sv.ws.efeat : nullptr);
  if (e) return e;
  return launch_forward(wpack, attr, dens, s_cur, (long long)N * 3, s_delta, csr, sv.ws, tape ? &tv.mk : nullptr,
                        s_pred, (long long)N * 3, B, N, st, nullptr, tc);
}

int pile_rollout_forward(const float* wpack, const float* attr, const float* dens, const float* s0,
                         const float* actions, const pile_pusher* pusher, float adj_thresh, int B, int N, int T,
                         void* scratch, void* tape, float* states, void* stream) {
  if (bad_dims(B, N) || T <= 0 || !wpack || !attr || !dens || !s0 || !actions || bad_pusher(pusher) || !scratch ||
      !states)
    return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  ScratchView sv = carve_scratch(scratch, B, N);
  const PushCam cam = make_cam(pusher);
  const size_t tape_step = carve_tape(nullptr, B, N).bytes;
  const long long sstride = (long long)T * N * 3;
  const bool tc = g_use_tensor_cores != 0;      // the tensor engine takes its relation rows from the search kernel
  for (int t = 0; t < T; ++t) {
    const float* s_cur = t == 0 ? s0 : states + (size_t)(t - 1) * N * 3;
    const long long cur_stride = t == 0 ? (long long)N * 3 : sstride;
    TapeView tv;
    Csr csr = sv.csr;
    if (tape) { tv = carve_tape(static_cast<char*>(tape) + (size_t)t * tape_step, B, N); csr = tv.csr; }
    int e = launch_nbr_search(s_cur, cur_stride, nullptr, actions + (size_t)t * 4, T * 4, cam, sv.ws.s_delta,
                              nullptr, B, N, adj_thresh * adj_thresh, csr, st, tc ? attr : nullptr, dens,
                              tc ? sv.ws.efeat : nullptr);
    if (e) return e;
    e = launch_forward(wpack, attr, dens, s_cur, cur_stride, sv.ws.s_delta, csr, sv.ws, tape ? &tv.mk : nullptr,
                       states + (size_t)t * N * 3, sstride, B, N, st, nullptr, tc);
    if (e) return e;
  }
  return 0;
}

int pile_forward_relations(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                           const float* s_delta, const int* rowptr, const int* col, const int* row, int B, int N,
                           void* scratch, void* tape, float* s_pred, void* stream) {
  if (bad_dims(B, N) || !wpack || !attr || !dens || !s_cur || !s_delta || !rowptr || !col || !row || !scratch ||
      !s_pred)
    return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  ScratchView sv = carve_scratch(scratch, B, N);
  Csr csr{const_cast<int*>(rowptr), const_cast<int*>(col), const_cast<int*>(row), nullptr, nullptr, nullptr};
  TapeView tv;
  if (tape) {
    tv = carve_tape(tape, B, N);
    const size_t E = (size_t)B * KMAX * N;
    cudaError_t e;
    if ((e = cudaMemcpyAsync(tv.csr.rowptr, rowptr, sizeof(int) * (size_t)B * (N + 1), cudaMemcpyDeviceToDevice, st))) return (int)e;
    if ((e = cudaMemcpyAsync(tv.csr.col, col, sizeof(int) * E, cudaMemcpyDeviceToDevice, st))) return (int)e;
    if ((e = cudaMemcpyAsync(tv.csr.row, row, sizeof(int) * E, cudaMemcpyDeviceToDevice, st))) return (int)e;
    csr = tv.csr;
    int r = launch_transpose_relations(csr, B, N, st);
    if (r) return r;
  }
  return launch_forward(wpack, attr, dens, s_cur, (long long)N * 3, s_delta, csr, sv.ws, tape ? &tv.mk : nullptr,
                        s_pred, (long long)N * 3, B, N, st);
}

int pile_gen_s_delta_backward(const float* s_cur, const float* action, int act_stride, const pile_pusher* pusher,
                              int B, int N, const float* g_s_delta, float* g_s_cur, float* g_action,
                              int g_act_stride, void* stream) {
  if (B <= 0 || N <= 0 || !s_cur || !action || bad_pusher(pusher) || !g_s_delta || !g_s_cur || !g_action)
    return (int)cudaErrorInvalidValue;
  return launch_gen_s_delta_bwd(s_cur, (long long)N * 3, action, act_stride, make_cam(pusher), B, N,
                                g_s_delta, g_s_cur, (long long)N * 3, g_action, g_act_stride, (cudaStream_t)stream);
}

namespace {
struct BwdView {
  void* kernels;      // launch_step_backward scratch
  float* g_cur;       // [B,N,3]
  float* g_sd;        // [B,N,3]
  size_t bytes;
};
BwdView carve_bwd_view(void* p, int B, int N) {
  Carver c(p);
  BwdView v;
  v.kernels = c.take<char>(bwd_scratch_bytes(B, N));
  v.g_cur = c.take<float>((size_t)B * N * 3);
  v.g_sd = c.take<float>((size_t)B * N * 3);
  v.bytes = c.off;
  return v;
}
}  // namespace

int pile_profile_step(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                      const float* action, int act_stride, const pile_pusher* pusher, float adj_thresh, int B,
                      int N, void* scratch, float* s_out, int reps, float* ms_out, void* stream) {
  if (bad_dims(B, N) || reps <= 0 || !wpack || !attr || !dens || !s_cur || !action || bad_pusher(pusher) ||
      !scratch || !s_out || !ms_out)
    return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  ScratchView sv = carve_scratch(scratch, B, N);
  const PushCam cam = make_cam(pusher);
  constexpr int NK = PROFILE_SLOTS;   // nbr, node_encode, edge_encode, 3 x (segmented sum, particle update)
  cudaEvent_t ev[NK + 1];
  for (auto& e : ev) cudaEventCreate(&e);
  for (int k = 0; k < NK; ++k) ms_out[k] = 0.f;
  int rc = 0;
  const bool tc = g_use_tensor_cores != 0;
  for (int r = 0; r < reps && !rc; ++r) {
    cudaEventRecord(ev[0], st);
    rc = launch_nbr_search(s_cur, (long long)N * 3, nullptr, action, act_stride, cam, sv.ws.s_delta, nullptr, B, N,
                           adj_thresh * adj_thresh, sv.csr, st, tc ? attr : nullptr, dens, tc ? sv.ws.efeat : nullptr);
    if (rc) break;
    rc = launch_forward(wpack, attr, dens, s_cur, (long long)N * 3, sv.ws.s_delta, sv.csr, sv.ws, nullptr, s_out,
                        (long long)N * 3, B, N, st, ev + 1, tc);
    if (rc) break;
    cudaEventSynchronize(ev[NK]);
    for (int k = 0; k < NK; ++k) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[k], ev[k + 1]);
      ms_out[k] += ms / reps;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

long long pile_bwd_scratch_bytes(int B, int N) {
  if (bad_dims(B, N)) return -1;
  return (long long)carve_bwd_view(nullptr, B, N).bytes;
}

int pile_step_backward(const float* wpack, const float* dens, const void* tape, int B, int N, const float* g_pred,
                       float* g_s_cur, float* g_s_delta, void* bwd_scratch, void* stream) {
  (void)dens;
  if (bad_dims(B, N) || !wpack || !tape || !g_pred || !g_s_cur || !g_s_delta || !bwd_scratch)
    return (int)cudaErrorInvalidValue;
  TapeView tv = carve_tape(const_cast<void*>(tape), B, N);
  BwdView bv = carve_bwd_view(bwd_scratch, B, N);
  return launch_step_backward(wpack, tv.csr, tv.mk, g_pred, (long long)N * 3, g_s_cur, g_s_delta, bv.kernels, B, N,
                              (cudaStream_t)stream);
}

int pile_rollout_backward(const float* wpack, const float* dens, const float* s0, const float* actions,
                          const pile_pusher* pusher, int B, int N, int T, const void* tape, const float* states,
                          float* g_states, void* bwd_scratch, float* g_actions, void* stream) {
  (void)dens;
  if (bad_dims(B, N) || T <= 0 || !wpack || !s0 || !actions || bad_pusher(pusher) || !tape || !states || !g_states ||
      !bwd_scratch || !g_actions)
    return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  const PushCam cam = make_cam(pusher);
  const size_t tape_step = carve_tape(nullptr, B, N).bytes;
  const long long sstride = (long long)T * N * 3;
  BwdView bv = carve_bwd_view(bwd_scratch, B, N);
  for (int t = T - 1; t >= 0; --t) {
    TapeView tv = carve_tape(const_cast<char*>(static_cast<const char*>(tape)) + (size_t)t * tape_step, B, N);
    const float* s_cur = t == 0 ? s0 : states + (size_t)(t - 1) * N * 3;
    const long long cur_stride = t == 0 ? (long long)N * 3 : sstride;
    int e = launch_step_backward(wpack, tv.csr, tv.mk, g_states + (size_t)t * N * 3, sstride, bv.g_cur, bv.g_sd,
                                 bv.kernels, B, N, st);
    if (e) return e;
    e = launch_gen_s_delta_bwd(s_cur, cur_stride, actions + (size_t)t * 4, T * 4, cam, B, N, bv.g_sd, bv.g_cur,
                               (long long)N * 3, g_actions + (size_t)t * 4, T * 4, st);
    if (e) return e;
    if (t > 0) {   // dL/ds_cur of step t joins the upstream gradient of step t-1's output
      e = launch_add_strided(g_states + (size_t)(t - 1) * N * 3, sstride, bv.g_cur, (long long)N * 3, B, N * 3, st);
      if (e) return e;
    }
  }
  return 0;
}

int pile_reward(const float* states, long long n_states, long long state_stride, int N, const float* goal_img,
                int Hh, int Ww, const float* goal_coor, int M, const float* cam, float off_x, float off_y,
                int normalize, float* reward, int* argmin, void* stream) {
  if (n_states < 0 || N <= 0 || M <= 0 || !states || !goal_img || !goal_coor || !cam || !reward)
    return (int)cudaErrorInvalidValue;
  return launch_reward(states, n_states, state_stride, N, goal_img, Hh, Ww, goal_coor, M, cam[0], cam[1], cam[2],
                       cam[3], off_x, off_y, normalize, reward, argmin, (cudaStream_t)stream);
}

int pile_reward_backward(const float* states, long long n_states, long long state_stride, int N,
                         const float* goal_img, int Hh, int Ww, const float* goal_coor, int M, const float* cam,
                         float off_x, float off_y, int normalize, const float* g_reward, const int* argmin,
                         float* g_states, long long g_stride, int accumulate, void* stream) {
  if (n_states < 0 || N <= 0 || M <= 0 || !states || !goal_img || !goal_coor || !cam || !g_reward || !argmin ||
      !g_states)
    return (int)cudaErrorInvalidValue;
  return launch_reward_bwd(states, n_states, state_stride, N, goal_img, Hh, Ww, goal_coor, M, cam[0], cam[1],
                           cam[2], cam[3], off_x, off_y, normalize, g_reward, argmin, g_states, g_stride,
                           accumulate, (cudaStream_t)stream);
}

int pile_fps(const float* pts, int n_sets, int n, int dim, int count, int init_idx, float* gap_workspace,
             int* out_idx, float* out_pts, float* out_radius, void* stream) {
  if (!pts || !gap_workspace || !out_idx || !out_pts) return (int)cudaErrorInvalidValue;
  return launch_fps(pts, n_sets, n, dim, count, init_idx, gap_workspace, out_idx, out_pts, out_radius,
                    (cudaStream_t)stream);
}

int pile_fps_sets(const float* pts, int shared_cloud, int n_sets, int n, int dim, int count, const int* init_idx,
                  int squared, float* gap_workspace, int* out_idx, float* out_pts, float* out_radius, void* stream) {
  if (!pts || !gap_workspace || !out_idx || !out_pts) return (int)cudaErrorInvalidValue;
  return launch_fps_sets(pts, shared_cloud, n_sets, n, dim, count, init_idx, squared, gap_workspace, out_idx, out_pts,
                         out_radius, (cudaStream_t)stream);
}

int pile_depth_counts_len(int H, int W) { return (int)(((long long)H * W + 1023) / 1024); }

int pile_depth_to_points(const float* depth, int H, int W, const double* cam4, float max_depth, double* out_pts,
                         int capacity, int* n_out, int* counts_workspace, void* stream) {
  if (!depth || !cam4 || !out_pts || !n_out || !counts_workspace) return (int)cudaErrorInvalidValue;
  return launch_depth_to_points(depth, H, W, cam4, max_depth, out_pts, capacity, n_out, counts_workspace,
                                (cudaStream_t)stream);
}

long long pile_voxel_downsample_bytes(int n) { return (long long)voxel_downsample_bytes(n); }

int pile_voxel_downsample(const double* pts, int n, double voxel_size, double* out_pts, int* m_out, void* workspace,
                          void* stream) {
  if (!pts || !out_pts || !m_out || !workspace) return (int)cudaErrorInvalidValue;
  return launch_voxel_downsample(pts, n, voxel_size, out_pts, m_out, workspace, (cudaStream_t)stream);
}

int pile_cover_radius(const double* cloud, int m, const float* picks, int n_sets, int count, double* radius,
                      void* stream) {
  if (!cloud || !picks || !radius) return (int)cudaErrorInvalidValue;
  return launch_cover_radius(cloud, m, picks, n_sets, count, radius, (cudaStream_t)stream);
}

int pile_recenter(const double* cloud, int m, const float* picks, int n_sets, int count, const double* radius,
                  double r_cap, double r_scale, float* out, void* stream) {
  if (!cloud || !picks || !radius || !out) return (int)cudaErrorInvalidValue;
  return launch_recenter(cloud, m, picks, n_sets, count, radius, r_cap, r_scale, out, (cudaStream_t)stream);
}

int pile_adam_clamp(float* actions, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, int step,
                    float lr, float beta1, float beta2, float eps, const float* lo4, const float* hi4, void* stream) {
  if (!actions || !grad || !exp_avg || !exp_avg_sq || n <= 0 || (n & 3) || step <= 0 || !lo4 || !hi4)
    return (int)cudaErrorInvalidValue;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  return launch_adam_clamp(actions, grad, exp_avg, exp_avg_sq, n, beta1, beta2, (float)((double)lr / bc1),
                           (float)sqrt(bc2), eps, lo4, hi4, (cudaStream_t)stream);
}

int pile_adam_clamp_dev(float* actions, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                        const int* iter_dev, float lr, float beta1, float beta2, float eps, const float* lo4,
                        const float* hi4, void* stream) {
  if (!actions || !grad || !exp_avg || !exp_avg_sq || n <= 0 || (n & 3) || !iter_dev || !lo4 || !hi4)
    return (int)cudaErrorInvalidValue;
  return launch_adam_clamp_dev(actions, grad, exp_avg, exp_avg_sq, n, iter_dev, lr, beta1, beta2, eps, lo4, hi4,
                               (cudaStream_t)stream);
}

int pile_counter_add(int* counter_dev, int delta, void* stream) {
  if (!counter_dev) return (int)cudaErrorInvalidValue;
  return launch_counter_add(counter_dev, delta, (cudaStream_t)stream);
}

int pile_gd_track(const float* reward, const float* actions, int n_sample, int n_batch, int T, float* max_reward,
                  int* max_idx, float* best_actions, float* rew_mean, float* rew_std, const int* iter_dev,
                  int stat_every, int stat_stride, void* stream) {
  if (!reward || !actions || n_sample <= 0 || n_batch <= 0 || T <= 0 || !max_reward || !max_idx || !best_actions ||
      !rew_mean || !rew_std || !iter_dev || stat_every <= 0 || stat_stride <= 0)
    return (int)cudaErrorInvalidValue;
  return launch_gd_track(reward, actions, n_sample, n_batch, T, max_reward, max_idx, best_actions, rew_mean, rew_std,
                         iter_dev, stat_every, stat_stride, (cudaStream_t)stream);
}

long long pile_train_tape_bytes(int B, int N) { return bad_dims(B, N) ? -1 : train_tape_bytes(B, N); }
long long pile_train_scratch_bytes(int B, int N) { return bad_dims(B, N) ? -1 : train_bwd_scratch_bytes(B, N); }
long long pile_train_grad_offset(int tensor_index) { return train_grad_offset(tensor_index); }

int pile_train_forward(const float* wpack, const float* attr, const float* dens, const int* particle_nums,
                       const float* s_cur, const float* s_delta, float adj_thresh, int B, int N, void* train_tape,
                       float* s_pred, void* stream) {
  if (bad_dims(B, N) || !wpack || !attr || !dens || !s_cur || !s_delta || !train_tape || !s_pred)
    return (int)cudaErrorInvalidValue;
  return launch_train_forward(wpack, attr, dens, particle_nums, s_cur, s_delta, adj_thresh, B, N, train_tape, s_pred,
                              (cudaStream_t)stream);
}

int pile_train_forward_relations(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                                 const float* s_delta, const int* rowptr, const int* col, const int* row, int B, int N,
                                 void* train_tape, float* s_pred, void* stream) {
  if (bad_dims(B, N) || !wpack || !attr || !dens || !s_cur || !s_delta || !rowptr || !col || !row || !train_tape || !s_pred)
    return (int)cudaErrorInvalidValue;
  return launch_train_forward_relations(wpack, attr, dens, s_cur, s_delta, rowptr, col, row, B, N, train_tape, s_pred,
                                        (cudaStream_t)stream);
}

int pile_train_backward(const float* wpack, const float* dens, void* train_tape, int B, int N, const float* g_pred,
                        float* g_s_cur, float* g_s_delta, float* grads, void* scratch, void* stream) {
  if (bad_dims(B, N) || !wpack || !dens || !train_tape || !g_pred || !g_s_cur || !g_s_delta || !grads || !scratch)
    return (int)cudaErrorInvalidValue;
  return launch_train_backward(wpack, dens, train_tape, B, N, g_pred, g_s_cur, g_s_delta, grads, scratch,
                               (cudaStream_t)stream);
}

int pile_train_relations_view(void* train_tape, int B, int N, int** rowptr, int** col, int** row) {
  if (!train_tape || bad_dims(B, N)) return (int)cudaErrorInvalidValue;
  return train_relations_view(train_tape, B, N, rowptr, col, row);
}

long long pile_general_wpack_slot_offset(int slot, int nf_effect) { return general_wpack_slot_offset(slot, nf_effect); }
long long pile_general_tape_bytes(int B, int N, int nf_effect) { return bad_dims(B, N) ? -1 : general_tape_bytes(B, N, nf_effect); }
long long pile_general_scratch_bytes(int B, int N, int nf_effect) {
  return bad_dims(B, N) ? -1 : general_bwd_scratch_bytes(B, N, nf_effect);
}
long long pile_general_grad_offset(int tensor_index, int nf_effect) { return general_grad_offset(tensor_index, nf_effect); }

int pile_general_forward(const float* wpack, int nf_effect, const float* attr, const float* dens,
                         const int* particle_nums, const float* s_cur, const float* s_delta, float adj_thresh, int B,
                         int N, void* tape, float* s_pred, void* stream) {
  if (bad_dims(B, N) || !wpack || !attr || !dens || !s_cur || !s_delta || !tape || !s_pred) return (int)cudaErrorInvalidValue;
  return launch_general_forward(wpack, nf_effect, attr, dens, particle_nums, s_cur, s_delta, adj_thresh, B, N, tape, s_pred,
                                (cudaStream_t)stream);
}

int pile_general_forward_inference(const float* wpack, int nf_effect, const float* attr, const float* dens,
                                   const int* particle_nums, const float* s_cur, const float* s_delta, float adj_thresh,
                                   int B, int N, void* tape, float* s_pred, void* stream) {
  if (bad_dims(B, N) || !wpack || !attr || !dens || !s_cur || !s_delta || !tape || !s_pred) return (int)cudaErrorInvalidValue;
  return launch_general_forward(wpack, nf_effect, attr, dens, particle_nums, s_cur, s_delta, adj_thresh, B, N, tape, s_pred,
                                (cudaStream_t)stream, true);
}

int pile_general_forward_relations(const float* wpack, int nf_effect, const float* attr, const float* dens,
                                   const float* s_cur, const float* s_delta, const int* rowptr, const int* col,
                                   const int* row, int B, int N, void* tape, float* s_pred, void* stream) {
  if (bad_dims(B, N) || !wpack || !attr || !dens || !s_cur || !s_delta || !rowptr || !col || !row || !tape || !s_pred)
    return (int)cudaErrorInvalidValue;
  return launch_general_forward_relations(wpack, nf_effect, attr, dens, s_cur, s_delta, rowptr, col, row, B, N, tape,
                                          s_pred, (cudaStream_t)stream);
}

int pile_general_backward(const float* wpack, int nf_effect, const float* dens, void* tape, int B, int N,
                          const float* g_pred, float* g_s_cur, float* g_s_delta, float* grads, void* scratch,
                          void* stream) {
  if (bad_dims(B, N) || !wpack || !dens || !tape || !g_pred || !g_s_cur || !g_s_delta || !scratch)
    return (int)cudaErrorInvalidValue;
  return launch_general_backward(wpack, nf_effect, dens, tape, B, N, g_pred, g_s_cur, g_s_delta, grads, scratch,
                                 (cudaStream_t)stream);
}

int pile_general_relations_view(void* tape, int B, int N, int nf_effect, int** rowptr, int** col, int** row) {
  if (!tape || bad_dims(B, N)) return (int)cudaErrorInvalidValue;
  return general_relations_view(tape, B, N, nf_effect, rowptr, col, row);
}

long long pile_rgr_param_offset(int tensor_index) {
  return (tensor_index < 0 || tensor_index > 20) ? -1 : rgr_param_offset(tensor_index);
}
long long pile_rgr_workspace_bytes(int B, int H, int W) { return rgr_workspace_bytes(B, H, W); }
int pile_rgr_forward(const float* params, const float* x, int B, int H, int W, void* workspace, float* y,
                     void* stream) {
  if (!params || !x || !workspace || !y) return (int)cudaErrorInvalidValue;
  return launch_rgr_forward(params, x, B, H, W, workspace, y, (cudaStream_t)stream);
}

int pile_mppi_num_chunks(int S) { return mppi_num_chunks(S); }

int pile_mppi_partials(const float* reward, const float* acts, int S, int T, float reward_weight, float* partials,
                       void* stream) {
  if (S <= 0 || T <= 0 || !reward || !acts || !partials) return (int)cudaErrorInvalidValue;
  return launch_mppi_partials(reward, acts, S, T, reward_weight, partials, (cudaStream_t)stream);
}

int pile_mppi_combine(const float* partials, int P, int T, float* out, void* stream) {
  if (P <= 0 || T <= 0 || !partials || !out) return (int)cudaErrorInvalidValue;
  return launch_mppi_combine(partials, P, T, out, (cudaStream_t)stream);
}

}  // extern "C"
