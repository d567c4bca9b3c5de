// K6/K7: particle-space target-shape reward and MPPI weighting.
//
// Reward (reference env/flex_rewards.py:156-214): project particles to pixels, sum the bilinear lookup of
// the shaped goal image (grid_sample, padding_mode='border', align_corners=False, BOTH axes normalised by
// H, :197) plus, for every goal point, the distance to the nearest projected particle (:207-209);
// divide by N, negate.  One CTA per state; the reference's [B, M, N] distance tensor and the B-fold tiled
// goal image never exist.
#include "common.cuh"
#include "kernels.h"

namespace pile {

constexpr int RW_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in warp 0
}

struct Bilinear {
  int x0, y0;
  float wx, wy;        // fractional parts
  float gx_scale, gy_scale;   // d(ix)/d(pix) (0 when the border clamp is active)
};

__device__ __forceinline__ Bilinear bilinear_setup(float px, float py, int Hh, int Ww) {
  // norm = pix / H * 2 - 1 ; ix = ((norm + 1) * W - 1) / 2 ; clamp to [0, size-1]
  const float nx = px / (float)Hh * 2.f - 1.f;
  const float ny = py / (float)Hh * 2.f - 1.f;
  float ix = ((nx + 1.f) * (float)Ww - 1.f) * 0.5f;
  float iy = ((ny + 1.f) * (float)Hh - 1.f) * 0.5f;
  Bilinear q;
  q.gx_scale = (ix > 0.f && ix < (float)(Ww - 1)) ? (float)Ww / (float)Hh : 0.f;
  q.gy_scale = (iy > 0.f && iy < (float)(Hh - 1)) ? 1.f : 0.f;
  ix = fminf(fmaxf(ix, 0.f), (float)(Ww - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(Hh - 1));
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  q.x0 = (int)fx0;
  q.y0 = (int)fy0;
  q.wx = ix - fx0;
  q.wy = iy - fy0;
  return q;
}

__device__ __forceinline__ float img_at(const float* __restrict__ img, int y, int x, int Hh, int Ww) {
  return (x >= 0 && x < Ww && y >= 0 && y < Hh) ? __ldg(img + (size_t)y * Ww + x) : 0.f;
}

__global__ void __launch_bounds__(RW_THREADS)
k_reward(const float* __restrict__ states, long long state_stride, int N, const float* __restrict__ goal_img,
         int Hh, int Ww, const float* __restrict__ goal_coor, int M, float fx, float fy, float cx, float cy,
         float off_x, float off_y, int normalize, float* __restrict__ reward, int* __restrict__ argmin_out) {
  extern __shared__ float sm[];
  float* pxs = sm;
  float* pys = sm + N;
  __shared__ float red[RW_THREADS / 32];
  const long long s = blockIdx.x;
  const float* st = states + s * state_stride;

  float part = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float x = st[n * 3 + 0], y = st[n * 3 + 1], z = st[n * 3 + 2];
    const float px = x * fx / z + cx + off_x;
    const float py = y * fy / z + cy + off_y;
    pxs[n] = px;
    pys[n] = py;
    const Bilinear q = bilinear_setup(px, py, Hh, Ww);
    const float v00 = img_at(goal_img, q.y0, q.x0, Hh, Ww), v01 = img_at(goal_img, q.y0, q.x0 + 1, Hh, Ww);
    const float v10 = img_at(goal_img, q.y0 + 1, q.x0, Hh, Ww), v11 = img_at(goal_img, q.y0 + 1, q.x0 + 1, Hh, Ww);
    part += v00 * (1.f - q.wx) * (1.f - q.wy) + v01 * q.wx * (1.f - q.wy) + v10 * (1.f - q.wx) * q.wy +
            v11 * q.wx * q.wy;
  }
  __syncthreads();
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const float gx = goal_coor[m * 2 + 0], gy = goal_coor[m * 2 + 1];
    float best = __int_as_float(0x7f800000);
    int arg = 0;
    for (int n = 0; n < N; ++n) {
      const float dx = gx - pxs[n], dy = gy - pys[n];
      const float d2 = dx * dx + dy * dy;
      if (d2 < best) { best = d2; arg = n; }
    }
    part += sqrtf(best);
    if (argmin_out) argmin_out[s * M + m] = arg;
  }
  const float total = block_sum(part, red);
  if (threadIdx.x == 0) reward[s] = -(normalize ? total / (float)N : total);
}

// Both kernels keep the projected particles (and the backward the arg-min table, M = 5N ints) in dynamic shared
// memory: 28 N bytes for the backward, so the default 48 KB would stop at N = 1755 although the relation search
// accepts N up to ~2690; raise the limit once per device.
constexpr size_t RW_MAX_SMEM = 160 * 1024;
__global__ void k_reward_bwd(const float*, long long, int, const float*, int, int, const float*, int, float, float, float,
                             float, float, float, int, const float*, const int*, float*, long long, int);
static int reward_configure() {
  static DeviceOnce once;
  const int dev = once.pending();
  if (dev < 0) return 0;
  cudaError_t e = cudaFuncSetAttribute(k_reward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RW_MAX_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_reward_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RW_MAX_SMEM);
  if (e != cudaSuccess) return (int)e;
  once.done(dev);
  return 0;
}

int launch_reward(const float* states, long long n_states, long long state_stride, int N, const float* goal_img,
                  int Hh, int Ww, const float* goal_coor, int M, float fx, float fy, float cx, float cy,
                  float off_x, float off_y, int normalize, float* reward, int* argmin_out, cudaStream_t st) {
  if (n_states <= 0) return 0;
  int e = reward_configure();
  if (e) return e;
  if (2 * (size_t)N * sizeof(float) > RW_MAX_SMEM) return (int)cudaErrorInvalidValue;
  k_reward<<<(unsigned)n_states, RW_THREADS, 2 * N * sizeof(float), st>>>(
      states, state_stride, N, goal_img, Hh, Ww, goal_coor, M, fx, fy, cx, cy, off_x, off_y, normalize, reward,
      argmin_out);
  PILE_CHECK_LAUNCH();
  return 0;
}

// d reward / d state for the states whose upstream gradient g_reward[s] is given.
// grid_sample backward zeroes the gradient where the border clamp is active; the min over particles
// passes the gradient to the arg-min particle only.
__global__ void __launch_bounds__(RW_THREADS)
k_reward_bwd(const float* __restrict__ states, long long state_stride, int N, const float* __restrict__ goal_img,
             int Hh, int Ww, const float* __restrict__ goal_coor, int M, float fx, float fy, float cx, float cy,
             float off_x, float off_y, int normalize, const float* __restrict__ g_reward,
             const int* __restrict__ argmin_in, float* __restrict__ g_states, long long g_stride, int accumulate) {
  extern __shared__ float sm[];
  float* pxs = sm;
  float* pys = sm + N;
  int* arg = reinterpret_cast<int*>(sm + 2 * N);   // [M]
  const long long s = blockIdx.x;
  const float* st = states + s * state_stride;
  const float scale = -g_reward[s] * (normalize ? 1.f / (float)N : 1.f);
  if (scale == 0.f) {            // states the loss does not look at (all but the last step, planners.py:438)
    if (!accumulate)
      for (int n = threadIdx.x; n < N * 3; n += blockDim.x) g_states[s * g_stride + n] = 0.f;
    return;
  }

  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float x = st[n * 3 + 0], y = st[n * 3 + 1], z = st[n * 3 + 2];
    pxs[n] = x * fx / z + cx + off_x;
    pys[n] = y * fy / z + cy + off_y;
  }
  for (int m = threadIdx.x; m < M; m += blockDim.x) arg[m] = argmin_in[s * M + m];
  __syncthreads();
  // one warp per particle: the lanes stride over the goal points that picked it (fixed order + fixed shuffle tree:
  // deterministic), lane 0 adds the bilinear image term and writes the three position gradients
  const int lane = threadIdx.x & 31, nwarps = (int)(blockDim.x >> 5);
  for (int n = (int)(threadIdx.x >> 5); n < N; n += nwarps) {
    const float px = pxs[n], py = pys[n];
    float ax = 0.f, ay = 0.f;
    for (int m = lane; m < M; m += 32) {
      if (arg[m] == n) {
        const float dx = px - goal_coor[m * 2 + 0], dy = py - goal_coor[m * 2 + 1];
        const float d = sqrtf(dx * dx + dy * dy);
        if (d > 0.f) { ax += dx / d; ay += dy / d; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ax += __shfl_xor_sync(0xffffffffu, ax, o);
      ay += __shfl_xor_sync(0xffffffffu, ay, o);
    }
    if (lane != 0) continue;
    const Bilinear q = bilinear_setup(px, py, Hh, Ww);
    const float v00 = img_at(goal_img, q.y0, q.x0, Hh, Ww), v01 = img_at(goal_img, q.y0, q.x0 + 1, Hh, Ww);
    const float v10 = img_at(goal_img, q.y0 + 1, q.x0, Hh, Ww), v11 = img_at(goal_img, q.y0 + 1, q.x0 + 1, Hh, Ww);
    float gpx = ((v01 - v00) * (1.f - q.wy) + (v11 - v10) * q.wy) * q.gx_scale + ax;
    float gpy = ((v10 - v00) * (1.f - q.wx) + (v11 - v01) * q.wx) * q.gy_scale + ay;
    gpx *= scale;
    gpy *= scale;
    const float x = st[n * 3 + 0], y = st[n * 3 + 1], z = st[n * 3 + 2];
    const float gx = gpx * fx / z, gy = gpy * fy / z;
    const float gz = -(gpx * x * fx + gpy * y * fy) / (z * z);
    float* g = g_states + s * g_stride + n * 3;
    if (accumulate) { g[0] += gx; g[1] += gy; g[2] += gz; }
    else { g[0] = gx; g[1] = gy; g[2] = gz; }
  }
}

int launch_reward_bwd(const float* states, long long n_states, long long state_stride, int N, const float* goal_img,
                      int Hh, int Ww, const float* goal_coor, int M, float fx, float fy, float cx, float cy,
                      float off_x, float off_y, int normalize, const float* g_reward, const int* argmin_in,
                      float* g_states, long long g_stride, int accumulate, cudaStream_t st) {
  if (n_states <= 0) return 0;
  const size_t smem = 2 * N * sizeof(float) + M * sizeof(int);
  if (smem > RW_MAX_SMEM) return (int)cudaErrorInvalidValue;
  int e = reward_configure();
  if (e) return e;
  k_reward_bwd<<<(unsigned)n_states, RW_THREADS, smem, st>>>(
      states, state_stride, N, goal_img, Hh, Ww, goal_coor, M, fx, fy, cx, cy, off_x, off_y, normalize, g_reward,
      argmin_in, g_states, g_stride, accumulate);
  PILE_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// MPPI weighting (reference planners.py:549-561 computes softmax(w * reward) @ act_seqs).
// Stage 1, one CTA per chunk of samples:  part[c] = (m_c, Z_c, A_c[T*4]) with z = reward_weight * reward,
//   m_c = max z, Z_c = sum exp(z - m_c), A_c = sum exp(z - m_c) * act.
// Stage 2, one CTA: log-sum-exp merge of any number of such partials (chunks of this GPU, or the
//   all-gathered partials of every rank): out = (m, Z, A) with the same meaning; the plan is A / Z.
// ------------------------------------------------------------------------------------------------
constexpr int MPPI_CHUNK = 128;

__global__ void __launch_bounds__(RW_THREADS)
k_mppi_partials(const float* __restrict__ reward, const float* __restrict__ acts, int S, int T, float weight,
                float* __restrict__ part) {
  __shared__ float w[MPPI_CHUNK];
  __shared__ float m_s;
  const int s0 = blockIdx.x * MPPI_CHUNK;
  const int n = min(MPPI_CHUNK, S - s0);
  const int K = 4 * T;
  if (threadIdx.x < MPPI_CHUNK) w[threadIdx.x] = threadIdx.x < n ? weight * reward[s0 + threadIdx.x] : -__int_as_float(0x7f800000);
  __syncthreads();
  if (threadIdx.x < 32) {
    float m = -__int_as_float(0x7f800000);
    for (int i = threadIdx.x; i < MPPI_CHUNK; i += 32) m = fmaxf(m, w[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) m_s = m;
  }
  __syncthreads();
  const float m = m_s;
  if (threadIdx.x < MPPI_CHUNK) w[threadIdx.x] = threadIdx.x < n ? expf(w[threadIdx.x] - m) : 0.f;
  __syncthreads();
  float* out = part + (size_t)blockIdx.x * (2 + K);
  if (threadIdx.x == 0) {
    float z = 0.f;
    for (int i = 0; i < n; ++i) z += w[i];
    out[0] = m;
    out[1] = z;
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float a = 0.f;
    for (int i = 0; i < n; ++i) a = fmaf(w[i], acts[(size_t)(s0 + i) * K + k], a);
    out[2 + k] = a;
  }
}

__global__ void __launch_bounds__(RW_THREADS)
k_mppi_combine(const float* __restrict__ part, int P, int T, float* __restrict__ out) {
  __shared__ float m_s;
  const int K = 4 * T;
  if (threadIdx.x == 0) {
    float m = -__int_as_float(0x7f800000);
    for (int p = 0; p < P; ++p) m = fmaxf(m, part[(size_t)p * (2 + K)]);
    m_s = m;
  }
  __syncthreads();
  const float m = m_s;
  for (int k = threadIdx.x; k < K + 1; k += blockDim.x) {     // k == K handles Z
    float a = 0.f;
    for (int p = 0; p < P; ++p) {
      const float* q = part + (size_t)p * (2 + K);
      const float sc = expf(q[0] - m);
      a = fmaf(sc, k == K ? q[1] : q[2 + k], a);
    }
    if (k == K) out[1] = a; else out[2 + k] = a;
  }
  if (threadIdx.x == 0) out[0] = m;
}

int mppi_num_chunks(int S) { return (S + MPPI_CHUNK - 1) / MPPI_CHUNK; }

int launch_mppi_partials(const float* reward, const float* acts, int S, int T, float weight, float* part,
                         cudaStream_t st) {
  if (S <= 0) return (int)cudaErrorInvalidValue;
  k_mppi_partials<<<mppi_num_chunks(S), RW_THREADS, 0, st>>>(reward, acts, S, T, weight, part);
  PILE_CHECK_LAUNCH();
  return 0;
}

int launch_mppi_combine(const float* part, int P, int T, float* out, cudaStream_t st) {
  k_mppi_combine<<<1, RW_THREADS, 0, st>>>(part, P, T, out);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile

// ------------------------------------------------------------------------------------------------
// Farthest-point sampling (reference utils.fps_np, utils.py:451-466; used by the planner to thin the goal
// pixels, planners.py:620-624, and by the MPC loop to resample observations).  One CTA per point set:
// start at init_idx, repeatedly take the point farthest from the chosen set (first index on ties, like
// numpy argmax); distances are sqrt(sum of squares) in float32 with numpy's left-to-right sum.
// ------------------------------------------------------------------------------------------------
namespace pile {

constexpr int FPS_THREADS = 1024;

__global__ void __launch_bounds__(FPS_THREADS)
k_fps(const float* __restrict__ pts, long long set_stride, int n, int dim, int count, int init_idx,
      const int* __restrict__ init_arr, int squared, float* __restrict__ gap_ws, int* __restrict__ out_idx,
      float* __restrict__ out_pts, float* __restrict__ out_radius) {
  __shared__ float s_val[FPS_THREADS / 32];
  __shared__ int s_arg[FPS_THREADS / 32];
  __shared__ int s_pick;
  const int set = blockIdx.x;
  pts += (size_t)set * set_stride;          // set_stride = 0: every set samples the same cloud
  gap_ws += (size_t)set * n;
  out_idx += (size_t)set * count;
  out_pts += (size_t)set * count * dim;
  int pick = init_arr ? min(max(init_arr[set], 0), n - 1) : init_idx;
  for (int it = 0; it < count; ++it) {
    float c[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < dim; ++k) c[k] = pts[(size_t)pick * dim + k];
    if (threadIdx.x == 0) {
      out_idx[it] = pick;
      for (int k = 0; k < dim; ++k) out_pts[(size_t)it * dim + k] = c[k];
    }
    float best = -1.f;
    int arg = 0x7fffffff;
    // four points per round, all loads first: the gap array is read and written in place every pick, and a
    // load -> store -> load chain through L2 (the compiler keeps that order) costs ~1000 cycles per point
    for (int i0 = threadIdx.x; i0 < n; i0 += 4 * blockDim.x) {
      float p[4][3], old[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        old[u] = 0.f;
        p[u][0] = p[u][1] = p[u][2] = 0.f;
        if (i < n) {
          p[u][0] = pts[(size_t)i * dim];
          if (dim > 1) p[u][1] = pts[(size_t)i * dim + 1];
          if (dim > 2) p[u][2] = pts[(size_t)i * dim + 2];
          if (it > 0) old[u] = gap_ws[i];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < n) {
          float d = __fsub_rn(p[u][0], c[0]);
          float s = __fmul_rn(d, d);
          if (dim > 1) { d = __fsub_rn(p[u][1], c[1]); s = __fadd_rn(s, __fmul_rn(d, d)); }
          if (dim > 2) { d = __fsub_rn(p[u][2], c[2]); s = __fadd_rn(s, __fmul_rn(d, d)); }
          float g = squared ? s : __fsqrt_rn(s);       // dgl's sampler compares squared distances, fps_np norms
          if (it > 0) g = fminf(old[u], g);
          gap_ws[i] = g;
          if (g > best) { best = g; arg = i; }       // ascending i per thread: first index wins inside a thread
        }
      }
    }
    // block arg-max, lowest index on ties
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_arg[threadIdx.x >> 5] = arg; }
    __syncthreads();
    if (threadIdx.x < 32) {
      best = threadIdx.x < (blockDim.x >> 5) ? s_val[threadIdx.x] : -1.f;
      arg = threadIdx.x < (blockDim.x >> 5) ? s_arg[threadIdx.x] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
      }
      if (threadIdx.x == 0) {
        s_pick = arg;
        if (it == count - 1 && out_radius) out_radius[set] = best;
      }
    }
    __syncthreads();
    pick = s_pick;
  }
}

int launch_fps(const float* pts, int n_sets, int n, int dim, int count, int init_idx, float* gap_ws, int* out_idx,
               float* out_pts, float* out_radius, cudaStream_t st) {
  if (n_sets <= 0 || n <= 0 || dim < 1 || dim > 3 || count <= 0 || count > n || init_idx < 0 || init_idx >= n)
    return (int)cudaErrorInvalidValue;
  k_fps<<<n_sets, FPS_THREADS, 0, st>>>(pts, (long long)n * dim, n, dim, count, init_idx, nullptr, 0, gap_ws, out_idx,
                                        out_pts, out_radius);
  PILE_CHECK_LAUNCH();
  return 0;
}

int launch_fps_sets(const float* pts, int shared_cloud, int n_sets, int n, int dim, int count, const int* init_idx,
                    int squared, float* gap_ws, int* out_idx, float* out_pts, float* out_radius, cudaStream_t st) {
  if (n_sets <= 0 || n <= 0 || dim < 1 || dim > 3 || count <= 0 || count > n || !init_idx)
    return (int)cudaErrorInvalidValue;
  k_fps<<<n_sets, FPS_THREADS, 0, st>>>(pts, shared_cloud ? 0 : (long long)n * dim, n, dim, count, 0, init_idx,
                                        squared, gap_ws, out_idx, out_pts, out_radius);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile

// ------------------------------------------------------------------------------------------------
// Adam step on the action sequences + clamp to the workspace box, one launch (reference planners.py:674,
// 742-764: torch.optim.Adam(lr, betas=(0.9, 0.999)) followed by four clamp_ calls).  Same update formula as
// torch's single-tensor Adam; step_size = lr / (1 - b1^t) and bc2_sqrt = sqrt(1 - b2^t) come from the host.
// ------------------------------------------------------------------------------------------------
namespace pile {

struct Box4 { float lo[4], hi[4]; };

__global__ void k_adam_clamp(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float b1, float b2, float step_size,
                             float bc2_sqrt, float eps, Box4 box) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = m[i] + (gi - m[i]) * (1.f - b1);
  const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  float x = p[i] - step_size * (mi / denom);
  const int c = (int)(i & 3);
  x = fminf(fmaxf(x, box.lo[c]), box.hi[c]);
  p[i] = x;
}

int launch_adam_clamp(float* p, const float* g, float* m, float* v, long long n, float b1, float b2, float step_size,
                      float bc2_sqrt, float eps, const float* lo4, const float* hi4, cudaStream_t st) {
  Box4 box;
  for (int i = 0; i < 4; ++i) { box.lo[i] = lo4[i]; box.hi[i] = hi4[i]; }
  k_adam_clamp<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, m, v, n, b1, b2, step_size, bc2_sqrt, eps, box);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
