// Device-side bookkeeping of the gradient-descent planner loop (reference planners.py:682-764), so that one
// planner iteration -- rollout, reward, best tracking, backward, Adam + clamp -- is a fixed sequence of launches
// with no host-side value in it and can be captured once as a CUDA graph and replayed n_iter times:
//   k_gd_track         per state variant the best trajectory so far (:721-727) and rew_mean / rew_std (:737-738)
//   k_adam_clamp_dev   torch.optim.Adam's update with the step number read from device memory, + clamp (:756-764)
//   k_counter_add      advances the iteration counter on the stream
#include "common.cuh"
#include "kernels.h"

namespace pile {

constexpr int TRK_THREADS = 128;

__global__ void __launch_bounds__(TRK_THREADS)
k_gd_track(const float* __restrict__ reward, const float* __restrict__ acts, int n_sample, int n_batch, int T,
           float* __restrict__ max_reward, int* __restrict__ max_idx, float* __restrict__ best_actions,
           float* __restrict__ rew_mean, float* __restrict__ rew_std, const int* __restrict__ iter_dev, int stat_every,
           int stat_stride) {
  __shared__ float s_best[TRK_THREADS / 32];
  __shared__ int s_arg[TRK_THREADS / 32];
  __shared__ double s_sum[TRK_THREADS / 32], s_sq[TRK_THREADS / 32];
  __shared__ int s_take;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float best = __int_as_float(0xff800000);   // -inf
  int arg = 0x7fffffff;
  double sum = 0.0;
  for (int s = threadIdx.x; s < n_sample; s += blockDim.x) {
    const float r = reward[(long long)s * n_batch + b];
    if (r > best || (r == best && s < arg)) { best = r; arg = s; }
    sum += (double)r;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  if (lane == 0) { s_best[warp] = best; s_arg[warp] = arg; s_sum[warp] = sum; }
  __syncthreads();
  double mean = 0.0;
  if (threadIdx.x == 0) {
    for (int w = 1; w < TRK_THREADS / 32; ++w) {
      if (s_best[w] > best || (s_best[w] == best && s_arg[w] < arg)) { best = s_best[w]; arg = s_arg[w]; }
      sum += s_sum[w];
    }
    const bool take = arg != 0x7fffffff && best > max_reward[b];
    s_take = take ? arg : -1;
    if (take) { max_reward[b] = best; max_idx[b] = arg; }
    s_sum[0] = sum / (double)n_sample;
  }
  __syncthreads();
  const int take = s_take;
  if (take >= 0) {
    const float* src = acts + ((long long)take * n_batch + b) * T * 4;
    for (int k = threadIdx.x; k < T * 4; k += blockDim.x) best_actions[(long long)b * T * 4 + k] = src[k];
  }
  if (b % stat_every != 0) return;
  // statistics of a scene's state variant 0 over the samples (reward_seqs[:, 0].mean() / .std(), unbiased like torch)
  mean = s_sum[0];
  double sq = 0.0;
  for (int s = threadIdx.x; s < n_sample; s += blockDim.x) {
    const double d = (double)reward[(long long)s * n_batch + b] - mean;
    sq += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) s_sq[warp] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < TRK_THREADS / 32; ++w) sq += s_sq[w];
    const long long it = (long long)(b / stat_every) * stat_stride + *iter_dev;
    rew_mean[it] = (float)mean;
    rew_std[it] = n_sample > 1 ? (float)sqrt(sq / (double)(n_sample - 1)) : __int_as_float(0x7fc00000);
  }
}

int launch_gd_track(const float* reward, const float* acts, int n_sample, int n_batch, int T, float* max_reward,
                    int* max_idx, float* best_actions, float* rew_mean, float* rew_std, const int* iter_dev,
                    int stat_every, int stat_stride, cudaStream_t st) {
  k_gd_track<<<n_batch, TRK_THREADS, 0, st>>>(reward, acts, n_sample, n_batch, T, max_reward, max_idx, best_actions,
                                              rew_mean, rew_std, iter_dev, stat_every, stat_stride);
  PILE_CHECK_LAUNCH();
  return 0;
}

struct Box4d { float lo[4], hi[4]; };

__global__ void k_adam_clamp_dev(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, const int* __restrict__ iter_dev, float lr,
                                 float b1, float b2, float eps, Box4d box) {
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    // torch.optim.Adam: bias corrections in double on the host; same arithmetic here, once per block
    const double step = (double)(*iter_dev + 1);
    const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
    s_step_size = (float)((double)lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = m[i] + (gi - m[i]) * (1.f - b1);
  const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / s_bc2_sqrt + eps;
  float x = p[i] - s_step_size * (mi / denom);
  const int c = (int)(i & 3);
  x = fminf(fmaxf(x, box.lo[c]), box.hi[c]);
  p[i] = x;
}

int launch_adam_clamp_dev(float* p, const float* g, float* m, float* v, long long n, const int* iter_dev, float lr,
                          float b1, float b2, float eps, const float* lo4, const float* hi4, cudaStream_t st) {
  Box4d box;
  for (int i = 0; i < 4; ++i) { box.lo[i] = lo4[i]; box.hi[i] = hi4[i]; }
  k_adam_clamp_dev<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, m, v, n, iter_dev, lr, b1, b2, eps, box);
  PILE_CHECK_LAUNCH();
  return 0;
}

__global__ void k_counter_add(int* c, int delta) { *c += delta; }

int launch_counter_add(int* counter, int delta, cudaStream_t st) {
  k_counter_add<<<1, 1, 0, st>>>(counter, delta);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
