// Observation -> particles: the planner-side work of one MPC step either side of the rollout
// (reference env/flex_env.py:910-951 obs2ptcl_fixed_num_batch, called with batch_size = 30 at :1028 and :1086).
//
//   depth image --k_depth_count/k_depth_scatter--> foreground points   utils.depth2fgpcd      (utils.py:491-506)
//   points --keys, radix sort, segmented mean--> one point per voxel   utils.downsample_pcd   (utils.py:533-544,
//                                                                       open3d PointCloud::VoxelDownSample)
//   voxel points --k_fps (reward.cu), one start index per set--> N picks utils.fps            (utils.py:423-437,
//                                                                       dgl.geometry.farthest_point_sampler)
//   picks --k_cover_radius--> particle_r ; --k_recenter--> particles    utils.fps :435-437, utils.recenter :468-477
//
// Arithmetic follows the reference's dtypes: the point cloud and every distance are float64 (numpy / open3d),
// the farthest-point sampler runs on float32 copies (torch .float()), the recentred particles are float32.
// Sums that numpy evaluates left to right ((dx*dx + dy*dy) + dz*dz) use explicit _rn operations so that no FMA
// contraction changes them.  Sizes are small (10^4..10^5 points): the kernels are written for determinism and
// zero host round trips, not for a roofline.
#include <cub/cub.cuh>

#include "kernels.h"

namespace pile {

constexpr int OBS_BLOCK = 1024;

__global__ void __launch_bounds__(OBS_BLOCK)
k_depth_count(const float* __restrict__ depth, int HW, float max_depth, int* __restrict__ counts) {
  const int i = blockIdx.x * OBS_BLOCK + threadIdx.x;
  const bool fg = i < HW && depth[i] > 0.f && depth[i] < max_depth;
  const int c = __syncthreads_count(fg);
  if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

// pixel order (row-major) is kept: block offset = sum of the preceding blocks' counts, then a block scan
__global__ void __launch_bounds__(OBS_BLOCK)
k_depth_scatter(const float* __restrict__ depth, int HW, int W, double fx, double fy, double cx, double cy,
                float max_depth, const int* __restrict__ counts, double* __restrict__ pts, int cap,
                int* __restrict__ n_out) {
  __shared__ int warp_sums[OBS_BLOCK / 32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int part = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += OBS_BLOCK) part += counts[b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) warp_sums[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < OBS_BLOCK / 32; ++w) s += warp_sums[w];
    s_base = s;
  }
  __syncthreads();
  const int base = s_base;
  __syncthreads();
  const int i = blockIdx.x * OBS_BLOCK + threadIdx.x;
  float z = 0.f;
  bool fg = false;
  if (i < HW) { z = depth[i]; fg = z > 0.f && z < max_depth; }
  const unsigned ballot = __ballot_sync(0xffffffffu, fg);
  const int in_warp = __popc(ballot & ((1u << lane) - 1u));
  if (lane == 0) warp_sums[warp] = __popc(ballot);
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += warp_sums[w];
  const int pos = base + before + in_warp;
  if (fg && pos < cap) {
    const int v = i / W, u = i - v * W;
    const double zd = (double)z;
    pts[3 * (size_t)pos + 0] = __ddiv_rn(__dmul_rn((double)u - cx, zd), fx);      // (pos_x - cx) * depth / fx
    pts[3 * (size_t)pos + 1] = __ddiv_rn(__dmul_rn((double)v - cy, zd), fy);
    pts[3 * (size_t)pos + 2] = zd;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == OBS_BLOCK - 1) *n_out = pos + (fg ? 1 : 0);
}

// ---- voxel-grid downsample ---------------------------------------------------------------------------
__global__ void __launch_bounds__(OBS_BLOCK)
k_bbox_min(const double* __restrict__ pts, int n, double* __restrict__ out3) {
  __shared__ double s[3][OBS_BLOCK / 32];
  double m[3] = {1e300, 1e300, 1e300};
  for (int i = threadIdx.x; i < n; i += OBS_BLOCK)
    for (int a = 0; a < 3; ++a) m[a] = fmin(m[a], pts[3 * (size_t)i + a]);
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m[a] = fmin(m[a], __shfl_xor_sync(0xffffffffu, m[a], o));
    if ((threadIdx.x & 31) == 0) s[a][threadIdx.x >> 5] = m[a];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = s[threadIdx.x][0];
    for (int w = 1; w < OBS_BLOCK / 32; ++w) v = fmin(v, s[threadIdx.x][w]);
    out3[threadIdx.x] = v;
  }
}

// open3d: voxel_min_bound = min_bound - voxel_size * 0.5; index = floor((p - voxel_min_bound) / voxel_size)
__global__ void k_voxel_keys(const double* __restrict__ pts, int n, const double* __restrict__ minb, double voxel,
                             unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long key = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double lo = __dsub_rn(minb[a], __dmul_rn(voxel, 0.5));
    const long long q = (long long)floor(__ddiv_rn(__dsub_rn(pts[3 * (size_t)i + a], lo), voxel));
    key = (key << 21) | (unsigned long long)(q & 0x1fffff);
  }
  keys[i] = key;
  vals[i] = i;
}

__global__ void k_seg_flags(const unsigned long long* __restrict__ keys, int n, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per voxel head: mean of its points in ascending input index (the stable sort keeps that order; it
// is the order open3d accumulates in)
__global__ void k_seg_mean(const double* __restrict__ pts, const unsigned long long* __restrict__ keys,
                           const int* __restrict__ vals, const int* __restrict__ flags,
                           const int* __restrict__ seg_id, int n, double* __restrict__ out, int* __restrict__ m_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == n - 1) *m_out = seg_id[i];               // inclusive scan of the head flags
  if (!flags[i]) return;
  double sx = 0.0, sy = 0.0, sz = 0.0;
  int cnt = 0;
  const unsigned long long key = keys[i];
  for (int j = i; j < n && keys[j] == key; ++j) {
    const double* p = pts + 3 * (size_t)vals[j];
    sx = __dadd_rn(sx, p[0]); sy = __dadd_rn(sy, p[1]); sz = __dadd_rn(sz, p[2]);
    ++cnt;
  }
  double* o = out + 3 * (size_t)(seg_id[i] - 1);
  o[0] = __ddiv_rn(sx, (double)cnt); o[1] = __ddiv_rn(sy, (double)cnt); o[2] = __ddiv_rn(sz, (double)cnt);
}

struct VoxelWs {
  double* minb;
  unsigned long long *keys_in, *keys_out;
  int *vals_in, *vals_out, *flags, *seg;
  void* cub_tmp;
  size_t cub_bytes;
};

static size_t voxel_cub_bytes(int n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, n, 0, 63);
  cub::DeviceScan::InclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, n);
  return a > b ? a : b;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t voxel_downsample_bytes(int n) {
  if (n < 1) n = 1;
  return 256 + 2 * align256(sizeof(unsigned long long) * (size_t)n) + 4 * align256(sizeof(int) * (size_t)n) +
         align256(voxel_cub_bytes(n));
}

int launch_voxel_downsample(const double* pts, int n, double voxel, double* out_pts, int* m_out, void* ws,
                            cudaStream_t st) {
  if (n <= 0 || !(voxel > 0.0)) return (int)cudaErrorInvalidValue;
  char* p = static_cast<char*>(ws);
  VoxelWs w;
  w.minb = reinterpret_cast<double*>(p); p += 256;
  w.keys_in = reinterpret_cast<unsigned long long*>(p); p += align256(sizeof(unsigned long long) * (size_t)n);
  w.keys_out = reinterpret_cast<unsigned long long*>(p); p += align256(sizeof(unsigned long long) * (size_t)n);
  w.vals_in = reinterpret_cast<int*>(p); p += align256(sizeof(int) * (size_t)n);
  w.vals_out = reinterpret_cast<int*>(p); p += align256(sizeof(int) * (size_t)n);
  w.flags = reinterpret_cast<int*>(p); p += align256(sizeof(int) * (size_t)n);
  w.seg = reinterpret_cast<int*>(p); p += align256(sizeof(int) * (size_t)n);
  w.cub_tmp = p;
  w.cub_bytes = voxel_cub_bytes(n);
  const int blocks = (n + 255) / 256;
  k_bbox_min<<<1, OBS_BLOCK, 0, st>>>(pts, n, w.minb);
  PILE_CHECK_LAUNCH();
  k_voxel_keys<<<blocks, 256, 0, st>>>(pts, n, w.minb, voxel, w.keys_in, w.vals_in);
  PILE_CHECK_LAUNCH();
  cudaError_t e = cub::DeviceRadixSort::SortPairs(w.cub_tmp, w.cub_bytes, w.keys_in, w.keys_out, w.vals_in, w.vals_out,
                                                  n, 0, 63, st);
  if (e != cudaSuccess) return (int)e;
  k_seg_flags<<<blocks, 256, 0, st>>>(w.keys_out, n, w.flags);
  PILE_CHECK_LAUNCH();
  e = cub::DeviceScan::InclusiveSum(w.cub_tmp, w.cub_bytes, w.flags, w.seg, n, st);
  if (e != cudaSuccess) return (int)e;
  k_seg_mean<<<blocks, 256, 0, st>>>(pts, w.keys_out, w.vals_out, w.flags, w.seg, n, out_pts, m_out);
  PILE_CHECK_LAUNCH();
  return 0;
}

int launch_depth_to_points(const float* depth, int H, int W, const double* cam4, float max_depth, double* out_pts,
                           int cap, int* n_out, int* ws_counts, cudaStream_t st) {
  const long long HW = (long long)H * W;
  if (H <= 0 || W <= 0 || HW > (1ll << 30) || cap <= 0) return (int)cudaErrorInvalidValue;
  const int blocks = (int)((HW + OBS_BLOCK - 1) / OBS_BLOCK);
  k_depth_count<<<blocks, OBS_BLOCK, 0, st>>>(depth, (int)HW, max_depth, ws_counts);
  PILE_CHECK_LAUNCH();
  k_depth_scatter<<<blocks, OBS_BLOCK, 0, st>>>(depth, (int)HW, W, cam4[0], cam4[1], cam4[2],
                                                 cam4[3], max_depth, ws_counts, out_pts, cap, n_out);
  PILE_CHECK_LAUNCH();
  return 0;
}

// ---- covering radius and recentring ---------------------------------------------------------------------
__device__ __forceinline__ double dist_rn(const double* p, double sx, double sy, double sz) {
  const double dx = __dsub_rn(p[0], sx), dy = __dsub_rn(p[1], sy), dz = __dsub_rn(p[2], sz);
  return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));   // np.linalg.norm
}

// particle_r[s] = max over cloud points of the distance to the nearest pick (utils.fps :435-437)
__global__ void __launch_bounds__(OBS_BLOCK)
k_cover_radius(const double* __restrict__ pcd, int m, const float* __restrict__ picks, int N,
               double* __restrict__ radius) {
  extern __shared__ double sp[];          // picks of this set as doubles [N][3]
  __shared__ double wmax[OBS_BLOCK / 32];
  const int s = blockIdx.x;
  for (int i = threadIdx.x; i < 3 * N; i += OBS_BLOCK) sp[i] = (double)picks[(size_t)s * N * 3 + i];
  __syncthreads();
  double worst = 0.0;
  for (int i = threadIdx.x; i < m; i += OBS_BLOCK) {
    double best = 1e300;
    for (int k = 0; k < N; ++k) best = fmin(best, dist_rn(pcd + 3 * (size_t)i, sp[3 * k], sp[3 * k + 1], sp[3 * k + 2]));
    worst = fmax(worst, best);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = worst;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < OBS_BLOCK / 32; ++w) worst = fmax(worst, wmax[w]);
    radius[s] = worst;
  }
}

// out[s][k] = mean of the cloud points closer than r_s = min(r_cap, r_scale * radius[s]) to pick k (utils.recenter)
__global__ void __launch_bounds__(128)
k_recenter(const double* __restrict__ pcd, int m, const float* __restrict__ picks, int N,
           const double* __restrict__ radius, double r_cap, double r_scale, float* __restrict__ out) {
  __shared__ double ssum[4][3];
  __shared__ int scnt[4];
  const int k = blockIdx.x, s = blockIdx.y;
  const float* c = picks + ((size_t)s * N + k) * 3;
  const double cx = (double)c[0], cy = (double)c[1], cz = (double)c[2];
  const double r = fmin(r_cap, __dmul_rn(r_scale, radius[s]));
  double sx = 0.0, sy = 0.0, sz = 0.0;
  int cnt = 0;
  for (int i = threadIdx.x; i < m; i += 128) {
    const double* p = pcd + 3 * (size_t)i;
    if (dist_rn(p, cx, cy, cz) < r) { sx += p[0]; sy += p[1]; sz += p[2]; ++cnt; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5][0] = sx; ssum[threadIdx.x >> 5][1] = sy; ssum[threadIdx.x >> 5][2] = sz; scnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 4; ++w) { sx += ssum[w][0]; sy += ssum[w][1]; sz += ssum[w][2]; cnt += scnt[w]; }
    float* o = out + ((size_t)s * N + k) * 3;
    // an empty neighbourhood is NaN, as numpy's mean of an empty selection
    o[0] = (float)(sx / (double)cnt); o[1] = (float)(sy / (double)cnt); o[2] = (float)(sz / (double)cnt);
  }
}

int launch_cover_radius(const double* pcd, int m, const float* picks, int S, int N, double* radius, cudaStream_t st) {
  if (m <= 0 || S <= 0 || N <= 0 || (size_t)N * 24 > 96 * 1024) return (int)cudaErrorInvalidValue;
  static DeviceOnce once;
  const int once_dev = once.pending();
  if (once_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(k_cover_radius, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return (int)e;
    once.done(once_dev);
  }
  k_cover_radius<<<S, OBS_BLOCK, (size_t)N * 24, st>>>(pcd, m, picks, N, radius);
  PILE_CHECK_LAUNCH();
  return 0;
}

int launch_recenter(const double* pcd, int m, const float* picks, int S, int N, const double* radius, double r_cap,
                    double r_scale, float* out, cudaStream_t st) {
  if (m <= 0 || S <= 0 || N <= 0) return (int)cudaErrorInvalidValue;
  k_recenter<<<dim3(N, S), 128, 0, st>>>(pcd, m, picks, N, radius, r_cap, r_scale, out);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
