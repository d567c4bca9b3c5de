// Shared definitions for the pile-GNN rollout kernels (sm_100a).
//
// Data layout in HBM (all float32 / int32, row-major, one "row" = one particle or one relation):
//   states      [Bt, N, 3]                     particle positions (camera frame)
//   CSR         rowptr [Bt, N+1] (local offsets), col/row [Bt, KMAX*N] (sender / receiver index,
//               relations of a sample sorted by (receiver, sender) = torch.nonzero() order,
//               reference model/gnn_dyn.py:247)
//   node feats  [Bt*N, H]      relation feats [Bt*KMAX*N, H]  (slot = b*KMAX*N + local edge id)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

namespace pile {

constexpr int H = 64;          // nf_effect compiled in (config/mpc/config.yaml:91); checked at run time
constexpr int KMAX = 10;       // max relations per receiver (model/gnn_dyn.py:231)
constexpr int PSTEP = 3;       // propagation steps (model/gnn_dyn.py:160)
constexpr int TILE = 128;      // rows per GEMM tile
constexpr int LDA = H + 4;     // shared-memory activation row stride (conflict-free float4 rows)
constexpr int LDX = 12;        // shared-memory stride of the 8-wide input feature tile
constexpr int NT = 256;        // threads per GEMM CTA: 8 warps x (4 rows/lane x 8 cols/warp)
constexpr int NSM = 148;
constexpr int TC_EDGE2_FLOATS = 13312;  // 52 KB, layout in edge_tmem.cu
constexpr int TC_BWD_EDGE_FLOATS = 13312;  // 52 KB, layout in bwd_tc.cu
constexpr int TC_BWD_NODE_FLOATS = 26624;  // 104 KB, layout in bwd_node_tc.cu
constexpr int TC_NODE_FLOATS = 29952;   // 117 KB of bf16 hi/lo weight images, layout in node_tc.cu

// ---- packed weight buffer (floats). Forward blocks are transposed [K][H]; backward blocks keep the
// checkpoint's [out][in] layout (that IS the [K=out][cols=in] operand of the dgrad GEMM).
enum WSlot {
  W_PE0T = 0,   // [8][H]  rows: s_delta xyz, attr, dens, 0,0,0
  B_PE0,        // [H]
  W_PE1T,       // [H][H]
  B_PE1,        // [H]
  W_RE0T,       // [8][H]  rows: attr_r, attr_s, dx,dy,dz, dens, 0,0
  B_RE0,
  W_RE1T, B_RE1, W_RE2T, B_RE2,
  W_ET, W_RT, W_ST,   // relation propagator split by input block [rel_enc | eff_r | eff_s]
  WD_RP, B_RP,        // its density column and bias
  W_PT, W_AT,         // particle propagator split [p_enc | agg]
  WD_PP, B_PP,
  W_V0T, B_V0,        // predictor
  W_V1T,              // [H][4] (3 outputs + pad)
  B_V1,               // [4]
  // backward ([out][in])
  W_PE0,              // [H][8]
  W_PE1, W_RE0 /*[H][8]*/, W_RE1, W_RE2, W_E, W_R, W_S, W_P, W_A, W_V0,
  W_V1,               // [4][H] (row 3 zero)
  // tensor-core operands of the relation encoder (edge_tc.cu): bf16 [hi | lo] images in the canonical
  // K-major UMMA layout (tc.cuh) of W0aug [64 x 16], W1aug, W2aug, WEaug [64 x 80]; 64 KB, one bulk copy
  TC_EDGE,
  // tensor-core operands of the particle kernels (node_tc.cu): PE0aug, PE1aug, WPaug, [W_r;W_s], W_a, V0aug, V1aug
  TC_NODE,
  // relation-encoder operands of the A-in-TMEM variant (edge_tmem.cu): W0 [64 x 16], RE1, RE2, W_e [64 x 64]
  TC_EDGE2,
  // relation-encoder backward on tcgen05 (bwd_tc.cu): W_e^T, RE2^T, RE1^T [64 x 64], RE0[:, 2:5]^T padded to [16 x 64]
  TC_BWD_EDGE,
  // particle-side backward on tcgen05 (bwd_node_tc.cu): W_a^T, W_r^T, W_s^T, W_p^T, PE1^T [64 x 64], PE0[:, 0:3]^T
  // padded to [16 x 64], V0^T [64 x 64], V1^T padded to [64 x 16]
  TC_BWD_NODE,
  W_NUM
};

__host__ __device__ inline int wslot_size(int s) {
  switch (s) {
    case W_PE0T: case W_RE0T: case W_PE0: case W_RE0: return 8 * H;
    case B_PE0: case B_PE1: case B_RE0: case B_RE1: case B_RE2: case WD_RP: case B_RP:
    case WD_PP: case B_PP: case B_V0: return H;
    case W_V1T: case W_V1: return 4 * H;
    case B_V1: return 4;
    case TC_EDGE: return 4 * H * H;
    case TC_NODE: return TC_NODE_FLOATS;
    case TC_EDGE2: return TC_EDGE2_FLOATS;
    case TC_BWD_EDGE: return TC_BWD_EDGE_FLOATS;
    case TC_BWD_NODE: return TC_BWD_NODE_FLOATS;
    default: return H * H;
  }
}
__host__ __device__ inline int wslot_offset(int s) {
  int o = 0;
  for (int i = 0; i < s; ++i) o += wslot_size(i);
  return o;
}

// pusher frame handed to the s_delta kernels: simulator (planners.py:192-257) or real robot (planners.py:259-300)
struct PushCam {
  float m[12];        // kind 0: rows 0..2 of the 4x4 world->camera(OpenCV) matrix
  float global_scale; // kind 0
  float pusher_w;     // half width of the pusher: 0.8/24 (sim, planners.py:225) or 0.048 (real, planners.py:279)
  float decay;        // 0.01
  int kind;           // 0 = simulator frame, 1 = real-robot frame
  float s2r_scale;    // kind 1: action -> metres
  float shift_x, shift_y;   // kind 1: workspace centre subtracted from the particles (0 for kind 0)
  float height;       // kind 1: pusher height 0.88
};

// One-time kernel attribute setup (cudaFuncSetAttribute) is per DEVICE: every translation unit remembers which
// devices it has configured.  Racing threads may configure a device twice, which is harmless.
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  // current device index if it still has to be configured, -1 otherwise
  int pending() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) d = 0;
    if (d >= 0 && d < 64 && ((mask.load(std::memory_order_acquire) >> d) & 1ull)) return -1;
    return d;
  }
  void done(int d) {
    if (d >= 0 && d < 64) mask.fetch_or(1ull << d, std::memory_order_release);
  }
};

#define PILE_CHECK_LAUNCH() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// acc[i][j] += sum_k A[(lane+32i)*lda + k] * Wt[k*H + c0 + j],   k < K (K % 4 == 0)
// A: shared, row-major activations; Wt: shared, [K][H].  One warp covers 128 rows x 8 columns.
template <int K>
__device__ __forceinline__ void gemm_rows4x8(const float* __restrict__ A, int lda,
                                             const float* __restrict__ Wt, int lane, int c0,
                                             float (&acc)[4][8]) {
  const float* a0 = A + lane * lda;
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = ld4(a0 + i * 32 * lda + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w0 = ld4(Wt + (k + kk) * H + c0);
      const float4 w1 = ld4(Wt + (k + kk) * H + c0 + 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
        acc[i][0] = fmaf(av, w0.x, acc[i][0]);
        acc[i][1] = fmaf(av, w0.y, acc[i][1]);
        acc[i][2] = fmaf(av, w0.z, acc[i][2]);
        acc[i][3] = fmaf(av, w0.w, acc[i][3]);
        acc[i][4] = fmaf(av, w1.x, acc[i][4]);
        acc[i][5] = fmaf(av, w1.y, acc[i][5]);
        acc[i][6] = fmaf(av, w1.z, acc[i][6]);
        acc[i][7] = fmaf(av, w1.w, acc[i][7]);
      }
    }
  }
}

// cooperative copy of n floats (n % 4 == 0, 16B aligned) global -> shared
__device__ __forceinline__ void load_block(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) st4(dst + i, ld4(src + i));
}

__device__ __forceinline__ void acc_set_bias(float (&acc)[4][8], const float* bias, int c0) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = bias[c0 + j];
}

// ReLU in place, returning the >0 bit mask of each of the 4 rows (bit j = column c0+j)
__device__ __forceinline__ void acc_relu(float (&acc)[4][8], unsigned (&bits)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool pos = acc[i][j] > 0.f;
      m |= pos ? (1u << j) : 0u;
      acc[i][j] = pos ? acc[i][j] : 0.f;
    }
    bits[i] = m;
  }
}

__device__ __forceinline__ void acc_apply_mask(float (&acc)[4][8], const unsigned (&bits)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = (bits[i] >> j) & 1u ? acc[i][j] : 0.f;
}

// registers -> shared activations (row-major, stride lda)
__device__ __forceinline__ void acc_to_smem(float* A, int lda, int lane, int c0, const float (&acc)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* p = A + (lane + 32 * i) * lda + c0;
    st4(p, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    st4(p + 4, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
  }
}

// registers -> global rows [row0 + lane + 32 i][c0..c0+7]  (each thread writes full 32 B sectors)
__device__ __forceinline__ void acc_to_global(float* __restrict__ G, long long row0, int nrows, int lane,
                                              int c0, const float (&acc)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = lane + 32 * i;
    if (r < nrows) {
      float* p = G + (row0 + r) * H + c0;
      st4(p, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
      st4(p + 4, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
    }
  }
}

__device__ __forceinline__ void mask_to_global(uint8_t* __restrict__ M, long long row0, int nrows, int lane,
                                               int warp, const unsigned (&bits)[4]) {
  if (M == nullptr) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = lane + 32 * i;
    if (r < nrows) M[(row0 + r) * 8 + warp] = (uint8_t)bits[i];
  }
}

__device__ __forceinline__ void mask_from_global(const uint8_t* __restrict__ M, long long row0, int nrows,
                                                 int lane, int warp, unsigned (&bits)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = lane + 32 * i;
    bits[i] = r < nrows ? M[(row0 + r) * 8 + warp] : 0u;
  }
}

}  // namespace pile
