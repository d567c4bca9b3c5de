// General-width engine, inference: every Hp-wide layer of the hoisted model step on the tensor cores -- relation side RE1,
// RE2 and C_e = W_e r3 + w_d d + b, particle side PE1, (P_r, P_s) = (W_r, W_s) eff, the particle propagator and V0 (reference
// model/gnn_dyn.py:62-111, 174-198 at any nf_effect, model/gnn_dyn.py:119).
//
//   y[rows, Hp] = act(sum_s x_s[rows, Hp] W_s^T + bias + d w_d + res)     Hp = 64 NB, NB = 1..4 (nf_effect <= 256 zero-padded)
//
// Same arithmetic as the width-64 planner engines (tc.cuh): every fp32 operand is split into bf16 hi + lo, a product is
// three tcgen05.mma passes accumulated in fp32 in tensor memory.  One 256-thread CTA per 128-row tile, two CTAs per SM
// (while one waits for its MMAs or streams rows the other one works).  Per tile and 64-wide K block: the rows' block is
// loaded with whole 128-byte lines, split and written as the canonical K-major A tile (32 KB); the K block's weight
// rows for ALL Hp outputs ([Hp x 64] hi | lo, <= 64 KB) arrive by one TMA bulk copy; 12 MMAs of M = 128, N = Hp, K = 16
// accumulate into Hp TMEM columns.  After the last K block the accumulator leaves through the (dead) A tile as a
// row-swizzled staging buffer so that the global stores are whole lines.
//
// The bf16 weight images are built on the device from the engine's fp32 [in][out] images (k_g_tc_image) into free tape
// space at the start of every inference step: no change to the packed-weights layout or to the ABI.
#include "kernels.h"
#include "tc_tile.cuh"

namespace pile {
namespace general {

constexpr int GT_THREADS = 256;

struct GtHdr {
  uint64_t mma_bar, w_bar;
  uint32_t tmem_base;
};

__host__ __device__ constexpr uint32_t gt_wpart(int Hp) { return (uint32_t)Hp * 128u; }          // [Hp x 64] bf16
__host__ __device__ constexpr uint32_t gt_smem(int Hp) { return 2 * A_BYTES + 2 * gt_wpart(Hp) + 128; }

// Wt: fp32 [Hp in][Hp out] (the forward image of general.cu) -> per K block kb: hi then lo canonical image of the B operand
// [N = Hp outputs x K = 64 inputs]: bf16 index (k / 8) * (8 Hp) + (n / 8) * 64 + (n % 8) * 8 + (k % 8)
struct TcImages {
  static constexpr int MAXN = 12;
  const float* src[MAXN];
};
__global__ void k_g_tc_image(const TcImages im, int Hp, __nv_bfloat16* __restrict__ img0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Hp * Hp) return;
  const float* __restrict__ Wt = im.src[blockIdx.y];
  __nv_bfloat16* __restrict__ img = img0 + (size_t)blockIdx.y * 2 * Hp * Hp;
  const int kin = idx / Hp, n = idx - kin * Hp;
  const float w = Wt[idx];
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  const int kb = kin >> 6, k = kin & 63;
  const size_t off = (size_t)kb * (2 * Hp * 64) + (size_t)(k >> 3) * (8 * Hp) + (n >> 3) * 64 + (n & 7) * 8 + (k & 7);
  img[off] = hi;
  img[off + (size_t)Hp * 64] = lo;
}

// rows [row0, row0 + nrows) x columns [col0, col0 + 64) of a row-major [*, ld] array -> hi/lo A tile (rows past nrows are
// zero), in two halves so that the global loads of the NEXT block are in flight while the tensor core works on this one:
// block_issue (32 registers per thread; a warp instruction covers 8 rows x 128 bytes, whole lines) and block_commit (split +
// conflict-free 16-byte shared-memory stores).
__device__ __forceinline__ void block_issue(const float* __restrict__ x, long long row0, int nrows, int ld, int col0, int t,
                                            float (&o)[4][8]) {
  const int w = t >> 5, lane = t & 31;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = w * 16 + (j >> 1) * 8 + (lane & 7), kc = (j & 1) * 4 + (lane >> 3);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[j][i] = 0.f;
    if (row < nrows) ld8(x + (row0 + row) * ld + col0 + kc * 8, o[j]);
  }
}
__device__ __forceinline__ void block_commit(const float (&o)[4][8], int t, uint8_t* a_hi, uint8_t* a_lo) {
  const int w = t >> 5, lane = t & 31;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = w * 16 + (j >> 1) * 8 + (lane & 7), kc = (j & 1) * 4 + (lane >> 3);
    store_chunk(a_hi, a_lo, (row >> 3) * A_SBO + (row & 7) * 16 + kc * A_LBO, o[j]);
  }
}

// the same A tile block, computed instead of loaded: columns [col0, col0 + 64) of ReLU(x8 W8 + b8), x8 [*, 8] rows
__device__ __forceinline__ void input_block_to_tile(const float* __restrict__ x8, const float* __restrict__ w8,
                                                    const float* __restrict__ b8, long long row0, int nrows, int ld, int col0,
                                                    int t, uint8_t* a_hi, uint8_t* a_lo) {
  const int w = t >> 5, lane = t & 31;
#pragma unroll
  for (int h = 0; h < 2; ++h) {          // this thread's two rows
    const int row = w * 16 + h * 8 + (lane & 7);
    float in[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) in[i] = 0.f;
    const bool ok = row < nrows;
    if (ok) {
      const float4 u = ld4(x8 + (row0 + row) * 8), v = ld4(x8 + (row0 + row) * 8 + 4);
      in[0] = u.x; in[1] = u.y; in[2] = u.z; in[3] = u.w; in[4] = v.x; in[5] = v.y; in[6] = v.z; in[7] = v.w;
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {        // and two of the eight K chunks per row
      const int kc = c * 4 + (lane >> 3);
      const float* wc = w8 + col0 + kc * 8;
      float o[8];
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b8 + col0 + kc * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(b8 + col0 + kc * 8 + 4));
      o[0] = b0.x; o[1] = b0.y; o[2] = b0.z; o[3] = b0.w; o[4] = b1.x; o[5] = b1.y; o[6] = b1.z; o[7] = b1.w;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wc + (long long)j * ld));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wc + (long long)j * ld + 4));
        o[0] = fmaf(in[j], w0.x, o[0]); o[1] = fmaf(in[j], w0.y, o[1]); o[2] = fmaf(in[j], w0.z, o[2]); o[3] = fmaf(in[j], w0.w, o[3]);
        o[4] = fmaf(in[j], w1.x, o[4]); o[5] = fmaf(in[j], w1.y, o[5]); o[6] = fmaf(in[j], w1.z, o[6]); o[7] = fmaf(in[j], w1.w, o[7]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = ok ? fmaxf(o[i], 0.f) : 0.f;
      store_chunk(a_hi, a_lo, (row >> 3) * A_SBO + (row & 7) * 16 + kc * A_LBO, o);
    }
  }
}

struct LinTcArgs {
  int nsrc;                 // 1 or 2 sources
  const float* x[2];        // [rows, Hp]
  const uint8_t* img[2];    // weight images (k_g_tc_image)
  const float* bias;        // [Hp] or nullptr
  const float* wd;          // [Hp] density column or nullptr
  const float* dens;        // [B]
  const float* res;         // residual rows [rows, Hp] added before the activation, or nullptr
  float* y;
  const int* rowptr;        // EDGE: relation rows, tiled per sample
  int relu, B, N;
  // optional: source 0 is not read but computed on the fly, x_0 = ReLU(x8 W8 + b8) -- the narrow input layer in front of
  // this one (RE0 in front of RE1): its [rows, Hp] output never goes through HBM
  const float* x8;          // [rows, 8] or nullptr
  const float* w8;          // [8][Hp]
  const float* b8;          // [Hp]
};

template <int NB, bool EDGE>
__global__ void __launch_bounds__(GT_THREADS, 2) k_g_lin_tc(const LinTcArgs a) {
  constexpr int Hp = 64 * NB;
  constexpr uint32_t WPART = gt_wpart(Hp), WROW = 2 * WPART;
  constexpr uint32_t TCOLS = NB == 1 ? 64u : (NB == 2 ? 128u : 256u);
  constexpr uint32_t BL = b_lbo(Hp);
  constexpr uint32_t idesc = tc::make_idesc_bf16(TILE, Hp);
  extern __shared__ __align__(128) unsigned char sm[];
  uint8_t* const a_hi = sm;
  uint8_t* const a_lo = sm + A_BYTES;
  uint8_t* const w_sm = sm + 2 * A_BYTES;
  GtHdr* const hdr = reinterpret_cast<GtHdr*>(w_sm + WROW);
  const int t = threadIdx.x, wig = t >> 5, lane = t & 31;
  const int r = (wig & 3) * 32 + lane, half = wig >> 2;
  const int B = a.B, N = a.N;

  if (t < 32) tc::tmem_alloc(&hdr->tmem_base, TCOLS);
  if (t == 0) {
    tc::mbar_init(&hdr->mma_bar, 1);
    tc::mbar_init(&hdr->w_bar, 1);
    tc::mbar_init_fence();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = hdr->tmem_base;
  const uint32_t taddr = tmem_d + ((uint32_t)((wig & 3) * 32) << 16);
  const uint32_t sa_hi = tc::smem_u32(a_hi), sa_lo = tc::smem_u32(a_lo);
  const uint32_t sw_hi = tc::smem_u32(w_sm), sw_lo = sw_hi + WPART;
  uint32_t mph = 0, wph = 0;

  // relation rows are tiled per sample (the slots of a sample are contiguous, its last tile is partial); particle rows
  // are one contiguous range
  const int tps = (KMAX * N + TILE - 1) / TILE;
  const long long R = (long long)B * N;
  const int ntiles = EDGE ? B * tps : (int)((R + TILE - 1) / TILE);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int nrows, b = 0;
    long long row0;
    if (EDGE) {
      b = tile / tps;
      const int e0 = (tile - b * tps) * TILE;
      nrows = min(TILE, a.rowptr[(long long)b * (N + 1) + N] - e0);
      row0 = (long long)b * KMAX * N + e0;
    } else {
      row0 = (long long)tile * TILE;
      nrows = (int)min((long long)TILE, R - row0);
    }
    if (nrows <= 0) continue;          // CTA-uniform
    int step = 0;
    const int nsteps = a.nsrc * NB;
    float blk[4][8];          // the block of the step ahead
    if (!a.x8) block_issue(a.x[0], row0, nrows, Hp, 0, t, blk);
#pragma unroll 1
    for (int s = 0; s < a.nsrc; ++s) {
#pragma unroll 1
      for (int kb = 0; kb < NB; ++kb, ++step) {
        // the previous MMAs have completed (waited below): both operand buffers are free
        if (t == 0) {
          tc::mbar_expect_tx(&hdr->w_bar, WROW);
          tc::bulk_g2s(w_sm, a.img[s] + (size_t)kb * WROW, WROW, &hdr->w_bar);
        }
        if (s == 0 && a.x8)
          input_block_to_tile(a.x8, a.w8, a.b8, row0, nrows, Hp, kb * 64, t, a_hi, a_lo);
        else
          block_commit(blk, t, a_hi, a_lo);
        tc::fence_async_smem();
        tc::fence_before_sync();          // this thread's tcgen05.ld of the previous tile are complete
        __syncthreads();
        tc::mbar_wait(&hdr->w_bar, wph);
        wph ^= 1;
        if (wig == 0) {
          tc::fence_after_sync();
          const uint32_t el = tc::elect_one();
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t pa = pass == 1 ? sa_lo : sa_hi;
            const uint32_t pw = pass == 2 ? sw_lo : sw_hi;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc::mma_bf16_if(el, tmem_d, tc::make_desc(pa + k * 2 * A_LBO, A_LBO, A_SBO),
                              tc::make_desc(pw + k * 2 * BL, BL, B_SBO), idesc, (step | pass | k) != 0 ? 1u : 0u);
          }
          if (el) tc::mma_commit(&hdr->mma_bar);
          __syncwarp();
        }
        // the next step's rows: on their way while the tensor core works
        if (step + 1 < nsteps) {
          const int s1 = kb + 1 < NB ? s : s + 1, kb1 = kb + 1 < NB ? kb + 1 : 0;
          if (!(s1 == 0 && a.x8)) block_issue(a.x[s1], row0, nrows, Hp, kb1 * 64, t, blk);
        }
        tc::mbar_wait(&hdr->mma_bar, mph);
        mph ^= 1;
        tc::fence_after_sync();
      }
    }
    // epilogue: + bias + d w_d + residual, activation, out through the staging tile (the A tile: its MMAs are done)
    float d = 0.f;
    if (a.dens) {
      const long long node = EDGE ? 0 : min(row0 + r, R - 1);
      d = a.dens[EDGE ? b : (int)(node / N)] / 5000.f;          // gnn_dyn.py:158
    }
#pragma unroll 1
    for (int ob = 0; ob < NB; ++ob) {
      if (a.res) {          // the residual block, whole lines -> staging tile
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int idx = it * GT_THREADS + t;
          const int row = idx >> 4, slot = idx & 15;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < nrows) v = ld4(a.res + (row0 + row) * Hp + ob * 64 + slot * 4);
          *reinterpret_cast<float4*>(a_hi + stage_off(row, slot)) = v;
        }
        __syncthreads();
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float v[16];
        tc::tmem_ld16(taddr + ob * 64 + half * 32 + q * 16, v);
        tc::tmem_ld_wait();
        const int c0 = ob * 64 + half * 32 + q * 16;
        if (a.res) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 u = *reinterpret_cast<const float4*>(a_hi + stage_off(r, ((half * 32 + q * 16) >> 2) + i));
            v[4 * i] += u.x; v[4 * i + 1] += u.y; v[4 * i + 2] += u.z; v[4 * i + 3] += u.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 add = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + c0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.wd) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.wd + c0 + j));
            add.x = fmaf(d, w4.x, add.x); add.y = fmaf(d, w4.y, add.y); add.z = fmaf(d, w4.z, add.z); add.w = fmaf(d, w4.w, add.w);
          }
          v[j] += add.x; v[j + 1] += add.y; v[j + 2] += add.z; v[j + 3] += add.w;
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        stage_put16(a_hi, r, half * 32 + q * 16, v);          // the slots this thread just read
      }
      __syncthreads();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * GT_THREADS + t;
        const int row = idx >> 4, slot = idx & 15;
        const float4 v = *reinterpret_cast<const float4*>(a_hi + stage_off(row, slot));
        if (row < nrows) st4(a.y + (row0 + row) * Hp + ob * 64 + slot * 4, v);
      }
      __syncthreads();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (t < 32) tc::tmem_dealloc(hdr->tmem_base, TCOLS);
}

size_t tc_image_floats(int Hp) { return (size_t)Hp * Hp; }          // hi + lo bf16 per weight = one float's worth

// `n` images in one launch: image i from wpack + slot_off[i] to img0 + i * Hp * Hp floats
int launch_tc_images(const float* wpack, const long long* slot_off, int n, int Hp, float* img0, cudaStream_t st) {
  if (n > TcImages::MAXN) return (int)cudaErrorInvalidValue;
  TcImages im;
  for (int i = 0; i < n; ++i) im.src[i] = wpack + slot_off[i];
  const dim3 grid((Hp * Hp + 255) / 256, n);
  k_g_tc_image<<<grid, 256, 0, st>>>(im, Hp, reinterpret_cast<__nv_bfloat16*>(img0));
  PILE_CHECK_LAUNCH();
  return 0;
}

template <int NB, bool EDGE>
static int launch_lin_tc_nb(const LinTcArgs& a, cudaStream_t st) {
  static DeviceOnce once;
  const int dev = once.pending();
  if (dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(k_g_lin_tc<NB, EDGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gt_smem(64 * NB));
    if (e != cudaSuccess) return (int)e;
    once.done(dev);
  }
  const long long ntiles = EDGE ? (long long)a.B * ((KMAX * a.N + TILE - 1) / TILE) : ((long long)a.B * a.N + TILE - 1) / TILE;
  const int grid = (int)(ntiles < 1 ? 1 : (ntiles < 2 * NSM ? ntiles : 2 * NSM));
  k_g_lin_tc<NB, EDGE><<<grid, GT_THREADS, gt_smem(64 * NB), st>>>(a);
  PILE_CHECK_LAUNCH();
  return 0;
}

template <bool EDGE>
static int launch_lin_tc_any(const LinTcArgs& a, int Hp, cudaStream_t st) {
  switch (Hp / 64) {
    case 1: return launch_lin_tc_nb<1, EDGE>(a, st);
    case 2: return launch_lin_tc_nb<2, EDGE>(a, st);
    case 3: return launch_lin_tc_nb<3, EDGE>(a, st);
    case 4: return launch_lin_tc_nb<4, EDGE>(a, st);
  }
  return (int)cudaErrorInvalidValue;
}

int launch_lin_tc_edge(const float* x, const float* img, const float* bias, const float* wd, const float* dens, int relu,
                       float* y, const int* rowptr, int B, int N, int Hp, cudaStream_t st, const float* x8, const float* w8,
                       const float* b8) {
  LinTcArgs a{};
  a.x8 = x8; a.w8 = w8; a.b8 = b8;
  a.nsrc = 1; a.x[0] = x; a.img[0] = reinterpret_cast<const uint8_t*>(img);
  a.bias = bias; a.wd = wd; a.dens = dens; a.res = nullptr; a.y = y; a.rowptr = rowptr; a.relu = relu; a.B = B; a.N = N;
  return launch_lin_tc_any<true>(a, Hp, st);
}

// particle rows: up to two sources, optional residual
int launch_lin_tc_node(const float* x0, const float* img0, const float* x1, const float* img1, const float* bias,
                       const float* wd, const float* dens, const float* res, int relu, float* y, int B, int N, int Hp,
                       cudaStream_t st) {
  LinTcArgs a{};
  a.nsrc = x1 ? 2 : 1;
  a.x[0] = x0; a.img[0] = reinterpret_cast<const uint8_t*>(img0);
  a.x[1] = x1; a.img[1] = reinterpret_cast<const uint8_t*>(img1);
  a.bias = bias; a.wd = wd; a.dens = dens; a.res = res; a.y = y; a.rowptr = nullptr; a.relu = relu; a.B = B; a.N = N;
  return launch_lin_tc_any<false>(a, Hp, st);
}

}  // namespace general
}  // namespace pile
