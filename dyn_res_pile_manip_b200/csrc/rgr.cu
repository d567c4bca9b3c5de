// Resolution regressor inference (reference model/res_regressor.py:106-144, MPCResRgrNoPool.forward):
//   5 x [Conv2d(k=4, s=2, p=1) + LeakyReLU(0.2)]  6-64-128-256-512-512 on a 224 x 224 input  ->  512 x 7 x 7
//   Flatten, Linear 25088-4096-1024-256-64-1 with LeakyReLU(0.2) between.
// It runs once per MPC step between the simulator and the planner (env/flex_env.py:981-998, 1080-1090) at batch 1,
// so it is a WEIGHT STREAM: 114.2 M parameters = 457 MB, of which the first linear layer alone is 411 MB.
//   * convolutions: implicit GEMM on the CUDA cores, 64 output channels x 64 output pixels per CTA, the reduction
//     (C_in * 16) split over several CTAs so that every layer fills the GPU; partial sums are combined in a fixed
//     order by k_conv_finish (deterministic), which also adds the bias and applies the activation;
//   * linear layers: one warp per output row, 128-bit loads, the row is read exactly once (HBM roofline).
// FP32 throughout (the reference truncates the output to an int, so no reduced precision here).
#include "common.cuh"
#include "kernels.h"

namespace pile {

constexpr int RGR_NCONV = 5;
constexpr int RGR_NFC = 5;
static const int kConvCh[RGR_NCONV + 1] = {6, 64, 128, 256, 512, 512};
static const int kFcW[RGR_NFC + 1] = {512 * 7 * 7, 4096, 1024, 256, 64, 1};
constexpr float LRELU = 0.2f;

// parameter buffer = the reference state_dict order: model.0.weight, model.0.bias, model.2.weight, ... (conv), then
// model.11.weight, model.11.bias, model.13..., model.19 (linear)
long long rgr_param_offset(int idx) {      // idx = 2 * layer + (0 weight | 1 bias); idx = 2 * 10 -> total
  long long off = 0;
  for (int l = 0; l < RGR_NCONV + RGR_NFC; ++l) {
    long long w, b;
    if (l < RGR_NCONV) { w = (long long)kConvCh[l + 1] * kConvCh[l] * 16; b = kConvCh[l + 1]; }
    else { w = (long long)kFcW[l - RGR_NCONV + 1] * kFcW[l - RGR_NCONV]; b = kFcW[l - RGR_NCONV + 1]; }
    if (idx == 2 * l) return off;
    off += w;
    if (idx == 2 * l + 1) return off;
    off += b;
  }
  return off;
}

// ------------------------------------------------------------------------------------------------
// conv 4x4 / stride 2 / pad 1 as implicit GEMM.  in [B, Cin, Hi, Wi], w [Cout, Cin, 4, 4],
// part [S, Cout, B*Ho*Wo]; CTA (x = pixel tile, y = cout tile, z = split) reduces input channels
// [z * cin_per, (z+1) * cin_per).
// ------------------------------------------------------------------------------------------------
constexpr int CT = 64;        // cout tile = pixel tile
constexpr int CKC = 4;        // input channels per shared-memory stage (K = 64)
constexpr int CONV_THREADS = 256;

__global__ void __launch_bounds__(CONV_THREADS)
k_conv4x4s2(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ part, int B, int Cin,
            int Hi, int Wi, int Cout, int Ho, int Wo, int cin_per) {
  __shared__ __align__(16) float Ws[CKC * 16][CT];     // [k][cout]
  __shared__ __align__(16) float Xs[CKC * 16][CT];     // [k][pixel]
  const int P = B * Ho * Wo;
  const int p0 = blockIdx.x * CT, c0 = blockIdx.y * CT;
  const int ci_lo = blockIdx.z * cin_per, ci_hi = min(Cin, ci_lo + cin_per);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // thread -> 4 pixels (tx) x 4 couts (ty)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // this thread's gather coordinates: it loads pixel column (threadIdx.x & 63) for 16 of the 64 k rows
  const int lp = threadIdx.x & 63, lk0 = threadIdx.x >> 6;      // k rows lk0, lk0 + 4, ...
  const int p = p0 + lp;
  const bool pv = p < P;
  int b = 0, oy = 0, ox = 0;
  if (pv) { b = p / (Ho * Wo); const int r = p - b * Ho * Wo; oy = r / Wo; ox = r - oy * Wo; }
  const float* inb = in + (long long)b * Cin * Hi * Wi;

  for (int ci = ci_lo; ci < ci_hi; ci += CKC) {
    // weights: Ws[k][c] = w[c0 + c][ci + k / 16][k % 16]; a warp reads 64 consecutive floats of one cout row
    for (int idx = threadIdx.x; idx < CKC * 16 * CT; idx += CONV_THREADS) {
      const int c = idx >> 6, k = idx & 63;
      const int cin = ci + (k >> 4);
      float v = 0.f;
      if (c0 + c < Cout && cin < ci_hi) v = __ldg(w + ((long long)(c0 + c) * Cin + cin) * 16 + (k & 15));
      Ws[k][c] = v;
    }
#pragma unroll 4
    for (int kk = 0; kk < 16; ++kk) {
      const int k = lk0 + kk * 4;
      const int cin = ci + (k >> 4), ky = (k >> 2) & 3, kx = k & 3;
      const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
      float v = 0.f;
      if (pv && cin < ci_hi && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi) v = __ldg(inb + ((long long)cin * Hi + iy) * Wi + ix);
      Xs[k][lp] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < CKC * 16; ++k) {
      const float4 wv = *reinterpret_cast<const float4*>(&Ws[k][ty * 4]);
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[k][tx * 4]);
      const float wa[4] = {wv.x, wv.y, wv.z, wv.w}, xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wa[i], xa[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = part + (long long)blockIdx.z * Cout * P;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty * 4 + i;
    if (c >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pp = p0 + tx * 4 + j;
      if (pp < P) out[(long long)c * P + pp] = acc[i][j];
    }
  }
}

// out[b][c][pix] = lrelu(bias[c] + sum_s part[s][c][b * HW + pix])   (fixed summation order)
__global__ void k_conv_finish(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ out,
                              int S, int Cout, int B, int HW) {
  const long long P = (long long)B * HW, total = (long long)Cout * P;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i / P);
  const long long p = i - (long long)c * P;
  float v = bias[c];
  for (int s = 0; s < S; ++s) v += part[(long long)s * total + i];
  v = v > 0.f ? v : LRELU * v;
  const int b = (int)(p / HW);
  const int pix = (int)(p - (long long)b * HW);
  out[((long long)b * Cout + c) * HW + pix] = v;
}

// ------------------------------------------------------------------------------------------------
// linear layer at small batch: y[b][o] = act(bias[o] + W[o, :] . x[b, :]); one warp per output row, up to 4 batch
// rows per pass (the weight row is read once per pass)
// ------------------------------------------------------------------------------------------------
constexpr int FC_THREADS = 256;

template <int NB>
__global__ void __launch_bounds__(FC_THREADS)
k_fc_rows(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ x, float* __restrict__ y,
          int In, int Out, int b0, int act) {
  const int warp = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= Out) return;
  const float4* wr = reinterpret_cast<const float4*>(W + (long long)warp * In);
  const int n4 = In >> 2;
  float acc[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b] = 0.f;
#pragma unroll 4
  for (int i = lane; i < n4; i += 32) {
    float4 wv;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(wv.x), "=f"(wv.y), "=f"(wv.z), "=f"(wv.w) : "l"(wr + i));
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)(b0 + b) * In) + i);
      acc[b] = fmaf(wv.x, xv.x, acc[b]);
      acc[b] = fmaf(wv.y, xv.y, acc[b]);
      acc[b] = fmaf(wv.z, xv.z, acc[b]);
      acc[b] = fmaf(wv.w, xv.w, acc[b]);
    }
  }
  for (int i = (n4 << 2) + lane; i < In; i += 32) {           // tail (In % 4 != 0): none of the shipped layers
    const float wv = W[(long long)warp * In + i];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = fmaf(wv, x[(long long)(b0 + b) * In + i], acc[b]);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
    if (lane == 0) {
      float v = acc[b] + bias[warp];
      if (act) v = v > 0.f ? v : LRELU * v;
      y[(long long)(b0 + b) * Out + warp] = v;
    }
  }
}

static int conv_out(int n) { return (n + 2 - 4) / 2 + 1; }

struct RgrPlan {
  int hi[RGR_NCONV + 1], wi[RGR_NCONV + 1];
  int splits[RGR_NCONV];
  size_t part_floats;      // split-K partial sums (max over layers)
  size_t act_floats;       // one activation buffer (max over layers), two are used in ping-pong
  bool ok;
};

static RgrPlan rgr_plan(int B, int H, int W) {
  RgrPlan p{};
  p.hi[0] = H; p.wi[0] = W;
  p.part_floats = 0; p.act_floats = 0;
  for (int l = 0; l < RGR_NCONV; ++l) {
    p.hi[l + 1] = conv_out(p.hi[l]);
    p.wi[l + 1] = conv_out(p.wi[l]);
    const long long P = (long long)B * p.hi[l + 1] * p.wi[l + 1];
    const long long tiles = ((P + CT - 1) / CT) * ((kConvCh[l + 1] + CT - 1) / CT);
    const int max_splits = (kConvCh[l] + CKC - 1) / CKC;
    long long s = (2 * NSM + tiles - 1) / tiles;               // about two CTAs per SM
    if (s > max_splits) s = max_splits;
    if (s > 32) s = 32;
    if (s < 1) s = 1;
    p.splits[l] = (int)s;
    const size_t part = (size_t)s * kConvCh[l + 1] * P;
    if (part > p.part_floats) p.part_floats = part;
    const size_t act = (size_t)kConvCh[l + 1] * P;
    if (act > p.act_floats) p.act_floats = act;
  }
  for (int l = 0; l < RGR_NFC; ++l) {
    const size_t act = (size_t)B * kFcW[l + 1];
    if (act > p.act_floats) p.act_floats = act;
  }
  p.ok = B > 0 && H > 0 && W > 0 && p.hi[RGR_NCONV] > 0 && p.wi[RGR_NCONV] > 0 &&
         (long long)kConvCh[RGR_NCONV] * p.hi[RGR_NCONV] * p.wi[RGR_NCONV] == kFcW[0];
  return p;
}

static size_t up256(size_t x) { return (x + 255) / 256 * 256; }

long long rgr_workspace_bytes(int B, int H, int W) {
  const RgrPlan p = rgr_plan(B, H, W);
  if (!p.ok) return -1;
  return (long long)(up256(p.part_floats * 4) + 2 * up256(p.act_floats * 4));
}

int launch_rgr_forward(const float* params, const float* x, int B, int H, int W, void* ws, float* y, cudaStream_t st) {
  const RgrPlan p = rgr_plan(B, H, W);
  if (!p.ok) return (int)cudaErrorInvalidValue;
  char* base = static_cast<char*>(ws);
  float* part = reinterpret_cast<float*>(base);
  float* act[2] = {reinterpret_cast<float*>(base + up256(p.part_floats * 4)),
                   reinterpret_cast<float*>(base + up256(p.part_floats * 4) + up256(p.act_floats * 4))};
  const float* cur = x;
  int flip = 0;
  for (int l = 0; l < RGR_NCONV; ++l) {
    const int Cin = kConvCh[l], Cout = kConvCh[l + 1];
    const int Ho = p.hi[l + 1], Wo = p.wi[l + 1];
    const long long P = (long long)B * Ho * Wo;
    const int S = p.splits[l];
    const int cin_per = ((Cin + S - 1) / S + CKC - 1) / CKC * CKC;
    const int S_eff = (Cin + cin_per - 1) / cin_per;
    dim3 grid((unsigned)((P + CT - 1) / CT), (unsigned)((Cout + CT - 1) / CT), (unsigned)S_eff);
    k_conv4x4s2<<<grid, CONV_THREADS, 0, st>>>(cur, params + rgr_param_offset(2 * l), part, B, Cin, p.hi[l], p.wi[l],
                                              Cout, Ho, Wo, cin_per);
    PILE_CHECK_LAUNCH();
    const long long total = (long long)Cout * P;
    k_conv_finish<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part, params + rgr_param_offset(2 * l + 1), act[flip],
                                                                   S_eff, Cout, B, Ho * Wo);
    PILE_CHECK_LAUNCH();
    cur = act[flip];
    flip ^= 1;
  }
  for (int l = 0; l < RGR_NFC; ++l) {
    const int In = kFcW[l], Out = kFcW[l + 1];
    const float* Wt = params + rgr_param_offset(2 * (RGR_NCONV + l));
    const float* bs = params + rgr_param_offset(2 * (RGR_NCONV + l) + 1);
    float* out = l == RGR_NFC - 1 ? y : act[flip];
    const int act_fn = l < RGR_NFC - 1;
    const unsigned blocks = (unsigned)(((long long)Out * 32 + FC_THREADS - 1) / FC_THREADS);
    for (int b0 = 0; b0 < B;) {
      const int nb = B - b0 >= 4 ? 4 : (B - b0);
      if (nb == 4) k_fc_rows<4><<<blocks, FC_THREADS, 0, st>>>(Wt, bs, cur, out, In, Out, b0, act_fn);
      else if (nb == 3) k_fc_rows<3><<<blocks, FC_THREADS, 0, st>>>(Wt, bs, cur, out, In, Out, b0, act_fn);
      else if (nb == 2) k_fc_rows<2><<<blocks, FC_THREADS, 0, st>>>(Wt, bs, cur, out, In, Out, b0, act_fn);
      else k_fc_rows<1><<<blocks, FC_THREADS, 0, st>>>(Wt, bs, cur, out, In, Out, b0, act_fn);
      PILE_CHECK_LAUNCH();
      b0 += nb;
    }
    cur = out;
    flip ^= 1;
  }
  return 0;
}

}  // namespace pile
