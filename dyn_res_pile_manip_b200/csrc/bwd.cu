// K8: dgrad-only backward of one model step and of the pusher model.
//
// The planner differentiates sum(-reward) w.r.t. the ACTIONS only (reference planners.py:742-745; the
// optimizer owns act_seqs_tensor alone, :674), so no weight gradients are needed and every ReLU is
// replayed from the sign bits the forward recorded (one byte per row and 8-column block).  No gradient
// flows through the relation sets (topk/nonzero) or the hard along-push mask, exactly as in autograd.
// All scatters are written as gathers over the receiver-major CSR and its sender-major transpose, so the
// result is deterministic (no float atomics).
//
// With g = dL/ds_pred:
//   g_q    = (g V1) * [q>0]                  g_eff3 = g_q V0
//   for p = 2,1,0:  g_z = g_eff_{p+1} * [eff_{p+1}>0];  g_Cp += g_z;  g_agg_p = g_z W_a
//                   g_m(e) = g_agg_p[recv e] * [m_p(e)>0]
//                   g_eff_p = g_z + (sum_{e in row i} g_m) W_r + (sum_{e: send e = i} g_m) W_s
//   g_penc = g_eff_0 + g_Cp W_p ;  g_sdelta = (((g_penc*[penc>0]) PE1)*[h>0]) PE0[:, :3]
//   g_Ce(e) = sum_p g_m_p(e);  g_x(e) = ((((g_Ce W_e)*[r3>0]) RE2)*[r2>0]) RE1)*[r1>0]) RE0[:, 2:5]
//   g_scur[i] = g[i] + sum_{e in row i} g_x(e) - sum_{e: send e = i} g_x(e)
#include "common.cuh"
#include "kernels.h"

namespace pile {

namespace {
constexpr size_t BALIGN = 256;
inline size_t bup(size_t x) { return (x + BALIGN - 1) / BALIGN * BALIGN; }
}  // namespace

struct BwdScratch {
  float* gz;        // [R,H] current g_z
  float* gcp;       // [R,H] accumulated g_Cp
  float* gagg[PSTEP];
  float* gx;        // [E,4]
  float* gpr;       // [R,H] receiver-side sums of the masked relation gradients of the current propagation step
  float* gps;       // [R,H] sender-side sums
  size_t bytes;
};

static BwdScratch carve_bwd(void* p, int B, int N) {
  const size_t R = (size_t)B * N, E = (size_t)B * KMAX * N;
  char* base = static_cast<char*>(p);
  size_t off = 0;
  auto take = [&](size_t n) { float* q = base ? reinterpret_cast<float*>(base + off) : nullptr; off += bup(n * 4); return q; };
  BwdScratch s;
  s.gz = take(R * H);
  s.gcp = take(R * H);
  for (int i = 0; i < PSTEP; ++i) s.gagg[i] = take(R * H);
  s.gx = take(E * 4);
  s.gpr = take(R * H);
  s.gps = take(R * H);
  s.bytes = off;
  return s;
}

size_t bwd_scratch_bytes(int B, int N) { return carve_bwd(nullptr, B, N).bytes; }

__device__ __forceinline__ void acc_zero(float (&acc)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

__device__ __forceinline__ void acc_load_global(const float* __restrict__ G, long long row0, int nrows, int lane,
                                                int c0, float (&acc)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = lane + 32 * i;
    if (r < nrows) {
      const float4 a = ld4(G + (row0 + r) * H + c0), b = ld4(G + (row0 + r) * H + c0 + 4);
      acc[i][0] = a.x; acc[i][1] = a.y; acc[i][2] = a.z; acc[i][3] = a.w;
      acc[i][4] = b.x; acc[i][5] = b.y; acc[i][6] = b.z; acc[i][7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
  }
}

__device__ __forceinline__ void acc_add_global(float* __restrict__ G, long long row0, int nrows, int lane, int c0,
                                               const float (&acc)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = lane + 32 * i;
    if (r < nrows) {
      float* p = G + (row0 + r) * H + c0;
      float4 a = ld4(p), b = ld4(p + 4);
      a.x += acc[i][0]; a.y += acc[i][1]; a.z += acc[i][2]; a.w += acc[i][3];
      b.x += acc[i][4]; b.y += acc[i][5]; b.z += acc[i][6]; b.w += acc[i][7];
      st4(p, a); st4(p + 4, b);
    }
  }
}

__device__ __forceinline__ float4 mask4(const float4 v, unsigned nib) {
  return make_float4(nib & 1u ? v.x : 0.f, nib & 2u ? v.y : 0.f, nib & 4u ? v.z : 0.f, nib & 8u ? v.w : 0.f);
}

// ------------------------------------------------------------------------------------------------
// head: predictor backward, then g_z / g_agg of propagation step 2
// ------------------------------------------------------------------------------------------------
struct BwdHeadSmem {
  float w_v0[H * H], w_a[H * H];
  float w_v1[4 * H];
  float A[TILE * LDA];
  float G[TILE * 4];
};

__global__ void __launch_bounds__(NT, 2)
k_bwd_head(const float* __restrict__ wpack, const float* __restrict__ g_pred, long long g_stride,
           const uint8_t* __restrict__ m_q, const uint8_t* __restrict__ m_eff2, float* __restrict__ gz,
           float* __restrict__ gcp, float* __restrict__ gagg2, int B, int N) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdHeadSmem& S = *reinterpret_cast<BwdHeadSmem*>(smem_raw);
  load_block(S.w_v0, wpack + wslot_offset(W_V0), H * H);
  load_block(S.w_a, wpack + wslot_offset(W_A), H * H);
  load_block(S.w_v1, wpack + wslot_offset(W_V1), 4 * H);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c0 = warp * 8;
  const long long R = (long long)B * N;
  const int ntiles = (int)((R + TILE - 1) / TILE);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * TILE;
    const int nrows = (int)min((long long)TILE, R - row0);
    __syncthreads();
    if (threadIdx.x < TILE) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((int)threadIdx.x < nrows) {
        const long long node = row0 + threadIdx.x;
        const float* p = g_pred + (node / N) * g_stride + (node % N) * 3;
        g = make_float4(p[0], p[1], p[2], 0.f);
      }
      st4(S.G + threadIdx.x * 4, g);
    }
    __syncthreads();
    float acc[4][8];
    unsigned bits[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 g = ld4(S.G + (lane + 32 * i) * 4);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        acc[i][j] = g.x * S.w_v1[0 * H + c0 + j] + g.y * S.w_v1[1 * H + c0 + j] + g.z * S.w_v1[2 * H + c0 + j];
    }
    mask_from_global(m_q, row0, nrows, lane, warp, bits);
    acc_apply_mask(acc, bits);
    acc_to_smem(S.A, LDA, lane, c0, acc);
    __syncthreads();
    acc_zero(acc);
    gemm_rows4x8<H>(S.A, LDA, S.w_v0, lane, c0, acc);          // g_eff3
    mask_from_global(m_eff2, row0, nrows, lane, warp, bits);
    acc_apply_mask(acc, bits);                                  // g_z(2)
    acc_to_global(gz, row0, nrows, lane, c0, acc);
    acc_to_global(gcp, row0, nrows, lane, c0, acc);
    __syncthreads();
    acc_to_smem(S.A, LDA, lane, c0, acc);
    __syncthreads();
    acc_zero(acc);
    gemm_rows4x8<H>(S.A, LDA, S.w_a, lane, c0, acc);
    acc_to_global(gagg2, row0, nrows, lane, c0, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// propagation step p backward (gathers + node GEMMs), then either g_z/g_agg of step p-1 or (FIRST) the
// particle-encoder backward producing g_s_delta
// ------------------------------------------------------------------------------------------------
// g_Pr[i] = sum_{e in row i} g_m(e),  g_Ps[i] = sum_{e: send e = i} g_m(e)  with  g_m(e) = g_agg[recv e] * [m(e) > 0].
// Latency-bound gathers (mask bytes, transposed CSR, rows of g_agg), so they run in their own high-occupancy
// kernel (a half-warp per particle, lane = 4 channels) instead of inside the one-CTA-per-SM GEMM kernel.  The
// sums run in CSR order: deterministic.
__global__ void __launch_bounds__(256)
k_bwd_gather(const int* __restrict__ rowptr, const int* __restrict__ trowptr, const int* __restrict__ trecv,
             const int* __restrict__ tedge, const uint8_t* __restrict__ m_edge, const float* __restrict__ gagg_p,
             float* __restrict__ gpr_out, float* __restrict__ gps_out, int B, int N) {
  const int l16 = threadIdx.x & 15;
  // a lane's four channels are one nibble of mask byte l16 / 2 of a relation's row: the low one on even lanes
  const unsigned odd = (unsigned)l16 & 1u;
  const unsigned sel_nib = odd ? 0xf0u : 0x0fu, sel_cnt = odd ? 0x10101010u : 0x01010101u, sh = odd * 4u;
  const int R = B * N;
  const int nhw = (int)gridDim.x * (int)(blockDim.x >> 4);
  for (int node = (int)blockIdx.x * (int)(blockDim.x >> 4) + (int)(threadIdx.x >> 4); node < R; node += nhw) {
    const int b = node / N, i = node - b * N;
    const long long slot = (long long)b * KMAX * N;
    const int* rp = rowptr + (long long)b * (N + 1) + i;
    const int* tp = trowptr + (long long)b * (N + 1) + i;
    const int e0 = rp[0], e1 = rp[1], k0 = tp[0], k1 = tp[1];
    const float4 ga = ld4(gagg_p + (long long)node * H + 4 * l16);
    // per-sample base pointers once; everything below is a 32-bit index on top of them
    const uint8_t* msample = m_edge + slot * 8 + (l16 >> 1);
    const float* gsample = gagg_p + (long long)b * N * H + 4 * l16;
    const int* tr = trecv + slot;
    const int* te = tedge + slot;
    // (kept opaque: ptxas otherwise re-derives every address from the kernel parameters, three instructions per load)
    asm volatile("" : "+l"(msample), "+l"(gsample), "+l"(tr), "+l"(te));
    // Receiver side: every relation of the row carries the SAME gradient row ga, so the masked sum is ga times the number
    // of relations whose sign bit is set, per channel.  The four bits of the lane's nibble are spread to the four bytes of
    // a word (multiplier 1 + 2^7 + 2^14 + 2^21: bit k of the nibble lands on bit 8k, the copies do not overlap) and the
    // words are added: four counters (<= 10) in one register.
    unsigned cnt4 = 0;
    {
      const uint8_t* mp = msample + (unsigned)e0 * 8u;
      asm volatile("" : "+l"(mp));
      const int cnt = e1 - e0;
#pragma unroll
      for (int q = 0; q < KMAX; ++q)
        if (q < cnt) cnt4 += (((unsigned)__ldg(mp + q * 8) & sel_nib) * 0x00204081u) & sel_cnt;
#pragma unroll 1
      for (int q = KMAX; q < cnt; ++q)          // denser caller-provided lists never reach this kernel; kept for safety
        cnt4 += (((unsigned)__ldg(mp + q * 8) & sel_nib) * 0x00204081u) & sel_cnt;
    }
    const float4 gpr = make_float4(ga.x * (float)((cnt4 >> sh) & 15u), ga.y * (float)((cnt4 >> (8u + sh)) & 15u),
                                   ga.z * (float)((cnt4 >> (16u + sh)) & 15u), ga.w * (float)((cnt4 >> (24u + sh)) & 15u));
    float4 gps = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = k0; k < k1; k += 4) {          // four independent gathers in flight
      float4 g[4];
      unsigned nib[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kk = min(k + u, k1 - 1);
        const unsigned rc = (unsigned)__ldg(tr + kk), e = (unsigned)__ldg(te + kk);
        nib[u] = (unsigned)__ldg(msample + e * 8u) >> sh;
        g[u] = __ldg(reinterpret_cast<const float4*>(gsample + rc * (unsigned)H));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k + u < k1) {
          if (nib[u] & 1u) gps.x += g[u].x;
          if (nib[u] & 2u) gps.y += g[u].y;
          if (nib[u] & 4u) gps.z += g[u].z;
          if (nib[u] & 8u) gps.w += g[u].w;
        }
      }
    }
    st4(gpr_out + (long long)node * H + 4 * l16, gpr);
    st4(gps_out + (long long)node * H + 4 * l16, gps);
  }
}

struct BwdPropSmem {
  float w_r[H * H], w_s[H * H], w_x[H * H], w_y[H * H];   // w_x: W_a (or W_p when FIRST); w_y: W_PE1 (FIRST)
  float w_pe0[8 * H];
  float A1[TILE * LDA];
  float A2[TILE * LDA];
};

template <bool FIRST>
__global__ void __launch_bounds__(NT, 1)
k_bwd_prop(const float* __restrict__ wpack, const float* __restrict__ gpr_in, const float* __restrict__ gps_in,
           const uint8_t* __restrict__ m_next /* eff[p-1] or pe1 */, const uint8_t* __restrict__ m_pe0,
           float* __restrict__ gz, float* __restrict__ gcp, float* __restrict__ gagg_out,
           float* __restrict__ g_s_delta, int B, int N) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdPropSmem& S = *reinterpret_cast<BwdPropSmem*>(smem_raw);
  load_block(S.w_r, wpack + wslot_offset(W_R), H * H);
  load_block(S.w_s, wpack + wslot_offset(W_S), H * H);
  if (FIRST) {
    load_block(S.w_x, wpack + wslot_offset(W_P), H * H);
    load_block(S.w_y, wpack + wslot_offset(W_PE1), H * H);
    load_block(S.w_pe0, wpack + wslot_offset(W_PE0), 8 * H);
  } else {
    load_block(S.w_x, wpack + wslot_offset(W_A), H * H);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c0 = warp * 8;
  const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
  const long long R = (long long)B * N;
  const int ntiles = (int)((R + TILE - 1) / TILE);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * TILE;
    const int nrows = (int)min((long long)TILE, R - row0);
    __syncthreads();
    for (int r = hw; r < TILE; r += NT / 16) {          // the gathered sums of this tile (k_bwd_gather)
      float4 gpr = make_float4(0.f, 0.f, 0.f, 0.f), gps = gpr;
      if (r < nrows) {
        gpr = ld4(gpr_in + (row0 + r) * H + 4 * l16);
        gps = ld4(gps_in + (row0 + r) * H + 4 * l16);
      }
      st4(S.A1 + r * LDA + 4 * l16, gpr);
      st4(S.A2 + r * LDA + 4 * l16, gps);
    }
    __syncthreads();
    float acc[4][8];
    unsigned bits[4];
    acc_load_global(gz, row0, nrows, lane, c0, acc);            // residual path: g_eff_p starts at g_z(p)
    gemm_rows4x8<H>(S.A1, LDA, S.w_r, lane, c0, acc);
    gemm_rows4x8<H>(S.A2, LDA, S.w_s, lane, c0, acc);           // acc = g_eff_p
    if (!FIRST) {
      mask_from_global(m_next, row0, nrows, lane, warp, bits);
      acc_apply_mask(acc, bits);                                // g_z(p-1)
      acc_to_global(gz, row0, nrows, lane, c0, acc);
      acc_add_global(gcp, row0, nrows, lane, c0, acc);
      __syncthreads();
      acc_to_smem(S.A1, LDA, lane, c0, acc);
      __syncthreads();
      acc_zero(acc);
      gemm_rows4x8<H>(S.A1, LDA, S.w_x, lane, c0, acc);
      acc_to_global(gagg_out, row0, nrows, lane, c0, acc);
    } else {
      // g_penc = g_eff_0 + g_Cp W_p
      __syncthreads();
      {
        float t[4][8];
        acc_load_global(gcp, row0, nrows, lane, c0, t);
        acc_to_smem(S.A1, LDA, lane, c0, t);
      }
      __syncthreads();
      gemm_rows4x8<H>(S.A1, LDA, S.w_x, lane, c0, acc);
      mask_from_global(m_next, row0, nrows, lane, warp, bits);  // pe1
      acc_apply_mask(acc, bits);
      acc_to_smem(S.A2, LDA, lane, c0, acc);
      __syncthreads();
      acc_zero(acc);
      gemm_rows4x8<H>(S.A2, LDA, S.w_y, lane, c0, acc);
      mask_from_global(m_pe0, row0, nrows, lane, warp, bits);
      acc_apply_mask(acc, bits);
      acc_to_smem(S.A1, LDA, lane, c0, acc);
      __syncthreads();
      if ((int)threadIdx.x < nrows) {
        const int r = threadIdx.x;
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 4
        for (int k = 0; k < H; k += 4) {
          const float4 a = ld4(S.A1 + r * LDA + k);
          const float* w = S.w_pe0 + k * 8;
          o0 = fmaf(a.x, w[0], o0);  o1 = fmaf(a.x, w[1], o1);  o2 = fmaf(a.x, w[2], o2);
          o0 = fmaf(a.y, w[8], o0);  o1 = fmaf(a.y, w[9], o1);  o2 = fmaf(a.y, w[10], o2);
          o0 = fmaf(a.z, w[16], o0); o1 = fmaf(a.z, w[17], o1); o2 = fmaf(a.z, w[18], o2);
          o0 = fmaf(a.w, w[24], o0); o1 = fmaf(a.w, w[25], o1); o2 = fmaf(a.w, w[26], o2);
        }
        float* o = g_s_delta + (row0 + r) * 3;
        o[0] = o0; o[1] = o1; o[2] = o2;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// relation encoder backward: g_x(e) [E,4] (xyz of the s_r - s_s feature)
// ------------------------------------------------------------------------------------------------
struct BwdEdgeSmem {
  float w_e[H * H], w_re2[H * H], w_re1[H * H];
  float w_re0[8 * H];
  float A[TILE * LDA];
};

__global__ void __launch_bounds__(NT, 2)
k_bwd_edge(const float* __restrict__ wpack, const int* __restrict__ rowptr, const int* __restrict__ row,
           const uint8_t* __restrict__ me0, const uint8_t* __restrict__ me1, const uint8_t* __restrict__ me2,
           const uint8_t* __restrict__ m_re0, const uint8_t* __restrict__ m_re1, const uint8_t* __restrict__ m_re2,
           const float* __restrict__ ga0, const float* __restrict__ ga1, const float* __restrict__ ga2,
           float* __restrict__ gx, int B, int N) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdEdgeSmem& S = *reinterpret_cast<BwdEdgeSmem*>(smem_raw);
  load_block(S.w_e, wpack + wslot_offset(W_E), H * H);
  load_block(S.w_re2, wpack + wslot_offset(W_RE2), H * H);
  load_block(S.w_re1, wpack + wslot_offset(W_RE1), H * H);
  load_block(S.w_re0, wpack + wslot_offset(W_RE0), 8 * H);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c0 = warp * 8;
  const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
  const int mb = l16 >> 1, msh = (l16 & 1) * 4;
  const int tps = (KMAX * N + TILE - 1) / TILE;
  const long long ntiles = (long long)B * tps;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = (int)(tile / tps);
    const int e0 = (int)(tile % tps) * TILE;
    const int ne = rowptr[(long long)b * (N + 1) + N];
    if (e0 >= ne) continue;
    const int nrows = min(TILE, ne - e0);
    const long long slot0 = (long long)b * KMAX * N + e0;
    __syncthreads();
    for (int r = hw; r < TILE; r += NT / 16) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows) {
        const long long node = (long long)b * N + row[slot0 + r];
        const float4 a0 = mask4(ld4(ga0 + node * H + 4 * l16), (me0[(slot0 + r) * 8 + mb] >> msh) & 15u);
        const float4 a1 = mask4(ld4(ga1 + node * H + 4 * l16), (me1[(slot0 + r) * 8 + mb] >> msh) & 15u);
        const float4 a2 = mask4(ld4(ga2 + node * H + 4 * l16), (me2[(slot0 + r) * 8 + mb] >> msh) & 15u);
        g = make_float4(a0.x + a1.x + a2.x, a0.y + a1.y + a2.y, a0.z + a1.z + a2.z, a0.w + a1.w + a2.w);
      }
      st4(S.A + r * LDA + 4 * l16, g);
    }
    __syncthreads();
    float acc[4][8];
    unsigned bits[4];
    acc_zero(acc);
    gemm_rows4x8<H>(S.A, LDA, S.w_e, lane, c0, acc);
    mask_from_global(m_re2, slot0, nrows, lane, warp, bits);
    acc_apply_mask(acc, bits);
    __syncthreads();
    acc_to_smem(S.A, LDA, lane, c0, acc);
    __syncthreads();
    acc_zero(acc);
    gemm_rows4x8<H>(S.A, LDA, S.w_re2, lane, c0, acc);
    mask_from_global(m_re1, slot0, nrows, lane, warp, bits);
    acc_apply_mask(acc, bits);
    __syncthreads();
    acc_to_smem(S.A, LDA, lane, c0, acc);
    __syncthreads();
    acc_zero(acc);
    gemm_rows4x8<H>(S.A, LDA, S.w_re1, lane, c0, acc);
    mask_from_global(m_re0, slot0, nrows, lane, warp, bits);
    acc_apply_mask(acc, bits);
    __syncthreads();
    acc_to_smem(S.A, LDA, lane, c0, acc);
    __syncthreads();
    if ((int)threadIdx.x < nrows) {
      const int r = threadIdx.x;
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 4
      for (int k = 0; k < H; k += 4) {
        const float4 a = ld4(S.A + r * LDA + k);
        const float* w = S.w_re0 + k * 8 + 2;      // feature columns 2..4 = s_r - s_s
        o0 = fmaf(a.x, w[0], o0);  o1 = fmaf(a.x, w[1], o1);  o2 = fmaf(a.x, w[2], o2);
        o0 = fmaf(a.y, w[8], o0);  o1 = fmaf(a.y, w[9], o1);  o2 = fmaf(a.y, w[10], o2);
        o0 = fmaf(a.z, w[16], o0); o1 = fmaf(a.z, w[17], o1); o2 = fmaf(a.z, w[18], o2);
        o0 = fmaf(a.w, w[24], o0); o1 = fmaf(a.w, w[25], o1); o2 = fmaf(a.w, w[26], o2);
      }
      st4(gx + (slot0 + r) * 4, make_float4(o0, o1, o2, 0.f));
    }
  }
}

// g_scur[i] = g_pred[i] + sum_{e in row i} g_x(e) - sum_{e: send e = i} g_x(e)
__global__ void k_bwd_positions(const int* __restrict__ rowptr, const int* __restrict__ trowptr,
                                const int* __restrict__ tedge, const float* __restrict__ gx,
                                const float* __restrict__ g_pred, long long g_stride, float* __restrict__ g_s_cur,
                                int B, int N) {
  const long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= (long long)B * N) return;
  const int b = (int)(node / N), i = (int)(node % N);
  const long long slot = (long long)b * KMAX * N;
  const float* gp = g_pred + (long long)b * g_stride + i * 3;
  float x = gp[0], y = gp[1], z = gp[2];
  const int* rp = rowptr + (long long)b * (N + 1) + i;
  for (int e = rp[0]; e < rp[1]; ++e) {
    const float4 g = ld4(gx + (slot + e) * 4);
    x += g.x; y += g.y; z += g.z;
  }
  const int* tp = trowptr + (long long)b * (N + 1) + i;
  for (int k = tp[0]; k < tp[1]; ++k) {
    const float4 g = ld4(gx + (slot + tedge[slot + k]) * 4);
    x -= g.x; y -= g.y; z -= g.z;
  }
  float* o = g_s_cur + node * 3;
  o[0] = x; o[1] = y; o[2] = z;
}

template <typename Kern>
static int set_smem_b(Kern k, size_t bytes) {
  return (int)cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int launch_step_backward(const float* wpack, const Csr& csr, const Masks& mk, const float* g_pred,
                         long long g_stride, float* g_s_cur, float* g_s_delta, void* scratch, int B, int N,
                         cudaStream_t st) {
  static DeviceOnce once;
  const int once_dev = once.pending();
  if (once_dev >= 0) {
    int e;
    if ((e = set_smem_b(k_bwd_head, sizeof(BwdHeadSmem)))) return e;
    if ((e = set_smem_b(k_bwd_prop<false>, sizeof(BwdPropSmem)))) return e;
    if ((e = set_smem_b(k_bwd_prop<true>, sizeof(BwdPropSmem)))) return e;
    if ((e = set_smem_b(k_bwd_edge, sizeof(BwdEdgeSmem)))) return e;
    once.done(once_dev);
  }
  if (!csr.trowptr) return (int)cudaErrorInvalidValue;
  BwdScratch s = carve_bwd(scratch, B, N);
  const long long R = (long long)B * N;
  const int node_tiles = (int)((R + TILE - 1) / TILE);
  const long long edge_tiles = (long long)B * ((KMAX * N + TILE - 1) / TILE);
  const int g2 = node_tiles < 2 * NSM ? node_tiles : 2 * NSM;
  const int g1 = node_tiles < NSM ? node_tiles : NSM;
  const int ge = edge_tiles < 2 * NSM ? (int)edge_tiles : 2 * NSM;

  if (g_use_tensor_cores) {
    const int e = launch_bwd_head_tc(wpack, g_pred, g_stride, mk.q, mk.eff[2], s.gz, s.gcp, s.gagg[2], B, N, st);
    if (e) return e;
  } else {
    k_bwd_head<<<g2, NT, sizeof(BwdHeadSmem), st>>>(wpack, g_pred, g_stride, mk.q, mk.eff[2], s.gz, s.gcp, s.gagg[2],
                                                    B, N);
    PILE_CHECK_LAUNCH();
  }
  const int gg = (int)((R + 15) / 16 < 16 * NSM ? (R + 15) / 16 : 16 * NSM);
  for (int p = PSTEP - 1; p >= 0; --p) {
    k_bwd_gather<<<gg, 256, 0, st>>>(csr.rowptr, csr.trowptr, csr.trecv, csr.tedge, mk.edge[p], s.gagg[p], s.gpr, s.gps,
                                     B, N);
    PILE_CHECK_LAUNCH();
    if (g_use_tensor_cores) {
      const int e = launch_bwd_prop_tc(wpack, p == 0, s.gpr, s.gps, p >= 1 ? mk.eff[p - 1] : mk.pe1, mk.pe0, s.gz,
                                       s.gcp, p >= 1 ? s.gagg[p - 1] : nullptr, g_s_delta, B, N, st);
      if (e) return e;
      continue;
    }
    if (p >= 1)
      k_bwd_prop<false><<<g1, NT, sizeof(BwdPropSmem), st>>>(wpack, s.gpr, s.gps, mk.eff[p - 1], nullptr, s.gz, s.gcp,
                                                             s.gagg[p - 1], nullptr, B, N);
    else
      k_bwd_prop<true><<<g1, NT, sizeof(BwdPropSmem), st>>>(wpack, s.gpr, s.gps, mk.pe1, mk.pe0, s.gz, s.gcp, nullptr,
                                                            g_s_delta, B, N);
    PILE_CHECK_LAUNCH();
  }
  if (g_use_tensor_cores) {
    const int e = launch_bwd_edge_tc(wpack, csr, mk, s.gagg[0], s.gagg[1], s.gagg[2], s.gx, B, N, st);
    if (e) return e;
  } else {
    k_bwd_edge<<<ge, NT, sizeof(BwdEdgeSmem), st>>>(wpack, csr.rowptr, csr.row, mk.edge[0], mk.edge[1], mk.edge[2],
                                                    mk.re0, mk.re1, mk.re2, s.gagg[0], s.gagg[1], s.gagg[2], s.gx, B, N);
    PILE_CHECK_LAUNCH();
  }
  k_bwd_positions<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(csr.rowptr, csr.trowptr, csr.tedge, s.gx, g_pred,
                                                                g_stride, g_s_cur, B, N);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
