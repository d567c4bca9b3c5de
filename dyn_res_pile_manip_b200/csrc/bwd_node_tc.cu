// K8 on the tensor cores, particle side: the GEMMs of one propagation step of the dgrad pass.
//
// Same math as k_bwd_prop (bwd.cu), after k_bwd_gather has produced the receiver- and sender-side sums
//   g_Pr[i] = sum_{e in row i} g_m(e),   g_Ps[i] = sum_{e: send e = i} g_m(e):
//   p >= 1 :  g_eff = g_z + g_Pr W_r + g_Ps W_s ;  g_z <- g_eff * [eff_{p-1} > 0] ;  g_Cp += g_z ;  g_agg_{p-1} = g_z W_a
//   p == 0 :  g_penc = g_z + g_Pr W_r + g_Ps W_s + g_Cp W_p ;
//             g_sdelta = (((g_penc * [penc > 0]) PE1) * [h > 0]) PE0[:, :3]
// as tcgen05 GEMMs per 128-particle tile (bf16 hi/lo split, three passes, fp32 accumulation in TMEM; machinery of
// tc_tile.cuh).  dgrad needs D[row][k] = sum_n A[row][n] W[n][k], so every B operand is the image of W^T; ReLUs are
// replayed from the recorded sign bits (one uint32 per row and 32-column half).
#include "kernels.h"
#include "tc_tile.cuh"

namespace pile {

constexpr uint32_t BN_W64 = 2 * b_bytes(64, 64);      // 16 KB  [hi | lo] of a transposed 64x64 weight
constexpr uint32_t BN_W16 = 2 * b_bytes(16, 64);      //  4 KB  rows 0..2 = PE0[:, j] (the three s_delta inputs)
// slot order: W_a^T | W_r^T | W_s^T | W_p^T | PE1^T | PE0sel | V0^T | V1sel ; p >= 1 loads the first three, p == 0
// W_r^T .. PE0sel, the head kernel W_a^T and the last two
constexpr uint32_t BN_W64x16 = 2 * b_bytes(64, 16);   //  4 KB  V1^T padded to [64 x 16] (the three outputs)
constexpr uint32_t BN_OFF_A = 0, BN_OFF_R = BN_W64, BN_OFF_V0 = 5 * BN_W64 + BN_W16,
                   TC_BWD_NODE_BYTES = 6 * BN_W64 + BN_W16 + BN_W64x16;
static_assert(TC_BWD_NODE_BYTES == 4 * TC_BWD_NODE_FLOATS, "TC_BWD_NODE slot size (common.cuh) out of sync");

template <bool FIRST>
struct BwdNodeTcSmem {
  alignas(128) uint8_t w[FIRST ? 4 * BN_W64 + BN_W16 : 3 * BN_W64];
  GroupTile t[TC_GROUPS];
  uint64_t bar[TC_GROUPS];
  uint64_t w_bar;
  uint32_t tmem_base;
};
static_assert(sizeof(BwdNodeTcSmem<true>) <= 227 * 1024, "shared memory budget");

// D (=|+=) A * W with the group's A tile (three passes, four K-steps); first_acc: accumulate onto D from the start
template <int N>
__device__ __forceinline__ void issue_dgrad(uint32_t elected, uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b,
                                            uint32_t b_part_bytes, uint32_t first_acc) {
  constexpr uint32_t idesc = tc::make_idesc_bf16(TILE, N);
  constexpr uint32_t BL = b_lbo(N);
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t a = pass == 1 ? a_lo : a_hi;
    const uint32_t bb = pass == 2 ? b + b_part_bytes : b;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      tc::mma_bf16_if(elected, tmem_d, tc::make_desc(a + k * 2 * A_LBO, A_LBO, A_SBO),
                      tc::make_desc(bb + k * 2 * BL, BL, B_SBO), idesc, (pass | k) != 0 ? 1u : first_acc);
  }
}

// 16 accumulator values, zeroed where the recorded sign bit is clear, -> two K chunks of the A tile (and back in v)
__device__ __forceinline__ void mask16_to_tile(uint8_t* a_hi, uint8_t* a_lo, uint32_t off0, float (&v)[16], uint32_t bits16) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] = (bits16 >> (h * 8 + j)) & 1u ? v[h * 8 + j] : 0.f;
      v[h * 8 + j] = o[j];
    }
    store_chunk(a_hi, a_lo, off0 + h * A_LBO, o);
  }
}

template <bool FIRST>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_bwd_prop_tc(const float* __restrict__ wpack, const float* __restrict__ gpr, const float* __restrict__ gps,
              const uint8_t* __restrict__ m_next /* eff[p-1] or pe1 */, const uint8_t* __restrict__ m_pe0,
              float* __restrict__ gz, float* __restrict__ gcp, float* __restrict__ gagg_out,
              float* __restrict__ g_s_delta, int B, int N) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdNodeTcSmem<FIRST>& S = *reinterpret_cast<BwdNodeTcSmem<FIRST>*>(smem_raw);
  const int g = threadIdx.x / GROUP_THREADS, t = threadIdx.x % GROUP_THREADS;
  const int wig = t >> 5;
  const int r = (wig & 3) * 32 + (t & 31), half = wig >> 2;

  if (threadIdx.x < 32) tc::tmem_alloc(&S.tmem_base, TC_GROUPS * H);
  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_GROUPS; ++i) tc::mbar_init(&S.bar[i], 1);
    tc::mbar_init(&S.w_bar, 1);
    tc::mbar_init_fence();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(&S.w_bar, sizeof(S.w));
    tc::bulk_g2s(S.w, reinterpret_cast<const uint8_t*>(wpack + wslot_offset(TC_BWD_NODE)) + (FIRST ? BN_W64 : 0),
                 sizeof(S.w), &S.w_bar);
  }
  GroupCtx c;
  c.g = g; c.wig = wig;
  c.tmem_d = S.tmem_base + g * H;
  c.taddr = c.tmem_d + ((uint32_t)((wig & 3) * 32) << 16);
  c.a_hi = tc::smem_u32(S.t[g].a[0]); c.a_lo = tc::smem_u32(S.t[g].a[1]);
  c.aux_hi = c.aux_lo = 0; c.zero = 0;
  c.bar = &S.bar[g]; c.phase = 0;
  uint8_t* const a_hi = S.t[g].a[0];
  uint8_t* const a_lo = S.t[g].a[1];
  const uint32_t row_off = (r >> 3) * A_SBO + (r & 7) * 16;
  const uint32_t w = tc::smem_u32(S.w);
  // offsets inside the loaded part of the slot
  const uint32_t w_a = w + BN_OFF_A, w_r = w + (FIRST ? 0 : BN_OFF_R), w_s = w_r + BN_W64;
  const uint32_t w_p = w + 2 * BN_W64, w_pe1 = w + 3 * BN_W64, w_pe0 = w + 4 * BN_W64;      // FIRST only
  const long long R = (long long)B * N;
  const int ntiles = (int)((R + TILE - 1) / TILE);
  tc::mbar_wait(&S.w_bar, 0);

  // group g of CTA c takes tiles g * gridDim + c, + 4 * gridDim, ...: a small workload spreads over all SMs (one
  // tile chain per SM) before any SM runs four chains side by side
  for (int tile = g * (int)gridDim.x + (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x * TC_GROUPS) {
    const long long row0 = (long long)tile * TILE;
    const long long row = row0 + r;
    const bool valid = row < R;
    // g_eff = g_Pr W_r + g_Ps W_s (+ g_Cp W_p) accumulated in TMEM, the residual g_z is added in the epilogue
    load_rows_to_tile(gpr, row0, R, t, a_hi, a_lo);
    run_gemm(c, [&](uint32_t el) { issue_dgrad<64>(el, c.tmem_d, c.a_hi, c.a_lo, w_r, BN_W64 / 2, 0u); });
    load_rows_to_tile(gps, row0, R, t, a_hi, a_lo);
    run_gemm(c, [&](uint32_t el) { issue_dgrad<64>(el, c.tmem_d, c.a_hi, c.a_lo, w_s, BN_W64 / 2, 1u); });
    if (FIRST) {
      load_rows_to_tile(gcp, row0, R, t, a_hi, a_lo);
      run_gemm(c, [&](uint32_t el) { issue_dgrad<64>(el, c.tmem_d, c.a_hi, c.a_lo, w_p, BN_W64 / 2, 1u); });
    }
    const uint32_t bits = valid ? *reinterpret_cast<const uint32_t*>(m_next + row * 8 + half * 4) : 0u;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float v[16], z[16];
      tc::tmem_ld16(c.taddr + half * 32 + q * 16, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] = 0.f;
      if (valid) { ld8(gz + row * H + half * 32 + q * 16, z); ld8(gz + row * H + half * 32 + q * 16 + 8, z + 8); }
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += z[j];
      mask16_to_tile(a_hi, a_lo, row_off + (half * 4 + q * 2) * A_LBO, v, bits >> (q * 16));
      if (!FIRST && valid) {          // g_z of the step below; g_Cp accumulates it
        float* zp = gz + row * H + half * 32 + q * 16;
        float* cp = gcp + row * H + half * 32 + q * 16;
        float a[16];
        ld8(cp, a); ld8(cp + 8, a + 8);
        st8(zp, v); st8(zp + 8, v + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] += v[j];
        st8(cp, a); st8(cp + 8, a + 8);
      }
    }
    if (!FIRST) {
      run_gemm(c, [&](uint32_t el) { issue_dgrad<64>(el, c.tmem_d, c.a_hi, c.a_lo, w_a, BN_W64 / 2, 0u); });
      // g_agg of the step below, row-major (k_bwd_gather reads whole rows): through the staging tile
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float v[16];
        tc::tmem_ld16(c.taddr + half * 32 + q * 16, v);
        tc::tmem_ld_wait();
        stage_put16(a_hi, r, half * 32 + q * 16, v);
      }
      group_barrier(g);
      stage_flush(a_hi, t, gagg_out, row0, R);
      group_barrier(g);
    } else {
      run_gemm(c, [&](uint32_t el) { issue_dgrad<64>(el, c.tmem_d, c.a_hi, c.a_lo, w_pe1, BN_W64 / 2, 0u); });
      const uint32_t b0 = valid ? *reinterpret_cast<const uint32_t*>(m_pe0 + row * 8 + half * 4) : 0u;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float v[16];
        tc::tmem_ld16(c.taddr + half * 32 + q * 16, v);
        tc::tmem_ld_wait();
        mask16_to_tile(a_hi, a_lo, row_off + (half * 4 + q * 2) * A_LBO, v, b0 >> (q * 16));
      }
      run_gemm(c, [&](uint32_t el) { issue_dgrad<16>(el, c.tmem_d, c.a_hi, c.a_lo, w_pe0, BN_W16 / 2, 0u); });
      if (half == 0) {
        float v[16];
        tc::tmem_ld16(c.taddr, v);
        tc::tmem_ld_wait();
        if (valid) {
          float* o = g_s_delta + row * 3;
          o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(S.tmem_base, TC_GROUPS * H);
}

// ------------------------------------------------------------------------------------------------
// head of the backward pass (k_bwd_head's math): g_q = (g V1) * [q > 0];  g_z = (g_q V0) * [eff_3 > 0];
// g_Cp = g_z;  g_agg_2 = g_z W_a
// ------------------------------------------------------------------------------------------------
struct BwdHeadTcSmem {
  alignas(128) uint8_t w[2 * BN_W64 + BN_W64x16];     // W_a^T | V0^T | V1sel
  GroupTile t[TC_GROUPS];
  alignas(128) uint8_t zero[A_LBO];
  uint64_t bar[TC_GROUPS];
  uint64_t w_bar;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
k_bwd_head_tc(const float* __restrict__ wpack, const float* __restrict__ g_pred, long long g_stride,
              const uint8_t* __restrict__ m_q, const uint8_t* __restrict__ m_eff2, float* __restrict__ gz,
              float* __restrict__ gcp, float* __restrict__ gagg2, int B, int N) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdHeadTcSmem& S = *reinterpret_cast<BwdHeadTcSmem*>(smem_raw);
  const int g = threadIdx.x / GROUP_THREADS, t = threadIdx.x % GROUP_THREADS;
  const int wig = t >> 5;
  const int r = (wig & 3) * 32 + (t & 31), half = wig >> 2;

  if (threadIdx.x < 32) tc::tmem_alloc(&S.tmem_base, TC_GROUPS * H);
  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_GROUPS; ++i) tc::mbar_init(&S.bar[i], 1);
    tc::mbar_init(&S.w_bar, 1);
    tc::mbar_init_fence();
  }
  for (int i = threadIdx.x * 16; i < (int)A_LBO; i += TC_THREADS * 16) *reinterpret_cast<uint4*>(S.zero + i) = make_uint4(0, 0, 0, 0);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (threadIdx.x == 0) {
    const uint8_t* src = reinterpret_cast<const uint8_t*>(wpack + wslot_offset(TC_BWD_NODE));
    tc::mbar_expect_tx(&S.w_bar, sizeof(S.w));
    tc::bulk_g2s(S.w, src + BN_OFF_A, BN_W64, &S.w_bar);
    tc::bulk_g2s(S.w + BN_W64, src + BN_OFF_V0, BN_W64 + BN_W64x16, &S.w_bar);
  }
  GroupCtx c;
  c.g = g; c.wig = wig;
  c.tmem_d = S.tmem_base + g * H;
  c.taddr = c.tmem_d + ((uint32_t)((wig & 3) * 32) << 16);
  c.a_hi = tc::smem_u32(S.t[g].a[0]); c.a_lo = tc::smem_u32(S.t[g].a[1]);
  c.aux_hi = c.aux_lo = 0; c.zero = tc::smem_u32(S.zero);
  c.bar = &S.bar[g]; c.phase = 0;
  uint8_t* const a_hi = S.t[g].a[0];
  uint8_t* const a_lo = S.t[g].a[1];
  const uint32_t row_off = (r >> 3) * A_SBO + (r & 7) * 16;
  const uint32_t w = tc::smem_u32(S.w);
  const uint32_t w_a = w, w_v0 = w + BN_W64, w_v1 = w + 2 * BN_W64;
  const long long R = (long long)B * N;
  const int ntiles = (int)((R + TILE - 1) / TILE);
  tc::mbar_wait(&S.w_bar, 0);

  // group g of CTA c takes tiles g * gridDim + c, + 4 * gridDim, ...: a small workload spreads over all SMs (one
  // tile chain per SM) before any SM runs four chains side by side
  for (int tile = g * (int)gridDim.x + (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x * TC_GROUPS) {
    const long long row0 = (long long)tile * TILE;
    const long long row = row0 + r;
    const bool valid = row < R;
    if (half == 0) {          // K chunk 0 = (g_x, g_y, g_z, 0, ...)
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid) {
        const float* p = g_pred + (row / N) * g_stride + (row % N) * 3;
        f[0] = p[0]; f[1] = p[1]; f[2] = p[2];
      }
      store_chunk(a_hi, a_lo, row_off, f);
    }
    run_gemm(c, [&](uint32_t el) { issue_gemm_k16<64>(el, c.tmem_d, c.a_hi, c.a_lo, c.zero, w_v1, w_v1 + BN_W64x16 / 2); });
    const uint32_t bq = valid ? *reinterpret_cast<const uint32_t*>(m_q + row * 8 + half * 4) : 0u;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float v[16];
      tc::tmem_ld16(c.taddr + half * 32 + q * 16, v);
      tc::tmem_ld_wait();
      mask16_to_tile(a_hi, a_lo, row_off + (half * 4 + q * 2) * A_LBO, v, bq >> (q * 16));
    }
    run_gemm(c, [&](uint32_t el) { issue_dgrad<64>(el, c.tmem_d, c.a_hi, c.a_lo, w_v0, BN_W64 / 2, 0u); });
    const uint32_t be = valid ? *reinterpret_cast<const uint32_t*>(m_eff2 + row * 8 + half * 4) : 0u;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float v[16];
      tc::tmem_ld16(c.taddr + half * 32 + q * 16, v);
      tc::tmem_ld_wait();
      mask16_to_tile(a_hi, a_lo, row_off + (half * 4 + q * 2) * A_LBO, v, be >> (q * 16));
      if (valid) {
        float* zp = gz + row * H + half * 32 + q * 16;
        float* cp = gcp + row * H + half * 32 + q * 16;
        st8(zp, v); st8(zp + 8, v + 8);
        st8(cp, v); st8(cp + 8, v + 8);
      }
    }
    run_gemm(c, [&](uint32_t el) { issue_dgrad<64>(el, c.tmem_d, c.a_hi, c.a_lo, w_a, BN_W64 / 2, 0u); });
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float v[16];
      tc::tmem_ld16(c.taddr + half * 32 + q * 16, v);
      tc::tmem_ld_wait();
      stage_put16(a_hi, r, half * 32 + q * 16, v);
    }
    group_barrier(g);
    stage_flush(a_hi, t, gagg2, row0, R);
    group_barrier(g);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(S.tmem_base, TC_GROUPS * H);
}

int launch_bwd_head_tc(const float* wpack, const float* g_pred, long long g_stride, const uint8_t* m_q,
                       const uint8_t* m_eff2, float* gz, float* gcp, float* gagg2, int B, int N, cudaStream_t st) {
  static DeviceOnce once;
  const int once_dev = once.pending();
  if (once_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(k_bwd_head_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdHeadTcSmem));
    if (e != cudaSuccess) return (int)e;
    once.done(once_dev);
  }
  const long long ntiles = ((long long)B * N + TILE - 1) / TILE;
  const int grid = (int)(ntiles < 1 ? 1 : (ntiles < NSM ? ntiles : NSM));
  k_bwd_head_tc<<<grid, TC_THREADS, sizeof(BwdHeadTcSmem), st>>>(wpack, g_pred, g_stride, m_q, m_eff2, gz, gcp, gagg2, B, N);
  PILE_CHECK_LAUNCH();
  return 0;
}

int launch_bwd_prop_tc(const float* wpack, bool first, const float* gpr, const float* gps, const uint8_t* m_next,
                       const uint8_t* m_pe0, float* gz, float* gcp, float* gagg_out, float* g_s_delta, int B, int N,
                       cudaStream_t st) {
  static DeviceOnce once;
  const int once_dev = once.pending();
  if (once_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(k_bwd_prop_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(BwdNodeTcSmem<false>));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_bwd_prop_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(BwdNodeTcSmem<true>));
    if (e != cudaSuccess) return (int)e;
    once.done(once_dev);
  }
  const long long ntiles = ((long long)B * N + TILE - 1) / TILE;
  const int grid = (int)(ntiles < 1 ? 1 : (ntiles < NSM ? ntiles : NSM));
  if (first)
    k_bwd_prop_tc<true><<<grid, TC_THREADS, sizeof(BwdNodeTcSmem<true>), st>>>(wpack, gpr, gps, m_next, m_pe0, gz, gcp,
                                                                              gagg_out, g_s_delta, B, N);
  else
    k_bwd_prop_tc<false><<<grid, TC_THREADS, sizeof(BwdNodeTcSmem<false>), st>>>(wpack, gpr, gps, m_next, m_pe0, gz, gcp,
                                                                                gagg_out, g_s_delta, B, N);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
