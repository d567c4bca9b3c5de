// K8 on the tensor cores: relation-encoder backward (the largest GEMM block of the dgrad pass).
//
// Same math as k_bwd_edge (bwd.cu):
//   g_Ce(e) = sum_p g_agg_p[recv e] * [m_p(e) > 0]
//   g_x(e)  = ((((g_Ce W_e) * [r3>0]) RE2) * [r2>0]) RE1) * [r1>0]) RE0[:, 2:5]
// as four tcgen05 GEMMs per 128-relation tile (bf16 hi/lo split, fp32 accumulation in TMEM, machinery of
// tc_tile.cuh).  dgrad needs D[row][k] = sum_n A[row][n] W[n][k], i.e. the B operand is W^T = the forward's
// [in][out] layout; ReLUs are replayed from the sign bits the forward recorded (one uint32 per row and
// 32-column half).
#include "kernels.h"
#include "tc_tile.cuh"

namespace pile {

constexpr uint32_t BE_W64 = 2 * b_bytes(64, 64);     // 16 KB  [hi | lo] of a transposed 64x64 weight
constexpr uint32_t BE_W16 = 2 * b_bytes(16, 64);     //  4 KB  rows 0..2 = RE0[:, 2+j] (position columns)
constexpr uint32_t TC_BWD_EDGE_BYTES = 3 * BE_W64 + BE_W16;
static_assert(TC_BWD_EDGE_BYTES == 4 * TC_BWD_EDGE_FLOATS, "TC_BWD_EDGE slot size (common.cuh) out of sync");

struct BwdEdgeTcSmem {
  alignas(128) uint8_t w[TC_BWD_EDGE_BYTES];     // W_e^T | RE2^T | RE1^T | RE0sel
  GroupTile t[TC_GROUPS];
  alignas(128) uint8_t zero[A_LBO];
  uint64_t bar[TC_GROUPS];
  uint64_t w_bar;
  uint32_t tmem_base;
};

__device__ __forceinline__ void masked_to_tile(uint8_t* a_hi, uint8_t* a_lo, uint32_t off0, const float (&v)[16],
                                               uint32_t bits16) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (bits16 >> (h * 8 + j)) & 1u ? v[h * 8 + j] : 0.f;
    store_chunk(a_hi, a_lo, off0 + h * A_LBO, o);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_bwd_edge_tc(const float* __restrict__ wpack, const int* __restrict__ rowptr, const int* __restrict__ row,
              const uint8_t* __restrict__ me0, const uint8_t* __restrict__ me1, const uint8_t* __restrict__ me2,
              const uint8_t* __restrict__ m_re0, const uint8_t* __restrict__ m_re1, const uint8_t* __restrict__ m_re2,
              const float* __restrict__ ga0, const float* __restrict__ ga1, const float* __restrict__ ga2,
              float* __restrict__ gx, int B, int N) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdEdgeTcSmem& S = *reinterpret_cast<BwdEdgeTcSmem*>(smem_raw);
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
    tc::mbar_expect_tx(&S.w_bar, TC_BWD_EDGE_BYTES);
    tc::bulk_g2s(S.w, wpack + wslot_offset(TC_BWD_EDGE), TC_BWD_EDGE_BYTES, &S.w_bar);
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
  const int tps = (KMAX * N + TILE - 1) / TILE;
  const int ntiles = B * tps;
  tc::mbar_wait(&S.w_bar, 0);

  // The chain  row index -> rows of g_agg  is the kernel's longest wait (the ncu source page shows half of all stall
  // samples on it), so the receiver index and the three relation-mask words of a group's NEXT tile are fetched while
  // its current tile runs through the layers.
  const uint8_t* mp[3] = {me0, me1, me2};
  struct Pre { int node; uint32_t bits[3]; int ne; };
  auto prefetch = [&](int tile) {
    Pre p;
    p.node = 0; p.bits[0] = p.bits[1] = p.bits[2] = 0u; p.ne = 0;
    if (tile < ntiles) {
      const int b = tile / tps;
      const int e0 = (tile - b * tps) * TILE;
      p.ne = rowptr[(long long)b * (N + 1) + N];
      if (e0 + r < p.ne) {
        const long long slot = (long long)b * KMAX * N + e0 + r;
        p.node = b * N + row[slot];
#pragma unroll
        for (int q = 0; q < 3; ++q) p.bits[q] = *reinterpret_cast<const uint32_t*>(mp[q] + slot * 8 + half * 4);
      }
    }
    return p;
  };
  const int tstride = (int)gridDim.x * TC_GROUPS;
  Pre cur = prefetch((int)blockIdx.x * TC_GROUPS + g);
  for (int tile = (int)blockIdx.x * TC_GROUPS + g; tile < ntiles; tile += tstride) {
    const int b = tile / tps;
    const int e0 = (tile - b * tps) * TILE;
    const int ne = cur.ne;
    if (e0 >= ne) { cur = prefetch(tile + tstride); continue; }          // group-uniform
    const int nrows = min(TILE, ne - e0);
    const long long slot0 = (long long)b * KMAX * N + e0;
    const bool valid = r < nrows;
    const long long mrow = (slot0 + r) * 8 + half * 4;     // this thread's uint32 of sign bits in every edge mask

    // A = g_Ce: masked sum of the three propagation steps' receiver gradients (32 of the 64 channels)
    {
      float acc[4][8];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
      if (valid) {
        const long long node = cur.node;
        const float* gp[3] = {ga0, ga1, ga2};
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const uint32_t bits = cur.bits[p];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v[8];
            ld8(gp[p] + node * H + half * 32 + j * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[j][i] += (bits >> (j * 8 + i)) & 1u ? v[i] : 0.f;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) store_chunk(a_hi, a_lo, row_off + (half * 4 + j) * A_LBO, acc[j]);
    }
    cur = prefetch(tile + tstride);
    // three 64x64 dgrad layers, each followed by the recorded ReLU mask of the layer below
#pragma unroll 1
    for (int layer = 0; layer < 3; ++layer) {
      const uint32_t wl = w + layer * BE_W64;
      // the recorded ReLU bits of the layer below are fetched before the GEMM is handed over: the load overlaps the MMAs
      const uint8_t* mk = layer == 0 ? m_re2 : (layer == 1 ? m_re1 : m_re0);
      const uint32_t bits = valid ? *reinterpret_cast<const uint32_t*>(mk + mrow) : 0u;
      run_gemm(c, [&](uint32_t el) { issue_gemm<64, 4, false>(el, c.tmem_d, c.a_hi, c.a_lo, 0, 0, 0, wl, wl + BE_W64 / 2); });
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float v[16];
        tc::tmem_ld16(c.taddr + half * 32 + q * 16, v);
        tc::tmem_ld_wait();
        masked_to_tile(a_hi, a_lo, row_off + (half * 4 + q * 2) * A_LBO, v, bits >> (q * 16));
      }
    }
    // g_x = A RE0[:, 2:5]  (N = 16 product, columns 0..2)
    run_gemm(c, [&](uint32_t el) {
      issue_gemm<16, 4, false>(el, c.tmem_d, c.a_hi, c.a_lo, 0, 0, 0, w + 3 * BE_W64, w + 3 * BE_W64 + BE_W16 / 2);
    });
    if (half == 0) {
      float v[16];
      tc::tmem_ld16(c.taddr, v);
      tc::tmem_ld_wait();
      if (valid) st4(gx + (slot0 + r) * 4, make_float4(v[0], v[1], v[2], 0.f));
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(S.tmem_base, TC_GROUPS * H);
}

int launch_bwd_edge_tc(const float* wpack, const Csr& csr, const Masks& mk, const float* ga0, const float* ga1,
                       const float* ga2, float* gx, int B, int N, cudaStream_t st) {
  static DeviceOnce once;
  const int once_dev = once.pending();
  if (once_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(k_bwd_edge_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(BwdEdgeTcSmem));
    if (e != cudaSuccess) return (int)e;
    once.done(once_dev);
  }
  const long long ntiles = (long long)B * ((KMAX * N + TILE - 1) / TILE);
  const long long want = (ntiles + TC_GROUPS - 1) / TC_GROUPS;
  const int grid = (int)(want < NSM ? want : NSM);
  k_bwd_edge_tc<<<grid, TC_THREADS, sizeof(BwdEdgeTcSmem), st>>>(wpack, csr.rowptr, csr.row, mk.edge[0], mk.edge[1],
                                                                  mk.edge[2], mk.re0, mk.re1, mk.re2, ga0, ga1, ga2, gx,
                                                                  B, N);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
