// K3 on the 5th-generation tensor cores: relation encoder + hoisted relation-propagator term.
//
// Same math as k_edge_encode (fwd.cu; reference model/gnn_dyn.py:179-180, 187) but the three 64x64
// layers of every 128-relation tile run as tcgen05.mma (bf16 hi/lo split, fp32 accumulation in TMEM):
//
//   layer 0 (K = 6)   CUDA cores, one relation per thread  -> ReLU -> split -> A tile in shared memory
//   layer 1, 2        D = A * W^T on tcgen05 -> tcgen05.ld -> +bias, ReLU, split -> A tile (same smem)
//   layer E (W_e)     D = A * W_e^T          -> tcgen05.ld -> + (w_d d + b) -> C_e rows in HBM
//
// A CTA is persistent (one per SM) and holds 4 independent 128-thread groups; each group owns one
// relation tile at a time (its A tile pair, 64 TMEM columns, one mbarrier) and issues its own MMAs from
// one elected thread, so while one group waits for the tensor pipe the others run their epilogues.
// Activations never leave the SM between layers; the weights (bf16 hi/lo, canonical K-major layout,
// 48 KB) arrive once per CTA with one bulk (TMA) copy.
#include "common.cuh"
#include "kernels.h"
#include "tc.cuh"

namespace pile {

constexpr int TC_GROUPS = 4;
constexpr int TC_THREADS = TC_GROUPS * 128;
constexpr uint32_t A_SBO = 128, A_LBO = (TILE / 8) * 128;   // 2048
constexpr uint32_t B_SBO = 128, B_LBO = (H / 8) * 128;      // 1024
constexpr uint32_t A_BYTES = TILE * H * 2;                  // 16 KB per part
constexpr uint32_t B_BYTES = H * H * 2;                     // 8 KB per part
constexpr uint32_t TMEM_COLS = TC_GROUPS * H;               // 256

struct EdgeTcSmem {
  alignas(128) uint8_t wb[3][2][B_BYTES];            // [layer RE1, RE2, E][hi, lo]
  alignas(128) uint8_t a[TC_GROUPS][2][A_BYTES];     // [group][hi, lo]
  float w_re0[8 * H];
  float b_re0[H], b_re1[H], b_re2[H], wd_rp[H], b_rp[H];
  uint64_t mma_bar[TC_GROUPS];
  uint64_t w_bar;
  uint32_t tmem_base;
};

__device__ __forceinline__ void group_barrier(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(128) : "memory"); }

// 3-pass split product of the group's A tile with weight block `layer`
__device__ __forceinline__ void issue_layer(EdgeTcSmem& S, int g, int layer, uint32_t tmem_d) {
  constexpr uint32_t idesc = tc::make_idesc_bf16(TILE, H);
  const uint32_t a_hi = tc::smem_u32(S.a[g][0]), a_lo = tc::smem_u32(S.a[g][1]);
  const uint32_t b_hi = tc::smem_u32(S.wb[layer][0]), b_lo = tc::smem_u32(S.wb[layer][1]);
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t a = pass == 1 ? a_lo : a_hi;
    const uint32_t b = pass == 2 ? b_lo : b_hi;
#pragma unroll
    for (int k = 0; k < H / 16; ++k) {
      const uint64_t da = tc::make_desc(a + k * 2 * A_LBO, A_LBO, A_SBO);
      const uint64_t db = tc::make_desc(b + k * 2 * B_LBO, B_LBO, B_SBO);
      tc::mma_bf16(tmem_d, da, db, idesc, (pass | k) != 0 ? 1u : 0u);
    }
  }
}

// 8 consecutive activations of row t -> one 16-byte K chunk of the hi and lo A tiles
__device__ __forceinline__ void store_chunk(EdgeTcSmem& S, int g, int t, int kc, const float (&v)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) tc::split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
  const uint32_t off = (t >> 3) * A_SBO + kc * A_LBO + (t & 7) * 16;
  *reinterpret_cast<uint4*>(S.a[g][0] + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(S.a[g][1] + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// epilogue of a hidden layer: TMEM -> +bias -> ReLU (sign bits to the tape) -> split -> A tile
__device__ __forceinline__ void hidden_epilogue(EdgeTcSmem& S, int g, int t, uint32_t taddr, const float* bias,
                                                uint8_t* __restrict__ mask, long long mrow, bool valid) {
  float v[4][16];
#pragma unroll
  for (int q = 0; q < 4; ++q) tc::tmem_ld16(taddr + q * 16, v[q]);
  tc::tmem_ld_wait();
  uint32_t mlo = 0, mhi = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float o[8];
      unsigned m = 0;
      const float4 b0 = ld4(bias + q * 16 + h * 8), b1 = ld4(bias + q * 16 + h * 8 + 4);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x = v[q][h * 8 + j] + bb[j];
        m |= x > 0.f ? (1u << j) : 0u;
        o[j] = fmaxf(x, 0.f);
      }
      const int kc = q * 2 + h;
      if (kc < 4) mlo |= m << (8 * kc); else mhi |= m << (8 * (kc - 4));
      store_chunk(S, g, t, kc, o);
    }
  }
  if (mask != nullptr && valid) *reinterpret_cast<uint2*>(mask + mrow * 8) = make_uint2(mlo, mhi);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_edge_encode_tc(const float* __restrict__ wpack, const float* __restrict__ attr, const float* __restrict__ dens,
                 const float* __restrict__ s_cur, long long s_stride, const int* __restrict__ rowptr,
                 const int* __restrict__ col, const int* __restrict__ row, uint8_t* __restrict__ m_re0,
                 uint8_t* __restrict__ m_re1, uint8_t* __restrict__ m_re2, float* __restrict__ Ce, int B, int N) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  EdgeTcSmem& S = *reinterpret_cast<EdgeTcSmem*>(smem_raw);
  const int g = threadIdx.x >> 7, t = threadIdx.x & 127, wig = t >> 5;

  if (threadIdx.x < 32) tc::tmem_alloc(&S.tmem_base, TMEM_COLS);
  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_GROUPS; ++i) tc::mbar_init(&S.mma_bar[i], 1);
    tc::mbar_init(&S.w_bar, 1);
    tc::mbar_init_fence();
  }
  load_block(S.w_re0, wpack + wslot_offset(W_RE0T), 8 * H);
  load_block(S.b_re0, wpack + wslot_offset(B_RE0), H);
  load_block(S.b_re1, wpack + wslot_offset(B_RE1), H);
  load_block(S.b_re2, wpack + wslot_offset(B_RE2), H);
  load_block(S.wd_rp, wpack + wslot_offset(WD_RP), H);
  load_block(S.b_rp, wpack + wslot_offset(B_RP), H);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(&S.w_bar, 3 * 2 * B_BYTES);
    tc::bulk_g2s(S.wb, wpack + wslot_offset(TC_RE1), 3 * 2 * B_BYTES, &S.w_bar);
  }
  tc::mbar_wait(&S.w_bar, 0);

  const uint32_t tmem_d = S.tmem_base + g * H;
  const uint32_t taddr = tmem_d + ((uint32_t)(wig * 32) << 16);
  uint32_t phase = 0;
  const int tps = (KMAX * N + TILE - 1) / TILE;
  const long long ntiles = (long long)B * tps;
  for (long long tile = (long long)blockIdx.x * TC_GROUPS + g; tile < ntiles; tile += (long long)gridDim.x * TC_GROUPS) {
    const int b = (int)(tile / tps);
    const int e0 = (int)(tile % tps) * TILE;
    const int ne = rowptr[(long long)b * (N + 1) + N];
    if (e0 >= ne) continue;                               // group-uniform
    const int nrows = min(TILE, ne - e0);
    const long long slot0 = (long long)b * KMAX * N + e0;
    const float d = dens[b] / 5000.f;
    const bool valid = t < nrows;

    // ---- layer 0 on CUDA cores: x = [attr_r, attr_s, s_r - s_s, d] -------------------------------------
    float x[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (valid) {
      const int r = row[slot0 + t], c = col[slot0 + t];
      const float* pr = s_cur + (long long)b * s_stride + r * 3;
      const float* ps = s_cur + (long long)b * s_stride + c * 3;
      x[0] = attr[(long long)b * N + r]; x[1] = attr[(long long)b * N + c];
      x[2] = pr[0] - ps[0]; x[3] = pr[1] - ps[1]; x[4] = pr[2] - ps[2]; x[5] = d;
    }
    {
      uint32_t mlo = 0, mhi = 0;
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        float o[8];
        const float4 b0 = ld4(S.b_re0 + kc * 8), b1 = ld4(S.b_re0 + kc * 8 + 4);
        o[0] = b0.x; o[1] = b0.y; o[2] = b0.z; o[3] = b0.w; o[4] = b1.x; o[5] = b1.y; o[6] = b1.z; o[7] = b1.w;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const float4 w0 = ld4(S.w_re0 + k * H + kc * 8), w1 = ld4(S.w_re0 + k * H + kc * 8 + 4);
          o[0] = fmaf(x[k], w0.x, o[0]); o[1] = fmaf(x[k], w0.y, o[1]); o[2] = fmaf(x[k], w0.z, o[2]);
          o[3] = fmaf(x[k], w0.w, o[3]); o[4] = fmaf(x[k], w1.x, o[4]); o[5] = fmaf(x[k], w1.y, o[5]);
          o[6] = fmaf(x[k], w1.z, o[6]); o[7] = fmaf(x[k], w1.w, o[7]);
        }
        unsigned m = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { m |= o[j] > 0.f ? (1u << j) : 0u; o[j] = fmaxf(o[j], 0.f); }
        if (kc < 4) mlo |= m << (8 * kc); else mhi |= m << (8 * (kc - 4));
        store_chunk(S, g, t, kc, o);
      }
      if (m_re0 != nullptr && valid) *reinterpret_cast<uint2*>(m_re0 + (slot0 + t) * 8) = make_uint2(mlo, mhi);
    }

    // ---- layers 1, 2, E on the tensor cores ---------------------------------------------------------------
#pragma unroll 1
    for (int layer = 0; layer < 3; ++layer) {
      tc::fence_async_smem();          // A tile written with st.shared -> visible to the tensor core
      tc::fence_before_sync();         // our tcgen05.ld of the previous accumulator are complete
      group_barrier(g);
      if (t == 0) {
        tc::fence_after_sync();
        issue_layer(S, g, layer, tmem_d);
        tc::mma_commit(&S.mma_bar[g]);
      }
      tc::mbar_wait(&S.mma_bar[g], phase);
      phase ^= 1;
      tc::fence_after_sync();
      if (layer == 0) {
        hidden_epilogue(S, g, t, taddr, S.b_re1, m_re1, slot0 + t, valid);
      } else if (layer == 1) {
        hidden_epilogue(S, g, t, taddr, S.b_re2, m_re2, slot0 + t, valid);
      } else {
        float v[4][16];
#pragma unroll
        for (int q = 0; q < 4; ++q) tc::tmem_ld16(taddr + q * 16, v[q]);
        tc::tmem_ld_wait();
        if (valid) {
          float* out = Ce + (slot0 + t) * H;
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const int c = q * 16 + j;
              const float4 wd = ld4(S.wd_rp + c), bb = ld4(S.b_rp + c);
              st4(out + c, make_float4(v[q][j] + fmaf(wd.x, d, bb.x), v[q][j + 1] + fmaf(wd.y, d, bb.y),
                                       v[q][j + 2] + fmaf(wd.z, d, bb.z), v[q][j + 3] + fmaf(wd.w, d, bb.w)));
            }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(S.tmem_base, TMEM_COLS);
}

int launch_edge_encode_tc(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                          long long s_stride, const Csr& csr, const Masks* mk, float* Ce, int B, int N,
                          cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_encode_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(EdgeTcSmem));
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const long long ntiles = (long long)B * ((KMAX * N + TILE - 1) / TILE);
  const long long want = (ntiles + TC_GROUPS - 1) / TC_GROUPS;
  const int grid = (int)(want < NSM ? want : NSM);
  k_edge_encode_tc<<<grid, TC_THREADS, sizeof(EdgeTcSmem), st>>>(
      wpack, attr, dens, s_cur, s_stride, csr.rowptr, csr.col, csr.row, mk ? mk->re0 : nullptr,
      mk ? mk->re1 : nullptr, mk ? mk->re2 : nullptr, Ce, B, N);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
