// K3 on the 5th-generation tensor cores: relation encoder + hoisted relation-propagator term.
//
// Same math as k_edge_encode (fwd.cu; reference model/gnn_dyn.py:179-180, 187); ALL four layers of every
// 128-relation tile run as tcgen05.mma (bf16 hi/lo split operands, three passes, fp32 accumulation in
// TMEM), and every bias is folded into the GEMM through one extra K chunk, so the CUDA cores only do
// ReLU + hi/lo split between layers and a plain TMEM -> HBM copy at the end:
//
//   A0 = [attr_r, attr_s, s_r - s_s, d, 1, 0..]        (K = 16)   x W0aug^T  -> ReLU -> split -> A
//   A  = [h (64) | 1, d, 0.. (aux chunk) | 0 (zero chunk)] (K = 80) x W1aug^T -> ReLU -> split -> A
//                                                               x W2aug^T -> ReLU -> split -> A
//                                                               x WEaug^T -> C_e rows (fp32) in HBM
//   W?aug = [W | bias | (w_d for the last layer) | 0..]: the aux chunk of A multiplies the bias / density
//   columns.  The aux and zero chunks are reached through the descriptor's leading-byte offset, so they
//   cost no per-layer work.
//
// A CTA is persistent (one per SM) and holds 4 independent 128-thread groups; each group owns one
// relation tile at a time (its A tile pair, 64 TMEM columns, one mbarrier) and issues its own MMAs from
// one elected thread, so while one group waits for the tensor pipe the others run their epilogues.
// The next tile's relation features are prefetched into registers while the current tile is in flight.
// The 64 KB of weights arrive once per CTA with one bulk (TMA) copy.
#include "common.cuh"
#include "kernels.h"
#include "tc_tile.cuh"

namespace pile {

constexpr uint32_t B_LBO = b_lbo(H);                        // 1024: one K chunk of a 64-row weight
constexpr uint32_t W0_BYTES = 2 * B_LBO;                    // K = 16
constexpr uint32_t WL_BYTES = 10 * B_LBO;                   // K = 80
constexpr uint32_t TMEM_COLS = TC_GROUPS * H;               // 256
constexpr uint32_t TC_EDGE_BYTES = 2 * W0_BYTES + 3 * 2 * WL_BYTES;   // 64 KB

struct EdgeTcSmem {
  alignas(128) uint8_t w0[2][W0_BYTES];              // [hi, lo]     } one contiguous 64 KB image of the
  alignas(128) uint8_t wl[3][2][WL_BYTES];           // [RE1,RE2,E]  } TC_EDGE weight slot
  alignas(128) uint8_t a[TC_GROUPS][2][A_BYTES];     // [group][hi, lo] activation tile, 8 K chunks
  alignas(128) uint8_t aux[TC_GROUPS][2][A_LBO];     // [group][hi, lo] K chunk (1, d, 0, ...)
  alignas(128) uint8_t zero[A_LBO];                  // all-zero K chunk
  uint64_t mma_bar[TC_GROUPS];
  uint64_t w_bar;
  uint32_t tmem_base;
};
static_assert(sizeof(EdgeTcSmem) <= 227 * 1024, "shared memory budget");
static_assert(TC_EDGE_BYTES == H * H * 4 * 4, "TC_EDGE slot size (common.cuh) out of sync");

// One split product D = A * B^T: KSTEPS data K-steps + the aux K-step (first chunk = (1, d, 0..), second
// chunk = the shared zero chunk).  Executed by a CONVERGED warp: descriptors are warp-uniform, only the
// tcgen05.mma itself is predicated on the elected lane.
template <int KSTEPS>
__device__ __forceinline__ void issue_layer(uint32_t elected, uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo,
                                            uint32_t aux_hi, uint32_t aux_lo, uint32_t zero, uint32_t b_hi,
                                            uint32_t b_lo) {
  constexpr uint32_t idesc = tc::make_idesc_bf16(TILE, H);
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t a = pass == 1 ? a_lo : a_hi;
    const uint32_t x = pass == 1 ? aux_lo : aux_hi;
    const uint32_t b = pass == 2 ? b_lo : b_hi;
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k)
      tc::mma_bf16_if(elected, tmem_d, tc::make_desc(a + k * 2 * A_LBO, A_LBO, A_SBO),
                      tc::make_desc(b + k * 2 * B_LBO, B_LBO, B_SBO), idesc, (pass | k) != 0 ? 1u : 0u);
    tc::mma_bf16_if(elected, tmem_d, tc::make_desc(x, zero - x, A_SBO),
                    tc::make_desc(b + KSTEPS * 2 * B_LBO, B_LBO, B_SBO), idesc, 1u);
  }
}

// layer 0: a single K-step whose first chunk holds the 6 features + the constant 1
__device__ __forceinline__ void issue_layer0(uint32_t elected, uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo,
                                             uint32_t zero, uint32_t b_hi, uint32_t b_lo) {
  constexpr uint32_t idesc = tc::make_idesc_bf16(TILE, H);
  tc::mma_bf16_if(elected, tmem_d, tc::make_desc(a_hi, zero - a_hi, A_SBO), tc::make_desc(b_hi, B_LBO, B_SBO), idesc, 0u);
  tc::mma_bf16_if(elected, tmem_d, tc::make_desc(a_lo, zero - a_lo, A_SBO), tc::make_desc(b_hi, B_LBO, B_SBO), idesc, 1u);
  tc::mma_bf16_if(elected, tmem_d, tc::make_desc(a_hi, zero - a_hi, A_SBO), tc::make_desc(b_lo, B_LBO, B_SBO), idesc, 1u);
}

// epilogue of a hidden layer for this thread's (row, 32-column half): TMEM -> ReLU (sign bits to the tape)
// -> split -> A tile
template <bool RECORD>
__device__ __forceinline__ void hidden_epilogue(uint8_t* a_hi, uint8_t* a_lo, uint32_t row_off, int half,
                                                uint32_t taddr, uint8_t* __restrict__ mask, long long mrow,
                                                bool valid) {
  float v[2][16];
#pragma unroll
  for (int q = 0; q < 2; ++q) tc::tmem_ld16(taddr + half * 32 + q * 16, v[q]);
  tc::tmem_ld_wait();
  uint32_t mbits = 0;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float o[8];
      unsigned m = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x = v[q][h * 8 + j];
        if (RECORD) m |= x > 0.f ? (1u << j) : 0u;
        o[j] = fmaxf(x, 0.f);
      }
      if (RECORD) mbits |= m << (8 * (q * 2 + h));
      store_chunk(a_hi, a_lo, row_off + (half * 4 + q * 2 + h) * A_LBO, o);
    }
  }
  if (RECORD && valid) *reinterpret_cast<uint32_t*>(mask + mrow * 8 + half * 4) = mbits;
}

// relation features [attr_r, attr_s, s_r - s_s (3), d, 0, 0] per relation slot, written once per step so
// that the tensor-core kernel's per-tile loads have no dependent index chain
__global__ void k_edge_features(const float* __restrict__ attr, const float* __restrict__ dens,
                                const float* __restrict__ s_cur, long long s_stride, const int* __restrict__ rowptr,
                                const int* __restrict__ col, const int* __restrict__ row, float* __restrict__ efeat,
                                int B, int N) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int ne = rowptr[(long long)b * (N + 1) + N];
  if (e >= ne) return;
  const long long slot = (long long)b * KMAX * N + e;
  const int r = row[slot], c = col[slot];
  const float* pr = s_cur + (long long)b * s_stride + r * 3;
  const float* ps = s_cur + (long long)b * s_stride + c * 3;
  st4(efeat + slot * 8, make_float4(attr[(long long)b * N + r], attr[(long long)b * N + c], pr[0] - ps[0], pr[1] - ps[1]));
  st4(efeat + slot * 8 + 4, make_float4(pr[2] - ps[2], dens[b] / 5000.f, 0.f, 0.f));
}

template <bool RECORD>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_edge_encode_tc(const float* __restrict__ wpack, const float* __restrict__ efeat, const int* __restrict__ rowptr,
                 uint8_t* __restrict__ m_re0, uint8_t* __restrict__ m_re1, uint8_t* __restrict__ m_re2,
                 float* __restrict__ Ce, int B, int N) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  EdgeTcSmem& S = *reinterpret_cast<EdgeTcSmem*>(smem_raw);
  const int g = threadIdx.x / GROUP_THREADS, t = threadIdx.x % GROUP_THREADS;
  const int wig = t >> 5;                  // warp in group, 0..7
  const int r = (wig & 3) * 32 + (t & 31); // tile row = TMEM lane
  const int half = wig >> 2;               // which 32 of the 64 columns this thread handles

  if (threadIdx.x < 32) tc::tmem_alloc(&S.tmem_base, TMEM_COLS);
  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_GROUPS; ++i) tc::mbar_init(&S.mma_bar[i], 1);
    tc::mbar_init(&S.w_bar, 1);
    tc::mbar_init_fence();
  }
  for (int i = threadIdx.x * 16; i < (int)A_LBO; i += TC_THREADS * 16) *reinterpret_cast<uint4*>(S.zero + i) = make_uint4(0, 0, 0, 0);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(&S.w_bar, TC_EDGE_BYTES);
    tc::bulk_g2s(S.w0, wpack + wslot_offset(TC_EDGE), TC_EDGE_BYTES, &S.w_bar);
  }

  const uint32_t tmem_d = S.tmem_base + g * H;
  const uint32_t taddr = tmem_d + ((uint32_t)((wig & 3) * 32) << 16);
  uint8_t* const a_hi = S.a[g][0];
  uint8_t* const a_lo = S.a[g][1];
  const uint32_t a_hi_u = tc::smem_u32(a_hi), a_lo_u = tc::smem_u32(a_lo);
  const uint32_t aux_hi_u = tc::smem_u32(S.aux[g][0]), aux_lo_u = tc::smem_u32(S.aux[g][1]);
  const uint32_t zero_u = tc::smem_u32(S.zero);
  const uint32_t row_off = (r >> 3) * A_SBO + (r & 7) * 16;

  const int tps = (KMAX * N + TILE - 1) / TILE;
  const int ntiles = B * tps;                  // 32-bit tile arithmetic: 64-bit div/mod is emulated (~100 instr)
  const int stride = (int)gridDim.x * TC_GROUPS;

  // per-tile inputs, fetched one tile ahead: relation count of the sample and this row's 8 features
  struct Pre { int ne; float4 f0, f1; };
  auto fetch = [&](int tile) {
    Pre p;
    p.ne = 0; p.f0 = make_float4(0.f, 0.f, 0.f, 0.f); p.f1 = p.f0;
    if (tile < ntiles) {
      const int b = tile / tps;
      const long long slot = (long long)b * KMAX * N + (tile - b * tps) * TILE + r;
      p.ne = tc::ldg_nc_s32(rowptr + (long long)b * (N + 1) + N);
      if (half == 0) p.f0 = tc::ldg_nc_f4(efeat + slot * 8);   // rows past the sample's last relation read
      p.f1 = tc::ldg_nc_f4(efeat + slot * 8 + 4);              // scratch: harmless, those rows are never stored
    }
    return p;
  };

  PILE_TRACE_DECL();
  int tile = (int)blockIdx.x * TC_GROUPS + g;
  Pre cur = fetch(tile);
  tc::mbar_wait(&S.w_bar, 0);
  uint32_t phase = 0;

  while (tile < ntiles) {
    const int b = tile / tps;
    const int e0 = (tile - b * tps) * TILE;
    const int nrows = min(TILE, cur.ne - e0);
    const long long slot0 = (long long)b * KMAX * N + e0;
    const Pre nxt = fetch(tile + stride);
    if (nrows > 0) {                                     // group-uniform
      PILE_TRACE(1);
      const bool valid = r < nrows;
      if (half == 0) {       // A0 chunk 0 = (x0..x5, 1, 0)
        const bool ok = valid;
        const float f[8] = {ok ? cur.f0.x : 0.f, ok ? cur.f0.y : 0.f, ok ? cur.f0.z : 0.f, ok ? cur.f0.w : 0.f,
                            ok ? cur.f1.x : 0.f, ok ? cur.f1.y : 0.f, 1.f, 0.f};
        store_chunk(a_hi, a_lo, row_off, f);
      } else {               // aux chunk = (1, d, 0, ...)
        const float f[8] = {1.f, valid ? cur.f1.y : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store_chunk(S.aux[g][0], S.aux[g][1], row_off, f);
      }
#pragma unroll 1
      for (int layer = 0; layer < 4; ++layer) {
        PILE_TRACE(2);
        tc::fence_async_smem();          // st.shared of the A tile -> visible to the tensor core
        tc::fence_before_sync();         // our tcgen05.ld of the previous accumulator are complete
        PILE_TRACE(3);
        group_barrier(g);
        PILE_TRACE(4);
        if (wig == g) {                  // (issuer warps of the 4 groups sit on 4 different SM sub-partitions) warp-uniform: the whole warp walks the descriptors, one lane issues
          tc::fence_after_sync();
          const uint32_t elected = tc::elect_one();
          if (layer == 0) {
            issue_layer0(elected, tmem_d, a_hi_u, a_lo_u, zero_u, tc::smem_u32(S.w0[0]), tc::smem_u32(S.w0[1]));
          } else {
            issue_layer<4>(elected, tmem_d, a_hi_u, a_lo_u, aux_hi_u, aux_lo_u, zero_u,
                           tc::smem_u32(S.wl[layer - 1][0]), tc::smem_u32(S.wl[layer - 1][1]));
          }
          if (elected) tc::mma_commit(&S.mma_bar[g]);
          __syncwarp();
        }
        tc::mbar_wait(&S.mma_bar[g], phase);
        phase ^= 1;
        tc::fence_after_sync();
        PILE_TRACE(5);
        if (layer < 3) {
          uint8_t* mk = layer == 0 ? m_re0 : (layer == 1 ? m_re1 : m_re2);
          hidden_epilogue<RECORD>(a_hi, a_lo, row_off, half, taddr, mk, slot0 + r, valid);
        } else {
          float v[2][16];
#pragma unroll
          for (int q = 0; q < 2; ++q) tc::tmem_ld16(taddr + half * 32 + q * 16, v[q]);
          tc::tmem_ld_wait();
          if (valid) {      // C_e stays row-major: k_edge_agg streams it from HBM in whole 256-byte rows
            float* out = Ce + (slot0 + r) * H + half * 32;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
              for (int h = 0; h < 2; ++h) st8(out + q * 16 + h * 8, &v[q][h * 8]);
          }
        }
      }
    }
    PILE_TRACE(6);
    cur = nxt;
    tile += stride;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(S.tmem_base, TMEM_COLS);
}

PILE_TRACE_SETTER(set_edge_trace)

int launch_edge_encode_tc(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                          long long s_stride, const Csr& csr, const Masks* mk, float* efeat, float* Ce, int B, int N,
                          cudaStream_t st, bool efeat_ready) {
  static DeviceOnce once;
  const int once_dev = once.pending();
  if (once_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_encode_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(EdgeTcSmem));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_edge_encode_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(EdgeTcSmem));
    if (e != cudaSuccess) return (int)e;
    once.done(once_dev);
  }
  if (!efeat_ready) {          // relations that did not come from launch_nbr_search (which writes the rows itself)
    const dim3 fgrid((KMAX * N + 255) / 256, B);
    k_edge_features<<<fgrid, 256, 0, st>>>(attr, dens, s_cur, s_stride, csr.rowptr, csr.col, csr.row, efeat, B, N);
    PILE_CHECK_LAUNCH();
  }
  if (g_use_tensor_cores == 2) return launch_edge_encode_tmem(wpack, efeat, csr, mk, Ce, B, N, st);
  const long long ntiles = (long long)B * ((KMAX * N + TILE - 1) / TILE);
  const long long want = (ntiles + TC_GROUPS - 1) / TC_GROUPS;
  const int grid = (int)(want < NSM ? want : NSM);
  if (mk) {
    k_edge_encode_tc<true><<<grid, TC_THREADS, sizeof(EdgeTcSmem), st>>>(wpack, efeat, csr.rowptr, mk->re0, mk->re1,
                                                                         mk->re2, Ce, B, N);
  } else {
    k_edge_encode_tc<false><<<grid, TC_THREADS, sizeof(EdgeTcSmem), st>>>(wpack, efeat, csr.rowptr, nullptr, nullptr,
                                                                          nullptr, Ce, B, N);
  }
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
