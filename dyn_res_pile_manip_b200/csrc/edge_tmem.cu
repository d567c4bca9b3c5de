// K3, second tensor-core variant: the activation operand never touches shared memory.
//
// k_edge_encode_tc (edge_tc.cu) stages every layer's activations in shared memory; with N = 64 an SS-mode
// MMA then needs 6 KB of shared-memory operands per 32 tensor cycles, more than the SM can deliver, and the
// st.shared + fence.proxy.async sit in each tile's serial chain.  Here the A operand lives in TENSOR MEMORY:
//
//   * a group owns 128 TMEM columns = two 64-column regions X, Y used in ping-pong: layer L reads its A from
//     one region and accumulates D into the other;
//   * the epilogue thread (row r = TMEM lane, 32-column half h) loads its 32 accumulator columns, applies
//     ReLU, splits into bf16 hi/lo and writes the packed pairs back IN PLACE with tcgen05.st (16 columns hi,
//     16 columns lo: exactly the 32 columns it just read), so no other thread's data is touched;
//   * biases are pre-loaded into the NEXT accumulator region with tcgen05.st and the layer's MMAs accumulate
//     on top (enable_input_d = 1 from the first MMA);
//   * tcgen05.mma takes A from TMEM ([taddr]) and only the 64x16 weight slice (2 KB) from shared memory.
//
// Column map of an A region at base P (k = input channel, 2 bf16 per 32-bit column):
//   hi(k in 0..31) -> P+0..15, lo(k in 0..31) -> P+16..31, hi(k in 32..63) -> P+32..47, lo(k in 32..63) -> P+48..63
#include "kernels.h"
#include "tc_tile.cuh"
#include "tmap.cuh"

namespace pile {

constexpr uint32_t TM_W0_BYTES = 2 * b_bytes(64, 16);     // 4 KB   [hi | lo] of W0 [64 x 16] (cols: 6 features)
constexpr uint32_t TM_WL_BYTES = 2 * b_bytes(64, 64);     // 16 KB  [hi | lo] of RE1 / RE2 / W_e
constexpr uint32_t TC_EDGE2_BYTES = TM_W0_BYTES + 3 * TM_WL_BYTES;   // 52 KB
static_assert(TC_EDGE2_BYTES == 4 * TC_EDGE2_FLOATS, "TC_EDGE2 slot size (common.cuh) out of sync");

struct EdgeTmemSmem {
  alignas(128) uint8_t w0[TM_W0_BYTES];
  alignas(128) uint8_t wl[3][TM_WL_BYTES];
  float b_re0[H], b_re1[H], b_re2[H], wd_rp[H], b_rp[H];
  // packed C_e rows on their way out: the 128-byte upper-half part and the 64-byte third-byte part of every row,
  // each in the swizzle pattern of its tensor map (one TMA tile store per part and tile)
  alignas(1024) uint8_t stage_hi[TC_GROUPS][TILE * 128];
  alignas(1024) uint8_t stage_lo[TC_GROUPS][TILE * 64];
  alignas(16) float4 feat[TC_GROUPS][2][TILE][2];         // next tile's input rows, one private copy per thread
  uint64_t bar[TC_GROUPS];
  uint64_t w_bar;
  uint32_t tmem_base;
};

constexpr int EDGE_TMEM_SMEM = (int)sizeof(EdgeTmemSmem) + 1024;          // + alignment slack

__device__ __forceinline__ void mma_bf16_ts_if(uint32_t pred, uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%6, %6, %6, %6}, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(pred), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16 fp32 values -> 16 32-bit words (this thread's TMEM lane) starting at column taddr
__device__ __forceinline__ void tmem_st16f(uint32_t taddr, const float* v) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[i]);
  tmem_st16(taddr, r);
}

// hidden layer: D (+)= A(TMEM region pa) * W^T, three passes x four K-steps; dw = descriptor of the weight image's
// hi part (the lo part follows TM_WL_BYTES / 2 later), built once per kernel so that an MMA costs one add
__device__ __forceinline__ void issue_hidden_ts(uint32_t elected, uint32_t d, uint32_t pa, uint64_t dw,
                                                uint32_t first_accumulates) {
  constexpr uint32_t idesc = tc::make_idesc_bf16(TILE, H);
  constexpr uint32_t BL = b_lbo(H);
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t part = pass == 1 ? 16u : 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t a = pa + (k < 2 ? k * 8 : 32 + (k - 2) * 8) + part;
      mma_bf16_ts_if(elected, d, a, tc::desc_advance(dw, (pass == 2 ? TM_WL_BYTES / 2 : 0u) + k * 2 * BL), idesc,
                     (pass | k) != 0 ? 1u : first_accumulates);
    }
  }
}

// q = n / d for n * d < 2^32 with magic = 2^32 / d + 1 (host: div_magic); magic 0 = plain division
__device__ __forceinline__ int fast_div(int n, int d, uint32_t magic) {
  return magic ? (int)__umulhi((uint32_t)n, magic) : n / d;
}

template <bool RECORD>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_edge_encode_tmem(const float* __restrict__ wpack, const float* __restrict__ efeat, const int* __restrict__ rowptr,
                   uint8_t* __restrict__ m_re0, uint8_t* __restrict__ m_re1, uint8_t* __restrict__ m_re2,
                   const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, int B, int N,
                   uint32_t tps_magic) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the swizzled staging boxes need 1024-byte alignment in the shared window
  EdgeTmemSmem& S = *reinterpret_cast<EdgeTmemSmem*>(
      smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int g = threadIdx.x / GROUP_THREADS, t = threadIdx.x % GROUP_THREADS;
  const int wig = t >> 5;
  const int r = (wig & 3) * 32 + (t & 31);
  const int half = wig >> 2;

  if (threadIdx.x < 32) tc::tmem_alloc(&S.tmem_base, 512);
  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_GROUPS; ++i) tc::mbar_init(&S.bar[i], 1);
    tc::mbar_init(&S.w_bar, 1);
    tc::mbar_init_fence();
  }
  load_block(S.b_re0, wpack + wslot_offset(B_RE0), H);
  load_block(S.b_re1, wpack + wslot_offset(B_RE1), H);
  load_block(S.b_re2, wpack + wslot_offset(B_RE2), H);
  load_block(S.wd_rp, wpack + wslot_offset(WD_RP), H);
  load_block(S.b_rp, wpack + wslot_offset(B_RP), H);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(&S.w_bar, TC_EDGE2_BYTES);
    tc::bulk_g2s(S.w0, wpack + wslot_offset(TC_EDGE2), TC_EDGE2_BYTES, &S.w_bar);
  }

  const uint32_t lane_off = (uint32_t)((wig & 3) * 32) << 16;
  const uint32_t X = S.tmem_base + g * 128, Y = X + 64;          // column bases (lane field 0: MMA operands)
  const uint32_t Xt = X + lane_off, Yt = Y + lane_off;           // this warp's lane quarter (ld / st)
  uint64_t* bar = &S.bar[g];
  uint32_t phase = 0;

  const int tps = (KMAX * N + TILE - 1) / TILE;
  const int ntiles = B * tps;                  // 32-bit tile arithmetic: 64-bit div/mod is emulated (~100 instr)
  const int stride = (int)gridDim.x * TC_GROUPS;
  // The next tile's relation rows (32 bytes each) are fetched with cp.async into a slot that only the issuing
  // thread reads, so no registers are tied up across the tile chain and no barrier is needed
  float4* my_feat = &S.feat[g][half][r][0];
  const uint32_t my_feat_u32 = tc::smem_u32(my_feat);
  int nxt_b = 0;                        // sample of the tile fetched last
  auto fetch = [&](int tile) {          // -> relation count of the tile's sample
    int ne = 0;
    if (tile < ntiles) {
      const int b = fast_div(tile, tps, tps_magic);
      nxt_b = b;
      const long long slot = (long long)b * KMAX * N + (tile - b * tps) * TILE + r;
      ne = tc::ldg_nc_s32(rowptr + (long long)b * (N + 1) + N);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(my_feat_u32), "l"(efeat + slot * 8) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(my_feat_u32 + 16), "l"(efeat + slot * 8 + 4) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    return ne;
  };
  // hand the MMAs of one layer to the tensor core and wait for them
  PILE_TRACE_DECL();
  auto run = [&](auto issue, auto during) {
    PILE_TRACE(2);
    tmem_st_wait();                  // this thread's tcgen05.st (A operand, bias pre-load) have landed
    tc::fence_before_sync();
    PILE_TRACE(3);
    group_barrier(g);
    PILE_TRACE(4);
    if (wig == g) {          // issuer warp g*8+g: the four groups issue from four different SM sub-partitions
      tc::fence_after_sync();
      const uint32_t elected = tc::elect_one();
      issue(elected);
      if (elected) tc::mma_commit(bar);
      __syncwarp();
    }
    during();                        // work that overlaps the tensor pipe
    tc::mbar_wait(bar, phase);
    phase ^= 1;
    tc::fence_after_sync();
    PILE_TRACE(5);
  };
  auto nothing = [] {};
  // ReLU + hi/lo split of this thread's 32 accumulator columns of region `reg`, written back in place as the
  // next layer's A operand; then pre-load `bias_next` (nullable) into the other region `other`
  auto epilogue = [&](uint32_t reg, uint32_t other, const float* bias_next, float bias_scale_d, const float* wd_next,
                      uint8_t* __restrict__ mask, long long mrow, bool valid) {
    float v[2][16];
    tc::tmem_ld16(reg + half * 32, v[0]);
    tc::tmem_ld16(reg + half * 32 + 16, v[1]);
    tc::tmem_ld_wait();
    uint32_t hi[16], lo[16], mbits = 0;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float a = v[q][j], b = v[q][j + 1];
        if (RECORD) mbits |= (a > 0.f ? 1u : 0u) << (q * 16 + j) | (b > 0.f ? 1u : 0u) << (q * 16 + j + 1);
        tc::split2_relu(a, b, hi[q * 8 + j / 2], lo[q * 8 + j / 2]);
      }
    tmem_st16(reg + half * 32, hi);
    tmem_st16(reg + half * 32 + 16, lo);
    if (RECORD && valid) *reinterpret_cast<uint32_t*>(mask + mrow * 8 + half * 4) = mbits;
    if (bias_next != nullptr) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float bv[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 b4 = ld4(bias_next + half * 32 + q * 16 + j);
          bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
          if (wd_next != nullptr) {
            const float4 w4 = ld4(wd_next + half * 32 + q * 16 + j);
            bv[j] = fmaf(w4.x, bias_scale_d, bv[j]); bv[j + 1] = fmaf(w4.y, bias_scale_d, bv[j + 1]);
            bv[j + 2] = fmaf(w4.z, bias_scale_d, bv[j + 2]); bv[j + 3] = fmaf(w4.w, bias_scale_d, bv[j + 3]);
          }
        }
        tmem_st16f(other + half * 32 + q * 16, bv);
      }
    }
  };

  int tile = (int)blockIdx.x * TC_GROUPS + g;
  int cur_ne = fetch(tile);
  tc::mbar_wait(&S.w_bar, 0);
  const uint64_t dw0 = tc::make_desc(tc::smem_u32(S.w0), b_lbo(H), B_SBO);
  const uint64_t dwl0 = tc::make_desc(tc::smem_u32(S.wl[0]), b_lbo(H), B_SBO);
  const uint64_t dwl1 = tc::desc_advance(dwl0, TM_WL_BYTES), dwl2 = tc::desc_advance(dwl0, 2 * TM_WL_BYTES);
  // C_e rows of the previous tile wait in the staging boxes and leave (two TMA tile stores issued by one thread
  // of a non-issuer warp) while the next tile's first MMAs run
  bool pending = false;
  int p_b = 0, p_e0 = 0;
  const bool store_thread = t == 32 * ((g + 4) & 7);
  uint8_t* const my_hi = S.stage_hi[g] + r * 128;
  uint8_t* const my_lo = S.stage_lo[g] + r * 64;
  const uint32_t sw_hi = (uint32_t)(r & 7), sw_lo = (uint32_t)((r >> 1) & 3);
  const uint32_t stage_hi_u32 = tc::smem_u32(S.stage_hi[g]), stage_lo_u32 = tc::smem_u32(S.stage_lo[g]);
  auto flush = [&] {          // after a group barrier that follows the staging writes
    if (store_thread) {
      tc::tma_store_3d(&tm_hi, stage_hi_u32, 0, p_e0, p_b);
      tc::tma_store_3d(&tm_lo, stage_lo_u32, 0, p_e0, p_b);
      tc::bulk_commit();
    }
  };

  while (tile < ntiles) {
    const int b = nxt_b;
    const int e0 = (tile - b * tps) * TILE;
    const int nrows = min(TILE, cur_ne - e0);
    const long long slot0 = (long long)b * KMAX * N + e0;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const float4 f0 = my_feat[0], f1 = my_feat[1];
    const int nxt_ne = fetch(tile + stride);          // overwrites the slot just read
    if (nrows > 0) {
      PILE_TRACE(1);
      const bool valid = r < nrows;
      const float d = f1.y;
      // layer-0 operand (K = 16: 6 features, zeros) into X; bias_0 pre-loaded into Y
      if (half == 0) {
        uint32_t hi[8], lo[8];
        const float f[8] = {valid ? f0.x : 0.f, valid ? f0.y : 0.f, valid ? f0.z : 0.f, valid ? f0.w : 0.f,
                            valid ? f1.x : 0.f, valid ? f1.y : 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) tc::split2(f[2 * i], f[2 * i + 1], hi[i], lo[i]);
#pragma unroll
        for (int i = 4; i < 8; ++i) { hi[i] = 0u; lo[i] = 0u; }
        tmem_st8(Xt, hi);
        tmem_st8(Xt + 16, lo);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float bv[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 b4 = ld4(S.b_re0 + half * 32 + q * 16 + j);
          bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
        }
        tmem_st16f(Yt + half * 32 + q * 16, bv);
      }
      // layer 0: A = X (one K-step), D = Y
      run([&](uint32_t el) {
        constexpr uint32_t idesc = tc::make_idesc_bf16(TILE, H);
        const uint64_t d_lo = tc::desc_advance(dw0, TM_W0_BYTES / 2);
        mma_bf16_ts_if(el, Y, X, dw0, idesc, 1u);
        mma_bf16_ts_if(el, Y, X + 16, dw0, idesc, 1u);
        mma_bf16_ts_if(el, Y, X, d_lo, idesc, 1u);
      }, [&] {
        if (pending) flush();
        pending = false;
      });
      epilogue(Yt, Xt, S.b_re1, 0.f, nullptr, m_re0, slot0 + r, valid);
      // layer 1: A = Y, D = X
      run([&](uint32_t el) { issue_hidden_ts(el, X, Y, dwl0, 1u); }, nothing);
      epilogue(Xt, Yt, S.b_re2, 0.f, nullptr, m_re1, slot0 + r, valid);
      // layer 2: A = X, D = Y; then pre-load the hoisted constant w_d d + b of the propagator into X
      run([&](uint32_t el) { issue_hidden_ts(el, Y, X, dwl1, 1u); }, nothing);
      epilogue(Yt, Xt, S.b_rp, d, S.wd_rp, m_re2, slot0 + r, valid);
      // layer E: A = Y, D = X -> C_e rows.  The previous tile's stores have long finished reading the staging
      // boxes; the wait makes that formal before this layer's group barrier releases the writers below
      if (store_thread) tc::bulk_wait_read0();
      run([&](uint32_t el) { issue_hidden_ts(el, X, Y, dwl2, 1u); }, nothing);
      {
        float v[2][16];
        tc::tmem_ld16(Xt + half * 32, v[0]);
        tc::tmem_ld16(Xt + half * 32 + 16, v[1]);
        tc::tmem_ld_wait();
        // this thread's 32 columns of row r, packed to 24 bits (tc_tile.cuh: pack24) in registers: four 16-byte
        // chunks of the row's upper-half part and two of its third-byte part, written at the swizzled positions
        // the tensor maps expect (chunk index XOR row bits: conflict-free for one row per lane)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint2 h[4];
          uint32_t l[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            pack24(make_float4(v[q][4 * i], v[q][4 * i + 1], v[q][4 * i + 2], v[q][4 * i + 3]), h[i], l[i]);
          const uint32_t c = (uint32_t)(half * 4 + q * 2);
          *reinterpret_cast<uint4*>(my_hi + ((c ^ sw_hi) << 4)) = make_uint4(h[0].x, h[0].y, h[1].x, h[1].y);
          *reinterpret_cast<uint4*>(my_hi + (((c + 1) ^ sw_hi) << 4)) = make_uint4(h[2].x, h[2].y, h[3].x, h[3].y);
          *reinterpret_cast<uint4*>(my_lo + (((uint32_t)(half * 2 + q) ^ sw_lo) << 4)) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        tc::fence_async_smem();          // generic-proxy writes -> visible to the copy engine
        pending = true;
        p_b = b;
        p_e0 = e0;
      }
    }
    PILE_TRACE(6);
    cur_ne = nxt_ne;
    tile += stride;
  }
  if (pending) {             // group-uniform
    group_barrier(g);
    flush();
  }
  if (store_thread) tc::bulk_wait0();
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(S.tmem_base, 512);
}

PILE_TRACE_SETTER(set_edge_tmem_trace)

int launch_edge_encode_tmem(const float* wpack, const float* efeat, const Csr& csr, const Masks* mk, float* Ce,
                            int B, int N, cudaStream_t st) {
  static DeviceOnce once;
  const int once_dev = once.pending();
  if (once_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_encode_tmem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         EDGE_TMEM_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_edge_encode_tmem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EDGE_TMEM_SMEM);
    if (e != cudaSuccess) return (int)e;
    once.done(once_dev);
  }
  const int tps = (KMAX * N + TILE - 1) / TILE;
  const long long ntiles = (long long)B * tps;
  const long long want = (ntiles + TC_GROUPS - 1) / TC_GROUPS;
  const int grid = (int)(want < NSM ? want : NSM);
  // tile -> sample by multiply-high (exact while (largest tile index) * tps < 2^32), else plain division
  const unsigned long long nmax = (unsigned long long)ntiles + (unsigned long long)grid * TC_GROUPS;
  const uint32_t magic = tps >= 2 && nmax * (unsigned long long)tps < (1ull << 32) ? (uint32_t)((1ull << 32) / tps + 1) : 0u;
  // packed C_e rows = [128 bytes of upper halves | 64 bytes of third bytes]; two maps over the same rows,
  // dimensions (bytes of the part, slot in the sample, sample): the engine clips a sample's last tile at KMAX*N
  static TmapCache<2> cache;
  CUtensorMap tm[2];
  const int ts = cache.get(Ce, B, N, tm, [&](CUtensorMap* m) {
    const uint64_t row = CE_PACKED_ROW, plane = (uint64_t)KMAX * N * CE_PACKED_ROW;
    int e = tmap_encode_rows(&m[0], Ce, 128, (uint64_t)KMAX * N, row, (uint64_t)B, plane, TILE);
    if (e) return e;
    return tmap_encode_rows(&m[1], reinterpret_cast<uint8_t*>(Ce) + 128, 64, (uint64_t)KMAX * N, row, (uint64_t)B, plane, TILE);
  });
  if (ts) return ts;
  if (mk)
    k_edge_encode_tmem<true><<<grid, TC_THREADS, EDGE_TMEM_SMEM, st>>>(wpack, efeat, csr.rowptr, mk->re0, mk->re1,
                                                                       mk->re2, tm[0], tm[1], B, N, magic);
  else
    k_edge_encode_tmem<false><<<grid, TC_THREADS, EDGE_TMEM_SMEM, st>>>(wpack, efeat, csr.rowptr, nullptr, nullptr,
                                                                        nullptr, tm[0], tm[1], B, N, magic);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
