// K2/K4/K5 on the tensor cores: particle encoder and the per-propagation-step particle update.
//
// The FP32 k_propagate (fwd.cu) interleaves a memory-bound gather with FFMA GEMMs inside one CTA; here
// the propagation step (reference model/gnn_dyn.py:182-193) is split into
//   k_edge_agg          agg[i] = sum_{e=(i<-j)} ReLU(C_e[e] + P_r[i] + P_s[j])        pure HBM/L2 streaming
//   k_node_update_tc    eff <- ReLU(W_a agg + C_p + eff);  (P_r, P_s) <- (W_r, W_s) eff  tcgen05 tiles
//                       last step: s_pred = s_cur + V1 ReLU(V0 eff + c0) + c1            (gnn_dyn.py:196-198)
// and the particle encoder (gnn_dyn.py:174-176) becomes k_node_encode_tc.  GEMMs use the bf16 hi/lo split
// (three passes, fp32 accumulation in TMEM) and fold biases through the aux K chunk, see tc_tile.cuh.
#include "kernels.h"
#include "tc_tile.cuh"

namespace pile {

// ---- TC_NODE weight slot: [hi | lo] canonical images, in this order ---------------------------------
constexpr uint32_t NB_PE0 = 2 * b_bytes(64, 16);    //   4 KB  PE0aug  [64 x 16]  cols: s_delta(3), attr, d, bias
constexpr uint32_t NB_K80 = 2 * b_bytes(64, 80);    //  20 KB  PE1aug / WPaug / V0aug [64 x 80]
constexpr uint32_t NB_WRS = 2 * b_bytes(128, 64);   //  32 KB  [W_r ; W_s]  [128 x 64]
constexpr uint32_t NB_WA = 2 * b_bytes(64, 64);     //  16 KB  W_a  [64 x 64]
constexpr uint32_t NB_V1 = 2 * b_bytes(16, 80);     //   5 KB  V1aug [16 x 80]
constexpr uint32_t OFF_PE0 = 0, OFF_PE1 = OFF_PE0 + NB_PE0, OFF_WP = OFF_PE1 + NB_K80, OFF_WRS = OFF_WP + NB_K80,
                   OFF_WA = OFF_WRS + NB_WRS, OFF_V0 = OFF_WA + NB_WA, OFF_V1 = OFF_V0 + NB_K80,
                   TC_NODE_BYTES = OFF_V1 + NB_V1;
static_assert(TC_NODE_BYTES == 4 * TC_NODE_FLOATS, "TC_NODE slot size (common.cuh) out of sync");

constexpr uint32_t NODE_TMEM_PER_GROUP = 128;

template <uint32_t WBYTES>
struct NodeTcSmem {
  alignas(128) uint8_t w[WBYTES];
  GroupTile t[TC_GROUPS];
  alignas(128) uint8_t zero[A_LBO];
  uint64_t bar[TC_GROUPS];
  uint64_t w_bar;
  uint32_t tmem_base;
};
using NodeEncSmemTc = NodeTcSmem<OFF_WA>;                 // PE0, PE1, WP, WRS = 76 KB
using NodeUpdSmemTc = NodeTcSmem<NB_WRS + NB_WA>;         // 48 KB (non-last: WRS, WA; last: WA, V0, V1 = 41 KB)
static_assert(sizeof(NodeEncSmemTc) <= 227 * 1024, "shared memory budget");

template <typename Smem>
__device__ __forceinline__ GroupCtx tile_prologue(Smem& S, const float* wpack, uint32_t w_off, uint32_t w_bytes) {
  const int g = threadIdx.x / GROUP_THREADS, t = threadIdx.x % GROUP_THREADS;
  if (threadIdx.x < 32) tc::tmem_alloc(&S.tmem_base, TC_GROUPS * NODE_TMEM_PER_GROUP);
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
    tc::mbar_expect_tx(&S.w_bar, w_bytes);
    tc::bulk_g2s(S.w, reinterpret_cast<const uint8_t*>(wpack + wslot_offset(TC_NODE)) + w_off, w_bytes, &S.w_bar);
  }
  GroupCtx c;
  c.g = g;
  c.wig = t >> 5;
  c.tmem_d = S.tmem_base + g * NODE_TMEM_PER_GROUP;
  c.taddr = c.tmem_d + ((uint32_t)((c.wig & 3) * 32) << 16);
  c.a_hi = tc::smem_u32(S.t[g].a[0]);
  c.a_lo = tc::smem_u32(S.t[g].a[1]);
  c.aux_hi = tc::smem_u32(S.t[g].aux[0]);
  c.aux_lo = tc::smem_u32(S.t[g].aux[1]);
  c.zero = tc::smem_u32(S.zero);
  c.bar = &S.bar[g];
  c.phase = 0;
  return c;
}

template <typename Smem>
__device__ __forceinline__ void tile_epilogue(Smem& S) {
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(S.tmem_base, TC_GROUPS * NODE_TMEM_PER_GROUP);
}

// accumulator columns [0, 64) = P_r, [64, 128) = P_s of the group's tile -> row-major global arrays, one array at
// a time through the staging tile (the A tile: its last MMA has completed).  Ends with the group barrier that
// frees the tile for the next A operand.
__device__ __forceinline__ void store_pr_ps(const GroupCtx& c, uint8_t* stage, int t, int r, int half,
                                            float* __restrict__ Pr, float* __restrict__ Ps, long long row0, long long R,
                                            int packed_ps) {
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
#pragma unroll 1
    for (int q = 0; q < 2; ++q) {
      float v[16];
      tc::tmem_ld16(c.taddr + which * 64 + half * 32 + q * 16, v);
      tc::tmem_ld_wait();
      stage_put16(stage, r, half * 32 + q * 16, v);
    }
    group_barrier(c.g);
    // P_s rows are gathered once per relation by k_edge_agg (through L2, next to the C_e stream): with the tensor
    // engine's packed C_e they are stored in the same 24-bit row format
    if (which == 1 && packed_ps) stage_flush_packed(stage, t, reinterpret_cast<uint8_t*>(Ps), row0, R);
    else stage_flush(stage, t, which == 0 ? Pr : Ps, row0, R);
    group_barrier(c.g);
  }
}

// ------------------------------------------------------------------------------------------------
// particle encoder: p_enc (= effect_0), C_p, and (P_r, P_s) of propagation step 0
// ------------------------------------------------------------------------------------------------
template <bool RECORD>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_node_encode_tc(const float* __restrict__ wpack, const float* __restrict__ attr, const float* __restrict__ dens,
                 const float* __restrict__ s_delta, uint8_t* __restrict__ m_pe0, uint8_t* __restrict__ m_pe1,
                 float* __restrict__ Cp, float* __restrict__ eff, float* __restrict__ Pr, float* __restrict__ Ps,
                 int B, int N, int packed_ps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  NodeEncSmemTc& S = *reinterpret_cast<NodeEncSmemTc*>(smem_raw);
  GroupCtx c = tile_prologue(S, wpack, OFF_PE0, OFF_WA);
  const int g = c.g, wig = c.wig, t = threadIdx.x % GROUP_THREADS;
  const int r = (wig & 3) * 32 + (threadIdx.x & 31), half = wig >> 2;
  uint8_t* const a_hi = S.t[g].a[0];
  uint8_t* const a_lo = S.t[g].a[1];
  const uint32_t row_off = (r >> 3) * A_SBO + (r & 7) * 16;
  const uint32_t w = tc::smem_u32(S.w);
  const int R = B * N;                         // 32-bit row arithmetic (launcher checks B*N < 2^31)
  const int ntiles = (R + TILE - 1) / TILE;
  tc::mbar_wait(&S.w_bar, 0);

  // group g of CTA c takes tiles g * gridDim + c, + 4 * gridDim, ...: a small workload spreads over all SMs (one
  // tile chain per SM) before any SM runs four chains side by side
  for (int tile = g * (int)gridDim.x + (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x * TC_GROUPS) {
    const int row = tile * TILE + r;
    const bool valid = row < R;
    float d = 0.f;
    if (valid) d = dens[row / N] / 5000.f;
    if (half == 0) {
      float f[8] = {0.f, 0.f, 0.f, 0.f, d, 1.f, 0.f, 0.f};
      if (valid) { const float* sd = s_delta + (long long)row * 3; f[0] = sd[0]; f[1] = sd[1]; f[2] = sd[2]; f[3] = attr[row]; }
      store_chunk(a_hi, a_lo, row_off, f);
    } else {
      const float f[8] = {1.f, d, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      store_chunk(S.t[g].aux[0], S.t[g].aux[1], row_off, f);
    }
    // PE layer 0
    run_gemm(c, [&](uint32_t el) { issue_gemm_k16<64>(el, c.tmem_d, c.a_hi, c.a_lo, c.zero, w + OFF_PE0, w + OFF_PE0 + NB_PE0 / 2); });
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float v[16];
      tc::tmem_ld16(c.taddr + half * 32 + q * 16, v);
      tc::tmem_ld_wait();
      const uint32_t m = relu_to_tile<RECORD>(a_hi, a_lo, row_off + (half * 4 + q * 2) * A_LBO, v);
      if (RECORD && valid) *reinterpret_cast<uint16_t*>(m_pe0 + (long long)row * 8 + half * 4 + q * 2) = (uint16_t)m;
    }
    // PE layer 1 -> particle_encode = effect_0
    run_gemm(c, [&](uint32_t el) {
      issue_gemm<64, 4, true>(el, c.tmem_d, c.a_hi, c.a_lo, c.aux_hi, c.aux_lo, c.zero, w + OFF_PE1, w + OFF_PE1 + NB_K80 / 2);
    });
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float v[16];
      tc::tmem_ld16(c.taddr + half * 32 + q * 16, v);
      tc::tmem_ld_wait();
      const uint32_t m = relu_to_tile<RECORD>(a_hi, a_lo, row_off + (half * 4 + q * 2) * A_LBO, v);
      if (valid) {
        st16_tbc(eff, tile, r, half * 4 + q * 2, v);
        if (RECORD) *reinterpret_cast<uint16_t*>(m_pe1 + (long long)row * 8 + half * 4 + q * 2) = (uint16_t)m;
      }
    }
    // C_p = W_p p_enc + w_d d + b
    run_gemm(c, [&](uint32_t el) {
      issue_gemm<64, 4, true>(el, c.tmem_d, c.a_hi, c.a_lo, c.aux_hi, c.aux_lo, c.zero, w + OFF_WP, w + OFF_WP + NB_K80 / 2);
    });
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float v[16];
      tc::tmem_ld16(c.taddr + half * 32 + q * 16, v);
      tc::tmem_ld_wait();
      if (valid) st16_tbc(Cp, tile, r, half * 4 + q * 2, v);
    }
    // (P_r, P_s) = (W_r, W_s) p_enc : one N = 128 product, column half 0 -> P_r, half 1 -> P_s
    run_gemm(c, [&](uint32_t el) {
      issue_gemm<128, 4, false>(el, c.tmem_d, c.a_hi, c.a_lo, c.aux_hi, c.aux_lo, c.zero, w + OFF_WRS, w + OFF_WRS + NB_WRS / 2);
    });
    store_pr_ps(c, a_hi, t, r, half, Pr, Ps, (long long)tile * TILE, R, packed_ps);
  }
  tile_epilogue(S);
}

// ------------------------------------------------------------------------------------------------
// receiver-segmented sum of the relation effects (memory-bound; no weights, high occupancy)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 add2x2(const float4& a, const float4& b) {
  float4 r;
  asm("{\n\t"
      ".reg .b64 a01, a23, b01, b23;\n\t"
      "mov.b64 a01, {%4, %5};\n\t"
      "mov.b64 a23, {%6, %7};\n\t"
      "mov.b64 b01, {%8, %9};\n\t"
      "mov.b64 b23, {%10, %11};\n\t"
      "add.rn.f32x2 a01, a01, b01;\n\t"
      "add.rn.f32x2 a23, a23, b23;\n\t"
      "mov.b64 {%0, %1}, a01;\n\t"
      "mov.b64 {%2, %3}, a23;\n\t"
      "}\n"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
      : "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w));
  return r;
}

// A/B switch (tools/ab_step.py with a variant library): -DPILE_PS_FP32 keeps the gathered P_s rows in fp32 on inference runs too
#ifdef PILE_PS_FP32
constexpr bool PS_PACK_OK = false;
#else
constexpr bool PS_PACK_OK = true;
#endif
constexpr int AGG_THREADS = 256;
template <bool PACKED>
__host__ __device__ constexpr int agg_edge_bytes(bool packed_ps) {      // C_e row + P_s row
  return (PACKED ? CE_PACKED_ROW : H * 4) + (packed_ps ? CE_PACKED_ROW : H * 4);
}
template <bool PACKED>
__host__ __device__ constexpr int agg_smem_bytes(bool packed_ps) { return (AGG_THREADS / 16) * KMAX * agg_edge_bytes<PACKED>(packed_ps); }

template <bool RECORD, bool PACKED>
__global__ void __launch_bounds__(AGG_THREADS)
k_edge_agg(const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ Ce,
           const float* __restrict__ Pr, const float* __restrict__ Ps, uint8_t* __restrict__ m_edge,
           float* __restrict__ agg, int B, int N) {
  // A half-warp owns one receiver at a time; lane l16 owns channels 4*l16 .. 4*l16+3.  The kernel streams C_e once
  // and gathers one P_s row per relation.  The rows do not pass through registers: per receiver ONE bulk copy (TMA,
  // cp.async.bulk) brings its contiguous C_e rows into the half-warp's shared-memory slab, and the P_s rows of its
  // relations follow as 16-byte cp.async pieces (lane l16 copies piece l16 of every row; the sender indices go round by
  // shuffle); completion of both is counted on the half-warp's mbarrier, then the rows are summed from the slab.  The
  // first two levels of the dependent chain rowptr -> col -> rows are prefetched (rowptr two receivers ahead, col and
  // P_r one ahead).  All loop bounds are warp-uniform so that the two half-warps stay converged.  Variants that were
  // measured and lost (fewer instructions, register-fed rows, L2 prefetch, other occupancies): DESIGN.md section 6.
  extern __shared__ __align__(128) unsigned char agg_smem[];
  __shared__ uint64_t bars[AGG_THREADS / 16];
  // P_s rows are packed like C_e when no tape is recorded; with a tape (gradient runs) they stay fp32, because the
  // extra rounding in front of the ReLU flips sign bits and triples the gradient noise (5.9e-3 vs 2e-3 relative)
  constexpr bool PS_PACKED = PACKED && !RECORD && PS_PACK_OK;
  constexpr int CE_ROW = PACKED ? CE_PACKED_ROW : H * 4;
  constexpr int PS_ROW = PS_PACKED ? CE_PACKED_ROW : H * 4;
  constexpr int EDGE_BYTES = agg_edge_bytes<PACKED>(PS_PACKED);
  const int hw = threadIdx.x >> 4;
  unsigned char* slab_ce = agg_smem + hw * (KMAX * EDGE_BYTES);
  unsigned char* slab_ps = slab_ce + KMAX * CE_ROW;
  uint64_t* bar = &bars[hw];
  const int l16 = threadIdx.x & 15;
  constexpr unsigned FULL = 0xffffffffu;
  const int R = B * N;                                   // 32-bit: 64-bit div/mod is emulated
  const int nhw = (int)gridDim.x * (int)(blockDim.x >> 4);
  // arrivals per phase: lane 0's arrive.expect_tx (the C_e bulk copy) + one asynchronous arrive per lane (its cp.async
  // pieces of the P_s rows have landed)
  if (l16 == 0) tc::mbar_init(bar, 17);
  tc::mbar_init_fence();
  __syncthreads();
  uint32_t phase = 0;
  struct Seg { int b, e_lo, e_hi; };          // e_hi stays raw: its subtraction would wait for the prefetching load
  auto load_seg = [&](int nd) {
    Seg s;
    s.b = 0; s.e_lo = 0; s.e_hi = 0;
    if (nd < R) {
      s.b = nd / N;
      const int* rp = rowptr + (long long)s.b * (N + 1) + (nd - s.b * N);
      s.e_lo = __ldg(rp);
      s.e_hi = __ldg(rp + 1);
    }
    return s;
  };
  int node = (int)blockIdx.x * (int)(blockDim.x >> 4) + hw;
  Seg cur = load_seg(node), nxt = load_seg(node + nhw);
  int mycol = 0;
  float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);
  if (node < R) {
    if (l16 < cur.e_hi - cur.e_lo) mycol = __ldg(col + (long long)cur.b * KMAX * N + cur.e_lo + l16);
    pr = ld4(Pr + (long long)node * H + 4 * l16);
  }
  while (__any_sync(FULL, node < R)) {
    const int cnt = cur.e_hi - cur.e_lo;
    const long long slot = (long long)cur.b * KMAX * N + cur.e_lo;
    const int cntw = max(cnt, __shfl_xor_sync(FULL, cnt, 16));
    // all rows of this receiver
    if (l16 == 0) {
      tc::mbar_expect_tx(bar, (uint32_t)(cnt * CE_ROW));
      if (cnt > 0) tc::bulk_g2s(slab_ce, reinterpret_cast<const unsigned char*>(Ce) + slot * CE_ROW, (uint32_t)(cnt * CE_ROW), bar);
    }
    __syncwarp();
    // the gathered P_s rows as 16-byte cp.async pieces: lane l16 (< PS_ROW / 16) takes piece l16 of every row, so the
    // only per-row work is the sender's index from the lane that holds it and one 64-bit multiply-add (a per-row bulk
    // copy needs uniform registers, i.e. one serialised elect / broadcast / UBLKCP round per row: a quarter of the
    // kernel's instructions and stall samples)
    {
      constexpr int CH = PS_ROW / 16;
      // 32-bit byte offsets into P_s (the launcher checks rows * 256 < 2^32): one multiply-add per row instead of a
      // 64-bit multiply chain
      const unsigned char* ps_bytes = reinterpret_cast<const unsigned char*>(Ps);
      const uint32_t row0_off = (uint32_t)(cur.b * N) * (uint32_t)PS_ROW + (uint32_t)(l16 * 16);
      const uint32_t slab_lane = tc::smem_u32(slab_ps) + (uint32_t)(l16 * 16);
      const int src_lane = threadIdx.x & 16;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k >= cntw) break;          // warp-uniform (the shuffle needs all lanes)
        const int c = __shfl_sync(FULL, mycol, src_lane + k);
        if (k < cnt && l16 < CH)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slab_lane + (uint32_t)(k * PS_ROW)),
                       "l"(ps_bytes + (row0_off + (uint32_t)c * (uint32_t)PS_ROW))
                       : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
    }
    // prefetch for the next two receivers of this half-warp while the rows are on their way
    const int n1 = node + nhw;
    int mycol1 = 0;
    float4 pr1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n1 < R) {
      if (l16 < nxt.e_hi - nxt.e_lo) mycol1 = __ldg(col + (long long)nxt.b * KMAX * N + nxt.e_lo + l16);
      pr1 = ld4(Pr + (long long)n1 * H + 4 * l16);
    }
    const Seg nn = load_seg(n1 + nhw);
    tc::mbar_wait(bar, phase);
    phase ^= 1;

    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k >= cntw) break;          // warp-uniform
      const bool act = k < cnt;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act) {
        const unsigned char* e = slab_ce + k * CE_ROW;
        float4 ce;
        if (PACKED)
          ce = unpack24(*reinterpret_cast<const uint2*>(e + l16 * 8), *reinterpret_cast<const uint32_t*>(e + 128 + l16 * 4));
        else
          ce = *reinterpret_cast<const float4*>(e + l16 * 16);
        const unsigned char* pe = slab_ps + k * PS_ROW;
        float4 ps;
        if (PS_PACKED)
          ps = unpack24(*reinterpret_cast<const uint2*>(pe + l16 * 8), *reinterpret_cast<const uint32_t*>(pe + 128 + l16 * 4));
        else
          ps = *reinterpret_cast<const float4*>(pe + l16 * 16);
        // packed fp32x2 adds (sm_100 FADD2: two IEEE round-to-nearest sums per instruction, same values as scalar adds)
        v = add2x2(add2x2(ce, pr), ps);
        sum = add2x2(sum, make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f)));
      }
      if (RECORD) {          // sign bits of 8 channels per byte: even lanes collect their right neighbour's nibble
        const unsigned m4 = (v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) | (v.w > 0.f ? 8u : 0u);
        const unsigned hi = __shfl_down_sync(FULL, m4, 1);
        if (act && (l16 & 1) == 0) m_edge[(slot + k) * 8 + (l16 >> 1)] = (uint8_t)(m4 | (hi << 4));
      }
    }
    if (node < R) st4(agg + (long long)node * H + 4 * l16, sum);
    __syncwarp();          // the slab is rewritten by the next receiver's copies
    node = n1;
    cur = nxt;
    nxt = nn;
    mycol = mycol1;
    pr = pr1;
  }
}

// ------------------------------------------------------------------------------------------------
// particle update of one propagation step (LAST: + predictor and residual state update)
// ------------------------------------------------------------------------------------------------
template <bool LAST, bool RECORD>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_node_update_tc(const float* __restrict__ wpack, const float* __restrict__ agg, const float* __restrict__ Cp,
                 float* __restrict__ eff, float* __restrict__ PrOut, float* __restrict__ PsOut,
                 uint8_t* __restrict__ m_eff, uint8_t* __restrict__ m_q, const float* __restrict__ s_cur,
                 long long s_stride, float* __restrict__ s_out, long long o_stride, int B, int N, int packed_ps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  NodeUpdSmemTc& S = *reinterpret_cast<NodeUpdSmemTc*>(smem_raw);
#ifdef PILE_ENABLE_TRACE
  const long long t_entry = clock64();
#endif
  // non-last: [WRS | WA]; last: [WA | V0 | V1]
  GroupCtx c = tile_prologue(S, wpack, LAST ? OFF_WA : OFF_WRS, LAST ? (NB_WA + NB_K80 + NB_V1) : (NB_WRS + NB_WA));
  const int g = c.g, wig = c.wig, t = threadIdx.x % GROUP_THREADS;
  const int r = (wig & 3) * 32 + (threadIdx.x & 31), half = wig >> 2;
  uint8_t* const a_hi = S.t[g].a[0];
  uint8_t* const a_lo = S.t[g].a[1];
  const uint32_t row_off = (r >> 3) * A_SBO + (r & 7) * 16;
  const uint32_t w = tc::smem_u32(S.w);
  const uint32_t w_a = LAST ? w : w + NB_WRS;
  const uint32_t w_rs = w;                                  // non-last only
  const uint32_t w_v0 = w + NB_WA, w_v1 = w + NB_WA + NB_K80;   // last only
  const int R = B * N;
  const int ntiles = (R + TILE - 1) / TILE;
  PILE_TRACE_DECL();
#ifdef PILE_ENABLE_TRACE
  if (trace && tr_cap > 0) trace[tr_n++] = (7ll << 56) | (t_entry & 0x00ffffffffffffffLL);
#endif
  PILE_TRACE(8);
  if (LAST && half == 1) {       // constant aux chunk (1, 0, ...) for the predictor biases
    const float f[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    store_chunk(S.t[g].aux[0], S.t[g].aux[1], row_off, f);
  }
  tc::mbar_wait(&S.w_bar, 0);
  PILE_TRACE(9);

  // group g of CTA c takes tiles g * gridDim + c, + 4 * gridDim, ...: a small workload spreads over all SMs (one
  // tile chain per SM) before any SM runs four chains side by side
  for (int tile = g * (int)gridDim.x + (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x * TC_GROUPS) {
    const int row = tile * TILE + r;
    const bool valid = row < R;
    PILE_TRACE(1);
    // A = split(agg tile): whole 128-byte lines per request, then the hi/lo split into the canonical tile
    load_rows_to_tile(agg, (long long)tile * TILE, R, t, a_hi, a_lo);
    // C_p + eff of the first 16 columns are fetched BEFORE the GEMM is handed to the tensor core, so their HBM/L2
    // latency overlaps the MMA; the second 16 columns are fetched while the first are processed
    float x[16], e[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { x[j] = 0.f; e[j] = 0.f; }
    if (valid) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        ld8_tbc(Cp, tile, r, half * 4 + h, x + h * 8);
        ld8_tbc(eff, tile, r, half * 4 + h, e + h * 8);
      }
    }
    PILE_TRACE(2);
    run_gemm(c, [&](uint32_t el) {
      issue_gemm<64, 4, false>(el, c.tmem_d, c.a_hi, c.a_lo, c.aux_hi, c.aux_lo, c.zero, w_a, w_a + NB_WA / 2);
    });
    PILE_TRACE(5);
    // eff <- ReLU(W_a agg + C_p + eff)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float v[16];
      tc::tmem_ld16(c.taddr + half * 32 + q * 16, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += x[j] + e[j];
      if (q == 0) {          // prefetch the second half now that x / e are consumed
#pragma unroll
        for (int j = 0; j < 16; ++j) { x[j] = 0.f; e[j] = 0.f; }
        if (valid) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            ld8_tbc(Cp, tile, r, half * 4 + 2 + h, x + h * 8);
            ld8_tbc(eff, tile, r, half * 4 + 2 + h, e + h * 8);
          }
        }
      }
      const uint32_t m = relu_to_tile<RECORD>(a_hi, a_lo, row_off + (half * 4 + q * 2) * A_LBO, v);
      if (valid) {
        if (!LAST) st16_tbc(eff, tile, r, half * 4 + q * 2, v);
        if (RECORD) *reinterpret_cast<uint16_t*>(m_eff + (long long)row * 8 + half * 4 + q * 2) = (uint16_t)m;
      }
    }
    if (!LAST) {
      PILE_TRACE(2);
      run_gemm(c, [&](uint32_t el) {
        issue_gemm<128, 4, false>(el, c.tmem_d, c.a_hi, c.a_lo, c.aux_hi, c.aux_lo, c.zero, w_rs, w_rs + NB_WRS / 2);
      });
      PILE_TRACE(5);
      store_pr_ps(c, a_hi, t, r, half, PrOut, PsOut, (long long)tile * TILE, R, packed_ps);
    } else {
      // predictor: q = ReLU(V0 eff + c0);  s_pred = s_cur + V1 q + c1
      run_gemm(c, [&](uint32_t el) {
        issue_gemm<64, 4, true>(el, c.tmem_d, c.a_hi, c.a_lo, c.aux_hi, c.aux_lo, c.zero, w_v0, w_v0 + NB_K80 / 2);
      });
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float v[16];
        tc::tmem_ld16(c.taddr + half * 32 + q * 16, v);
        tc::tmem_ld_wait();
        const uint32_t m = relu_to_tile<RECORD>(a_hi, a_lo, row_off + (half * 4 + q * 2) * A_LBO, v);
        if (RECORD && valid) *reinterpret_cast<uint16_t*>(m_q + (long long)row * 8 + half * 4 + q * 2) = (uint16_t)m;
      }
      run_gemm(c, [&](uint32_t el) {
        issue_gemm<16, 4, true>(el, c.tmem_d, c.a_hi, c.a_lo, c.aux_hi, c.aux_lo, c.zero, w_v1, w_v1 + NB_V1 / 2);
      });
      if (half == 0) {
        float v[16];
        tc::tmem_ld16(c.taddr, v);
        tc::tmem_ld_wait();
        if (valid) {
          const int b = row / N, i = row - b * N;
          const float* sc = s_cur + (long long)b * s_stride + i * 3;
          float* so = s_out + (long long)b * o_stride + i * 3;
          so[0] = v[0] + sc[0];
          so[1] = v[1] + sc[1];
          so[2] = v[2] + sc[2];
        }
      }
    }
    PILE_TRACE(6);
  }
  PILE_TRACE(10);
  tile_epilogue(S);
}

PILE_TRACE_SETTER(set_node_trace)

template <typename Kern>
static int set_smem_tc(Kern k, size_t bytes) {
  return (int)cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

static int tc_grid(long long ntiles) {
  return (int)(ntiles < 1 ? 1 : (ntiles < NSM ? ntiles : NSM));
}

int launch_node_encode_tc(const float* wpack, const float* attr, const float* dens, const float* s_delta,
                          const Masks* mk, const StepScratch& ws, int B, int N, cudaStream_t st) {
  static DeviceOnce once;
  const int once_dev = once.pending();
  if (once_dev >= 0) {
    int e;
    if ((e = set_smem_tc(k_node_encode_tc<false>, sizeof(NodeEncSmemTc)))) return e;
    if ((e = set_smem_tc(k_node_encode_tc<true>, sizeof(NodeEncSmemTc)))) return e;
    if ((e = set_smem_tc(k_node_update_tc<false, false>, sizeof(NodeUpdSmemTc)))) return e;
    if ((e = set_smem_tc(k_node_update_tc<false, true>, sizeof(NodeUpdSmemTc)))) return e;
    if ((e = set_smem_tc(k_node_update_tc<true, false>, sizeof(NodeUpdSmemTc)))) return e;
    if ((e = set_smem_tc(k_node_update_tc<true, true>, sizeof(NodeUpdSmemTc)))) return e;
    if ((e = set_smem_tc(k_edge_agg<false, false>, agg_smem_bytes<false>(false)))) return e;
    if ((e = set_smem_tc(k_edge_agg<true, false>, agg_smem_bytes<false>(false)))) return e;
    if ((e = set_smem_tc(k_edge_agg<false, true>, agg_smem_bytes<true>(PS_PACK_OK)))) return e;
    if ((e = set_smem_tc(k_edge_agg<true, true>, agg_smem_bytes<true>(false)))) return e;
    once.done(once_dev);
  }
  const long long ntiles = ((long long)B * N + TILE - 1) / TILE;
  if (mk)
    k_node_encode_tc<true><<<tc_grid(ntiles), TC_THREADS, sizeof(NodeEncSmemTc), st>>>(
        wpack, attr, dens, s_delta, mk->pe0, mk->pe1, ws.Cp, ws.eff, ws.Pr[0], ws.Ps[0], B, N, PS_PACK_OK && g_use_tensor_cores == 2 && mk == nullptr);
  else
    k_node_encode_tc<false><<<tc_grid(ntiles), TC_THREADS, sizeof(NodeEncSmemTc), st>>>(
        wpack, attr, dens, s_delta, nullptr, nullptr, ws.Cp, ws.eff, ws.Pr[0], ws.Ps[0], B, N, PS_PACK_OK && g_use_tensor_cores == 2 && mk == nullptr);
  PILE_CHECK_LAUNCH();
  return 0;
}

int launch_propagate_tc(const float* wpack, const Csr& csr, const StepScratch& ws, const Masks* mk, int p,
                        const float* s_cur, long long s_stride, float* s_out, long long o_stride, int B, int N,
                        cudaStream_t st, cudaEvent_t mid) {
  const int in = p & 1, out = in ^ 1;
  const long long R = (long long)B * N;
  const long long ntiles = (R + TILE - 1) / TILE;
  if (R * 256 >= (1ll << 32)) return (int)cudaErrorInvalidValue;          // 32-bit byte offsets in k_edge_agg
  const int agg_blocks = (int)((R + 15) / 16 < 3 * NSM ? (R + 15) / 16 : 3 * NSM);      // 3 resident blocks per SM: one wave
  // tensor engine 2 (edge_tmem.cu) writes packed C_e rows, engine 1 (edge_tc.cu) plain fp32 rows
  const bool packed = g_use_tensor_cores == 2;
  uint8_t* me = mk ? mk->edge[p] : nullptr;
  auto agg_kernel = mk ? (packed ? k_edge_agg<true, true> : k_edge_agg<true, false>)
                       : (packed ? k_edge_agg<false, true> : k_edge_agg<false, false>);
  const int agg_smem = packed ? agg_smem_bytes<true>(PS_PACK_OK && mk == nullptr) : agg_smem_bytes<false>(false);
  agg_kernel<<<agg_blocks, AGG_THREADS, agg_smem, st>>>(csr.rowptr, csr.col, ws.Ce, ws.Pr[in], ws.Ps[in], me, ws.agg, B, N);
  PILE_CHECK_LAUNCH();
  if (mid) cudaEventRecord(mid, st);
  const int grid = tc_grid(ntiles);
  const size_t sm = sizeof(NodeUpdSmemTc);
  if (p < PSTEP - 1) {
    if (mk)
      k_node_update_tc<false, true><<<grid, TC_THREADS, sm, st>>>(wpack, ws.agg, ws.Cp, ws.eff, ws.Pr[out], ws.Ps[out],
                                                                  mk->eff[p], nullptr, s_cur, s_stride, s_out, o_stride, B, N, PS_PACK_OK && packed && mk == nullptr);
    else
      k_node_update_tc<false, false><<<grid, TC_THREADS, sm, st>>>(wpack, ws.agg, ws.Cp, ws.eff, ws.Pr[out], ws.Ps[out],
                                                                   nullptr, nullptr, s_cur, s_stride, s_out, o_stride, B, N, PS_PACK_OK && packed && mk == nullptr);
  } else {
    if (mk)
      k_node_update_tc<true, true><<<grid, TC_THREADS, sm, st>>>(wpack, ws.agg, ws.Cp, ws.eff, nullptr, nullptr, mk->eff[p],
                                                                 mk->q, s_cur, s_stride, s_out, o_stride, B, N, PS_PACK_OK && packed && mk == nullptr);
    else
      k_node_update_tc<true, false><<<grid, TC_THREADS, sm, st>>>(wpack, ws.agg, ws.Cp, ws.eff, nullptr, nullptr, nullptr,
                                                                  nullptr, s_cur, s_stride, s_out, o_stride, B, N, PS_PACK_OK && packed && mk == nullptr);
  }
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
