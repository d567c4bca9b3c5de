// Shared building blocks of the tcgen05 tile kernels (edge_tc.cu, node_tc.cu).
//
// Execution model: a persistent CTA (one per SM, 1024 threads) holds 4 independent GROUPS of 8 warps.
// A group owns one 128-row tile at a time: an activation tile pair (bf16 hi / lo, canonical K-major
// layout) in shared memory, a column range of TMEM for its accumulator and one mbarrier.  Its first warp
// issues the MMAs (descriptor arithmetic warp-uniform, tcgen05.mma predicated on the elected lane), all 8
// warps run the epilogues: thread (r, half) handles row r = TMEM lane and 32 of the accumulator columns
// (warps w and w+4 of a group share a TMEM lane quarter).  While one group waits for the tensor pipe the
// other groups run their epilogues.
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace pile {

// optional cycle trace of one warp (measurement hook, pile_debug_set_trace): every translation unit that
// includes this header gets its own copy and exposes a setter
static __device__ long long* g_trace = nullptr;
static __device__ int g_trace_cap = 0;
// The trace points cost a clock read + a predicated store each in every tile chain, so they are compiled in only
// with -DPILE_ENABLE_TRACE (tools/trace_edge_tc.py builds such a variant library); the setters always exist.
#ifdef PILE_ENABLE_TRACE
#define PILE_TRACE_DECL()                                                                                  \
  long long* trace = (blockIdx.x == 0 && threadIdx.x == 32 * 9) ? g_trace : nullptr; /* group 1, warp 1 */ \
  const int tr_cap = g_trace_cap;                                                                          \
  int tr_n = 0;
#define PILE_TRACE(tag_)                                                                                     \
  do {                                                                                                       \
    if (trace && tr_n < tr_cap) trace[tr_n++] = ((long long)(tag_) << 56) | (clock64() & 0x00ffffffffffffffLL); \
  } while (0)
#else
#define PILE_TRACE_DECL()
#define PILE_TRACE(tag_) do { } while (0)
#endif
#define PILE_TRACE_SETTER(name_)                                              \
  int name_(long long* buf, int cap) {                                        \
    cudaError_t e = cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf));           \
    if (e != cudaSuccess) return (int)e;                                      \
    return (int)cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap));           \
  }

constexpr int TC_GROUPS = 4;
constexpr int GROUP_THREADS = 256;
constexpr int TC_THREADS = TC_GROUPS * GROUP_THREADS;
constexpr uint32_t A_SBO = 128, A_LBO = (TILE / 8) * 128;   // 2048: one 16-byte K chunk of a 128-row tile
constexpr uint32_t A_BYTES = 8 * A_LBO;                     // 64 data columns: 16 KB per part
constexpr uint32_t B_SBO = 128;
__host__ __device__ constexpr uint32_t b_lbo(int n_rows) { return (uint32_t)(n_rows / 8) * 128; }
// bytes of one part (hi or lo) of a canonical [n_rows x k] bf16 weight image
__host__ __device__ constexpr uint32_t b_bytes(int n_rows, int k) { return (uint32_t)(k / 8) * b_lbo(n_rows); }

// Arrays that k_edge_agg touches (agg, P_r, P_s, C_e) stay row-major ([rows][64] fp32): a half-warp there streams
// whole 256-byte rows.  The tile kernels own one row per thread, so a direct access moves one 32-byte sector per
// lane and 32 different lines per instruction -- measured to be what bounds the particle kernels (the LSU/L1
// request rate, not HBM).  They therefore go through shared memory for these arrays (load_rows_to_tile,
// stage_put16 / stage_flush below: whole 128-byte lines per request).  (Making ALL arrays tile-blocked was
// measured too: it turns k_edge_agg's row reads into eight scattered sectors and costs more than it saves.)
__host__ __device__ __forceinline__ long long tb_off(long long tile, int r, int kc) {
  return (tile * TILE + r) * (long long)H + kc * 8;
}
__device__ __forceinline__ long long tb_row(long long row, int kc) { return row * (long long)H + kc * 8; }
// C_p and eff are touched by the tile kernels only (one row per thread), so they use a tile-blocked chunk-major
// layout [tile][chunk 0..7][row 0..127][8 floats]: the 32 lanes of a warp (32 consecutive rows, same chunk) move
// one contiguous 1 KB block per instruction.
__host__ __device__ __forceinline__ long long tbc_off(long long tile, int r, int kc) {
  return ((tile * 8 + kc) * TILE + r) * 8;
}
__device__ __forceinline__ void st8(float* __restrict__ p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld8(const float* __restrict__ p, float* v) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}

__device__ __forceinline__ void ld8_tbc(const float* __restrict__ base, long long tile, int r, int kc, float* v) {
  ld8(base + tbc_off(tile, r, kc), v);
}
__device__ __forceinline__ void st16_tbc(float* __restrict__ base, long long tile, int r, int kc, const float (&v)[16]) {
  st8(base + tbc_off(tile, r, kc), &v[0]);
  st8(base + tbc_off(tile, r, kc + 1), &v[8]);
}

struct GroupTile {          // per-group shared-memory operands
  alignas(128) uint8_t a[2][A_BYTES];     // [hi, lo] activation tile, 8 K chunks
  alignas(128) uint8_t aux[2][A_LBO];     // [hi, lo] K chunk (1, d, 0, ...): multiplies bias / density columns
};

__device__ __forceinline__ void group_barrier(int g) {
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GROUP_THREADS) : "memory");
}

// D[128 x N] (=|+=) A * B^T with bf16 hi/lo split operands (three passes): KSTEPS data K-steps of the group's
// A tile and, when AUX, one more K-step whose first chunk is the aux chunk and whose second chunk is the
// shared all-zero chunk.  B: canonical image of [N x 16*(KSTEPS+AUX)], hi at b_hi, lo at b_lo.
// To be executed by a CONVERGED warp; `elected` selects the issuing lane.
template <int N, int KSTEPS, bool AUX>
__device__ __forceinline__ void issue_gemm(uint32_t elected, uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo,
                                           uint32_t aux_hi, uint32_t aux_lo, uint32_t zero, uint32_t b_hi,
                                           uint32_t b_lo) {
  constexpr uint32_t idesc = tc::make_idesc_bf16(TILE, N);
  constexpr uint32_t BL = b_lbo(N);
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t a = pass == 1 ? a_lo : a_hi;
    const uint32_t x = pass == 1 ? aux_lo : aux_hi;
    const uint32_t b = pass == 2 ? b_lo : b_hi;
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k)
      tc::mma_bf16_if(elected, tmem_d, tc::make_desc(a + k * 2 * A_LBO, A_LBO, A_SBO),
                      tc::make_desc(b + k * 2 * BL, BL, B_SBO), idesc, (pass | k) != 0 ? 1u : 0u);
    if (AUX)
      tc::mma_bf16_if(elected, tmem_d, tc::make_desc(x, zero - x, A_SBO),
                      tc::make_desc(b + KSTEPS * 2 * BL, BL, B_SBO), idesc, (KSTEPS | pass) != 0 ? 1u : 0u);
  }
}

// input layer: a single K-step whose first chunk is chunk 0 of the A tile (features + constant 1) and whose
// second chunk is the zero chunk; B = canonical [N x 16]
template <int N>
__device__ __forceinline__ void issue_gemm_k16(uint32_t elected, uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo,
                                               uint32_t zero, uint32_t b_hi, uint32_t b_lo) {
  constexpr uint32_t idesc = tc::make_idesc_bf16(TILE, N);
  constexpr uint32_t BL = b_lbo(N);
  tc::mma_bf16_if(elected, tmem_d, tc::make_desc(a_hi, zero - a_hi, A_SBO), tc::make_desc(b_hi, BL, B_SBO), idesc, 0u);
  tc::mma_bf16_if(elected, tmem_d, tc::make_desc(a_lo, zero - a_lo, A_SBO), tc::make_desc(b_hi, BL, B_SBO), idesc, 1u);
  tc::mma_bf16_if(elected, tmem_d, tc::make_desc(a_hi, zero - a_hi, A_SBO), tc::make_desc(b_lo, BL, B_SBO), idesc, 1u);
}

// 8 consecutive fp32 values of a row -> one 16-byte K chunk of a hi and a lo tile (off = row + chunk offset)
__device__ __forceinline__ void store_chunk(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, const float (&v)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) tc::split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
  *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ReLU of 16 accumulator columns -> two K chunks of the A tile; returns the 16 sign bits
template <bool RECORD>
__device__ __forceinline__ uint32_t relu_to_tile(uint8_t* a_hi, uint8_t* a_lo, uint32_t off0, float (&v)[16]) {
  uint32_t m = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = v[h * 8 + j];
      if (RECORD) m |= x > 0.f ? (1u << (h * 8 + j)) : 0u;
      o[j] = fmaxf(x, 0.f);
      v[h * 8 + j] = o[j];
    }
    store_chunk(a_hi, a_lo, off0 + h * A_LBO, o);
  }
  return m;
}

// ---- row-major [128 x 64] fp32 tiles <-> one-row-per-thread registers, through the (dead) A tile --------
// The epilogue thread owns a whole row, a row-major global array wants whole 128-byte lines per request.  A
// 32 KB staging tile (the group's a[0..1], free once its last MMA has completed) transposes between the two:
// 16-byte slots, XOR-swizzled by the row so both sides are bank-conflict free.
__device__ __forceinline__ uint32_t stage_off(int r, int slot) { return (uint32_t)r * 256u + (uint32_t)((slot ^ (r & 7)) << 4); }
// this thread's 16 values of row r, columns c0 .. c0+15
__device__ __forceinline__ void stage_put16(uint8_t* stage, int r, int c0, const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(stage + stage_off(r, (c0 >> 2) + i)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// all 256 threads of the group: staging tile -> rows [row0, row0 + 128) of a row-major [R][64] array
__device__ __forceinline__ void stage_flush(const uint8_t* stage, int t, float* __restrict__ dst, long long row0, long long R) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int idx = it * GROUP_THREADS + t;
    const int row = idx >> 4, slot = idx & 15;
    const float4 v = *reinterpret_cast<const float4*>(stage + stage_off(row, slot));
    if (row0 + row < R) st4(dst + (row0 + row) * H + slot * 4, v);
  }
}
// ---- packed C_e rows (tensor engine 2) ------------------------------------------------------------------
// k_edge_agg re-reads C_e once per propagation step, so the tensor engine stores the rows as 24-bit words: per row
// 64 x 16-bit upper halves (128 bytes) followed by 64 x 8-bit third bytes (64 bytes) = 192 bytes instead of 256.
// The reader rebuilds a float with ONE byte permute per value and no mask: [upper half | b | b], i.e. the low 16
// mantissa bits become 257 * b.  The writer therefore stores b = round(low16 / 257) (multiply-high by 2^24 / 257):
// |257 b - low16| <= 128.5, a relative error <= 2^-16 like the bf16 hi/lo operands of the products that made the
// row -- the same bound as rounding to 24 bits, without the reader's AND.
constexpr int CE_PACKED_ROW = 192;
__device__ __forceinline__ uint32_t third_byte_scaled(float x) {          // byte 3 of the result = b
  return (__float_as_uint(x) & 0xffffu) * 65281u + 0x800000u;              // <= 65535 * 65281 + 2^23 < 2^32
}
__device__ __forceinline__ void pack24(const float4& v, uint2& hi, uint32_t& lo) {
  hi.x = __byte_perm(__float_as_uint(v.x), __float_as_uint(v.y), 0x7632);                  // [x.2 x.3 y.2 y.3]
  hi.y = __byte_perm(__float_as_uint(v.z), __float_as_uint(v.w), 0x7632);
  lo = __byte_perm(__byte_perm(third_byte_scaled(v.x), third_byte_scaled(v.y), 0x0073),
                   __byte_perm(third_byte_scaled(v.z), third_byte_scaled(v.w), 0x0073), 0x5410);   // [bx by bz bw]
}
__device__ __forceinline__ float4 unpack24(const uint2& hi, uint32_t lo) {
  float4 v;
  v.x = __uint_as_float(__byte_perm(hi.x, lo, 0x1044));       // [lo.0 lo.0 hi.0 hi.1]
  v.y = __uint_as_float(__byte_perm(hi.x, lo, 0x3255));
  v.z = __uint_as_float(__byte_perm(hi.y, lo, 0x1066));
  v.w = __uint_as_float(__byte_perm(hi.y, lo, 0x3277));
  return v;
}
// all 256 threads of the group: staging tile -> packed rows [row0, row_end) (row_end - row0 <= 128)
__device__ __forceinline__ void stage_flush_packed(const uint8_t* stage, int t, uint8_t* __restrict__ dst, long long row0,
                                                   long long row_end) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int idx = it * GROUP_THREADS + t;
    const int row = idx >> 4, slot = idx & 15;
    const float4 v = *reinterpret_cast<const float4*>(stage + stage_off(row, slot));
    uint2 hi;
    uint32_t lo;
    pack24(v, hi, lo);
    if (row0 + row < row_end) {
      uint8_t* p = dst + (row0 + row) * CE_PACKED_ROW;
      *reinterpret_cast<uint2*>(p + slot * 8) = hi;
      *reinterpret_cast<uint32_t*>(p + 128 + slot * 4) = lo;
    }
  }
}

// all 256 threads of the group: rows [row0, row0+128) of a row-major [R][64] array -> hi/lo A tile (rows past R
// are zero-filled).  A warp instruction covers 8 rows x 128 bytes (whole lines) and 8 consecutive rows of one
// chunk per quarter-warp on the shared-memory side (conflict-free).
__device__ __forceinline__ void load_rows_to_tile(const float* __restrict__ src, long long row0, long long R, int t,
                                                  uint8_t* a_hi, uint8_t* a_lo) {
  const int w = t >> 5, lane = t & 31;
  float o[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = w * 16 + (j >> 1) * 8 + (lane & 7), kc = (j & 1) * 4 + (lane >> 3);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[j][i] = 0.f;
    if (row0 + row < R) ld8(src + (row0 + row) * H + kc * 8, o[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = w * 16 + (j >> 1) * 8 + (lane & 7), kc = (j & 1) * 4 + (lane >> 3);
    store_chunk(a_hi, a_lo, (row >> 3) * A_SBO + (row & 7) * 16 + kc * A_LBO, o[j]);
  }
}

// everything a group needs to hand a layer to the tensor cores and wait for it
struct GroupCtx {
  int g, wig;
  uint32_t tmem_d;          // accumulator base column of the group
  uint32_t taddr;           // + lane-quarter offset of this warp
  uint32_t a_hi, a_lo, aux_hi, aux_lo, zero;
  uint64_t* bar;
  uint32_t phase;
};

// call with the A tile (and aux chunk) written by this group's threads; `issue(elected)` enqueues the MMAs
template <typename Issue>
__device__ __forceinline__ void run_gemm(GroupCtx& c, Issue issue) {
  tc::fence_async_smem();          // st.shared of the A tile -> visible to the tensor core
  tc::fence_before_sync();         // this thread's tcgen05.ld of the previous accumulator are complete
  group_barrier(c.g);
  if (c.wig == c.g) {      // issuer warp g*8+g: the four groups issue from four different SM sub-partitions
    tc::fence_after_sync();
    const uint32_t elected = tc::elect_one();
    issue(elected);
    if (elected) tc::mma_commit(c.bar);
    __syncwarp();
  }
  tc::mbar_wait(c.bar, c.phase);
  c.phase ^= 1;
  tc::fence_after_sync();
}

}  // namespace pile
