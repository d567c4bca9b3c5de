// K1/K2: pusher influence (s_delta) and the radius ^ top-10 neighbour search.
//
// Replaces the dense O(N^2) tensors of reference model/gnn_dyn.py:221-251 (repeat -> dis -> topk ->
// scatter -> nonzero -> one-hot Rr/Rs) by one CTA per sample that keeps the pushed positions in
// shared memory and emits a compact int32 CSR (+COO receiver ids, + optional sender-major transpose
// for the deterministic backward scatter).  The relation SET is bit-exact w.r.t. the reference:
// distances use separately rounded fp32 mul/add in the reference's x,y,z order on
// p = fl(s_cur + s_delta); ties at the 10th place resolve to the lower sender index.
#include "common.cuh"
#include "kernels.h"

namespace pile {

struct PushFrame {
  float sx, sy, sz, ex, ey, ez, ux, uy, uz, len;
};

__device__ __forceinline__ void cam_point(const PushCam& c, float wx, float wy, float wz, float& x, float& y,
                                          float& z) {
  x = __fdiv_rn(c.m[0] * wx + c.m[1] * wy + c.m[2] * wz + c.m[3], c.global_scale);
  y = __fdiv_rn(c.m[4] * wx + c.m[5] * wy + c.m[6] * wz + c.m[7], c.global_scale);
  z = __fdiv_rn(c.m[8] * wx + c.m[9] * wy + c.m[10] * wz + c.m[11], c.global_scale);
}

// planners.py:218-240 (simulator) / :274-285 (real robot): action (sx,sy,ex,ey) -> start/end of the push in the
// frame the particles live in, unit push direction
__device__ __forceinline__ PushFrame make_push_frame(const PushCam& c, const float* __restrict__ act) {
  PushFrame f;
  if (c.kind == 0) {
    cam_point(c, act[0], 0.f, -act[1], f.sx, f.sy, f.sz);
    cam_point(c, act[2], 0.f, -act[3], f.ex, f.ey, f.ez);
  } else {
    f.sx = __fdiv_rn(act[0], c.s2r_scale);
    f.sy = -__fdiv_rn(act[1], c.s2r_scale);
    f.ex = __fdiv_rn(act[2], c.s2r_scale);
    f.ey = -__fdiv_rn(act[3], c.s2r_scale);
    f.sz = c.height;
    f.ez = c.height;
  }
  const float dx = __fsub_rn(f.ex, f.sx), dy = __fsub_rn(f.ey, f.sy), dz = __fsub_rn(f.ez, f.sz);
  f.len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  f.ux = __fdiv_rn(dx, f.len);
  f.uy = __fdiv_rn(dy, f.len);
  f.uz = __fdiv_rn(dz, f.len);
  return f;
}

__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

// planners.py:241-254 (:286-297 for the real robot, whose particles are first shifted by the workspace centre;
// shift = 0 in the simulator frame and x - 0 is exact) for one particle
__device__ __forceinline__ void push_delta(const PushCam& c, const PushFrame& f, float x, float y, float z,
                                           float& ox, float& oy, float& oz) {
  x = __fsub_rn(x, c.shift_x);
  y = __fsub_rn(y, c.shift_y);
  const float rx = __fsub_rn(x, f.sx), ry = __fsub_rn(y, f.sy), rz = __fsub_rn(z, f.sz);
  const float across = dot3_rn(rx, ry, rz, -f.uy, f.ux, 0.f);
  const float along = dot3_rn(rx, ry, rz, f.ux, f.uy, f.uz);
  const float hard = (along < f.len && along > 0.f) ? 1.f : 0.f;
  const float excess = fmaxf(fmaxf(__fsub_rn(-c.pusher_w, across), 0.f), fmaxf(__fsub_rn(across, c.pusher_w), 0.f));
  const float soft = expf(__fdiv_rn(-excess, c.decay));
  const float to_end = dot3_rn(__fsub_rn(f.ex, x), __fsub_rn(f.ey, y), __fsub_rn(f.ez, z), f.ux, f.uy, f.uz);
  ox = __fmul_rn(__fmul_rn(__fmul_rn(to_end, f.ux), hard), soft);
  oy = __fmul_rn(__fmul_rn(__fmul_rn(to_end, f.uy), hard), soft);
  oz = __fmul_rn(__fmul_rn(__fmul_rn(to_end, f.uz), hard), soft);
}

__global__ void k_gen_s_delta(const float* __restrict__ s_cur, long long s_stride,
                              const float* __restrict__ action, int act_stride,
                              PushCam cam, int N, float* __restrict__ s_delta) {
  const int b = blockIdx.x;
  const PushFrame f = make_push_frame(cam, action + (size_t)b * act_stride);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float* p = s_cur + (long long)b * s_stride + i * 3;
    float ox, oy, oz;
    push_delta(cam, f, p[0], p[1], p[2], ox, oy, oz);
    float* o = s_delta + ((size_t)b * N + i) * 3;
    o[0] = ox; o[1] = oy; o[2] = oz;
  }
}

// dis[i][j] exactly as torch evaluates sum((p_j - p_i)^2, -1)   (gnn_dyn.py:224-230)
__device__ __forceinline__ float sqdist_rn(float xi, float yi, float zi, float xj, float yj, float zj) {
  const float dx = __fsub_rn(xj, xi), dy = __fsub_rn(yj, yi), dz = __fsub_rn(zj, zi);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

#define PILE_CE(a, b)                                   \
  {                                                     \
    const int lo__ = min(id[a], id[b]);                 \
    const int hi__ = max(id[a], id[b]);                 \
    id[a] = lo__;                                       \
    id[b] = hi__;                                       \
  }

// 29-comparator sorting network for 10 keys (verified with the 0-1 principle in tests/test_host_logic.py)
__device__ __forceinline__ void sort10(int (&id)[KMAX]) {
  PILE_CE(0, 8) PILE_CE(1, 9) PILE_CE(2, 7) PILE_CE(3, 5) PILE_CE(4, 6)
  PILE_CE(0, 2) PILE_CE(1, 4) PILE_CE(5, 8) PILE_CE(7, 9)
  PILE_CE(0, 3) PILE_CE(2, 4) PILE_CE(5, 7) PILE_CE(6, 9)
  PILE_CE(0, 1) PILE_CE(3, 6) PILE_CE(8, 9)
  PILE_CE(1, 5) PILE_CE(2, 3) PILE_CE(4, 8) PILE_CE(6, 7)
  PILE_CE(1, 2) PILE_CE(3, 5) PILE_CE(4, 6) PILE_CE(7, 8)
  PILE_CE(2, 3) PILE_CE(4, 5) PILE_CE(6, 7)
  PILE_CE(3, 4) PILE_CE(5, 6)
}

constexpr int NBR_THREADS = 1024;   // upper bound; the search kernel runs round_up(N, 32) threads so one pass covers all receivers

// exclusive scan of one int per thread over the CTA; returns the exclusive prefix, total via *total
__device__ __forceinline__ int block_exscan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int nwarps = (int)(blockDim.x >> 5);
    int w = lane < nwarps ? warp_sums[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < nwarps) warp_sums[lane] = winc - w;
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  return warp_sums[warp] + inc - v;
}

// packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2, two IEEE round-to-nearest results per instruction).  Only the
// subtraction and the multiplication are packed: ptxas contracts a packed multiply feeding a packed add into FFMA2
// even with explicit .rn (seen in SASS), which would break bit-exactness against torch's separately rounded
// (d*d).sum(-1); with scalar __fadd_rn it does not.
__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t sub2_rn(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2_rn(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

constexpr int NBR_BLOCK = 32;     // candidates per admission mask

// One CTA per sample.  Optional fused s_delta (action != nullptr).
// dynamic smem: pos[N] (float4) | cutd[N] | cuti[N] | deg[N] | roff[N+1] | sel[N*KMAX] | perm[N] | xs[NP] | ys[NP] | zs[NP] |
// cur[N] (float4: un-pushed position + attribute, for the relation-encoder input rows) | transposed lists (tape runs)
// (NP = N rounded up to a multiple of 32; the tail of xs holds +inf so that padded candidates are never admitted)
__global__ void __launch_bounds__(NBR_THREADS)
k_nbr_search(const float* __restrict__ s_cur, long long s_stride, const float* __restrict__ s_delta_in,
             const float* __restrict__ action, int act_stride, PushCam cam, float* __restrict__ s_delta_out,
             const int* __restrict__ particle_nums, int N, float thr,
             int* __restrict__ rowptr, int* __restrict__ col, int* __restrict__ row,
             int* __restrict__ trowptr, int* __restrict__ trecv, int* __restrict__ tedge,
             const float* __restrict__ attr, const float* __restrict__ dens, float* __restrict__ efeat) {
  extern __shared__ __align__(16) float smem[];
  float4* pos = reinterpret_cast<float4*>(smem);      // pushed positions (x, y, z, 0): one 16-byte load per pair
  float* cutd = smem + 4 * N;
  int* cuti = reinterpret_cast<int*>(cutd + N);
  int* deg = cuti + N;
  int* roff = deg + N;           // N+1
  int* sel = roff + (N + 1);     // N*KMAX
  int* perm = sel + N * KMAX;    // N: receivers in coarse spatial order (which lane handles which receiver)
  const int NP = (N + NBR_BLOCK - 1) / NBR_BLOCK * NBR_BLOCK;
  // SoA copy of the positions, 16-byte aligned (the pair loads of phase A are merged into 128-bit loads)
  float* xs = smem + (19 * N + 1 + 3) / 4 * 4;
  float* ys = xs + NP;
  float* zs = ys + NP;
  float4* cur = reinterpret_cast<float4*>(zs + NP);      // (s_cur, attr): the epilogue reads no global memory
  __shared__ int warp_sums[NBR_THREADS / 32];
  __shared__ int total_s;
  __shared__ float box[6];
  __shared__ int cell_cnt[65];

  const int b = blockIdx.x;
  const int nvalid = particle_nums ? min(particle_nums[b], N) : N;
  const size_t base = (size_t)b * N;

  PushFrame f;
  if (action) f = make_push_frame(cam, action + (size_t)b * act_stride);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float* p = s_cur + (long long)b * s_stride + i * 3;
    const float x = p[0], y = p[1], z = p[2];
    float dx, dy, dz;
    if (action) {
      push_delta(cam, f, x, y, z, dx, dy, dz);
      float* o = s_delta_out + (base + i) * 3;
      o[0] = dx; o[1] = dy; o[2] = dz;
    } else {
      const float* d = s_delta_in + (base + i) * 3;
      dx = d[0]; dy = d[1]; dz = d[2];
    }
    const float px = __fadd_rn(x, dx), py = __fadd_rn(y, dy), pz = __fadd_rn(z, dz);
    pos[i] = make_float4(px, py, pz, 0.f);
    xs[i] = px; ys[i] = py; zs[i] = pz;
    cur[i] = make_float4(x, y, z, efeat != nullptr ? attr[base + i] : 0.f);
  }
  for (int i = N + threadIdx.x; i < NP; i += blockDim.x) { xs[i] = __int_as_float(0x7f800000); ys[i] = 0.f; zs[i] = 0.f; }
  __syncthreads();

  // Which lane handles which receiver does not change any result (every receiver scans all candidates in ascending
  // index), but it decides how often a warp runs the divergent park/insert code: with 32 NEARBY receivers per warp
  // a candidate is either interesting to many lanes at once or to none.  So receivers are binned into an 8 x 8 grid
  // over the two widest axes of the sample's bounding box and handed out in Morton order of the cells (counting
  // sort with shared-memory atomics; the order inside a cell is arbitrary and irrelevant).
  {
    float lo[3] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000)};
    float hi[3] = {__int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000)};
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const float4 p = pos[i];
      lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
      hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
        hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
      }
    float* wl = cutd;            // scratch (cutd .. sel are written later): [warps][3] minima, then [warps][3] maxima
    const int nw = (int)(blockDim.x >> 5);
    if ((threadIdx.x & 31) == 0)
      for (int a = 0; a < 3; ++a) { wl[(threadIdx.x >> 5) * 3 + a] = lo[a]; wl[(nw + (threadIdx.x >> 5)) * 3 + a] = hi[a]; }
    if (threadIdx.x < 65) cell_cnt[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x < 3) {
      float l = wl[threadIdx.x], h = wl[nw * 3 + threadIdx.x];
      for (int w = 1; w < nw; ++w) { l = fminf(l, wl[w * 3 + threadIdx.x]); h = fmaxf(h, wl[(nw + w) * 3 + threadIdx.x]); }
      box[threadIdx.x] = l;
      box[3 + threadIdx.x] = h - l;
    }
    __syncthreads();
    // the two widest axes
    const float ex = box[3], ey = box[4], ez = box[5];
    const int drop = (ex <= ey && ex <= ez) ? 0 : ((ey <= ez) ? 1 : 2);
    const int a0 = drop == 0 ? 1 : 0, a1 = drop == 2 ? 1 : 2;
    const float s0 = box[3 + a0] > 0.f ? 8.f / box[3 + a0] : 0.f, s1 = box[3 + a1] > 0.f ? 8.f / box[3 + a1] : 0.f;
    auto cell_of = [&](int i) {
      const float4 p = pos[i];
      const float c[3] = {p.x, p.y, p.z};
      const unsigned q0 = (unsigned)min(7, max(0, (int)((c[a0] - box[a0]) * s0)));
      const unsigned q1 = (unsigned)min(7, max(0, (int)((c[a1] - box[a1]) * s1)));
      return (int)((q0 & 1u) | ((q1 & 1u) << 1) | ((q0 & 2u) << 1) | ((q1 & 2u) << 2) | ((q0 & 4u) << 2) | ((q1 & 4u) << 3));
    };
    for (int i = threadIdx.x; i < N; i += blockDim.x) atomicAdd(&cell_cnt[cell_of(i)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      int run = 0;
      for (int cidx = 0; cidx < 64; ++cidx) { const int n = cell_cnt[cidx]; cell_cnt[cidx] = run; run += n; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) perm[atomicAdd(&cell_cnt[cell_of(i)], 1)] = i;
    __syncthreads();
  }

  // pass 1: per receiver (one thread each), the (up to) 10 nearest in-radius senders, then sort them by index.
  // The sorted insertion is ~45 instructions and diverges (almost every j makes SOME lane of the warp insert),
  // so candidates are parked in a 3-deep per-lane FIFO and the warp runs the insertion code only when a lane's
  // FIFO is full: ~5x fewer executions on a 300-particle pile.  The pre-filter then uses a slightly stale
  // 10th-best distance, which only lets a few extra candidates through; the insertion itself re-checks.
  // pass 1: per receiver (one thread each), the (up to) 10 nearest in-radius senders in (distance, index) order, then
  // sorted by index.  Candidates are taken 32 at a time in two phases:
  //   A  distances of the 32 candidates (packed fp32x2 subtract / multiply, no branches, no cross-lane traffic) and
  //      a 32-bit mask of those below the lane's admission bound min(radius^2, current 10th best);
  //   B  while any lane of the warp has mask bits left, every such lane pops its lowest one and runs the sorted
  //      insertion (~50 instructions, divergent) on it.
  // The bound is up to 32 candidates stale in phase A, which only lets a few extra candidates through; phase B
  // re-checks.  A lane visits its candidates in ascending index, so equal distances keep the lower index first.
  for (int base_i = 0; base_i < N; base_i += blockDim.x) {     // every thread takes part in the warp votes
    const bool active = base_i + (int)threadIdx.x < N;
    const int i = active ? perm[base_i + threadIdx.x] : N;
    float bd[KMAX];
    int id[KMAX];
#pragma unroll
    for (int s = 0; s < KMAX; ++s) { bd[s] = __int_as_float(0x7f800000); id[s] = 0x7fffffff; }
    const int ic = active ? i : 0;
    const float4 pi = pos[ic];
    const float xi = pi.x, yi = pi.y, zi = pi.z;
    const uint64_t X2 = pack2(xi, xi), Y2 = pack2(yi, yi), Z2 = pack2(zi, zi);
    // Warm start of the admission bound from whatever the output lists hold on entry -- in a rollout the relations of
    // the step before, in the planner's loop those of the iteration before: ANY ten distinct candidates are at most
    // D = max of their current distances away, so the tenth smallest distance is <= D and no candidate with d > D can be
    // selected.  Candidates with d == D still can (ties go to the lower index), hence the bound is the next float above
    // D.  The selected set is unchanged; the sorted insertion runs ~10 + few times per receiver instead of
    // ~10 (1 + ln(in-radius / 10)) times.  The old list is validated (ten strictly ascending indices in range), so
    // stale or uninitialised memory is harmless; a sample's old lists are read here, its new ones written after the
    // barrier below by the same CTA.
    float lim0 = thr;
    if (active) {
      const int* prp = rowptr + (size_t)b * (N + 1) + i;
      const int plo = prp[0];
      if (prp[1] - plo == KMAX && plo >= 0 && plo <= KMAX * N - KMAX) {
        const int* pc = col + (size_t)b * KMAX * N + plo;
        int c[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) c[k] = pc[k];
        bool ok = c[0] >= 0 && c[KMAX - 1] < N;
#pragma unroll
        for (int k = 1; k < KMAX; ++k) ok = ok && c[k - 1] < c[k];
        if (ok) {
          float D = 0.f;
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            const float4 pj = pos[c[k]];
            D = fmaxf(D, sqdist_rn(xi, yi, zi, pj.x, pj.y, pj.z));
          }
          lim0 = fminf(thr, __int_as_float(__float_as_int(D) + 1));
        }
      }
    }
    // admission bound; -1 keeps lanes without a receiver out (d >= 0)
    float lim = active ? lim0 : -1.f;
    for (int jb = 0; jb < NP; jb += NBR_BLOCK) {
      unsigned m = 0;
#pragma unroll
      for (int q = 0; q < NBR_BLOCK / 2; ++q) {
        const uint64_t dx = sub2_rn(*reinterpret_cast<const uint64_t*>(xs + jb + 2 * q), X2);
        const uint64_t dy = sub2_rn(*reinterpret_cast<const uint64_t*>(ys + jb + 2 * q), Y2);
        const uint64_t dz = sub2_rn(*reinterpret_cast<const uint64_t*>(zs + jb + 2 * q), Z2);
        float a0, a1, b0, b1, c0, c1;
        unpack2(mul2_rn(dx, dx), a0, a1);
        unpack2(mul2_rn(dy, dy), b0, b1);
        unpack2(mul2_rn(dz, dz), c0, c1);
        const float d0 = __fadd_rn(__fadd_rn(a0, b0), c0), d1 = __fadd_rn(__fadd_rn(a1, b1), c1);
        m |= (d0 < lim ? 1u : 0u) << (2 * q);
        m |= (d1 < lim ? 1u : 0u) << (2 * q + 1);
      }
      while (__any_sync(0xffffffffu, m != 0u)) {
        if (m != 0u) {
          const int j = jb + __ffs((int)m) - 1;
          m &= m - 1u;
          const float4 pj = pos[j];
          const float d = sqdist_rn(xi, yi, zi, pj.x, pj.y, pj.z);
          if (d < lim) {
#pragma unroll
            for (int s = KMAX - 1; s > 0; --s) {
              const bool shift = d < bd[s - 1];
              const bool here = !shift && d < bd[s];
              const float nd = shift ? bd[s - 1] : (here ? d : bd[s]);
              const int ni = shift ? id[s - 1] : (here ? j : id[s]);
              bd[s] = nd; id[s] = ni;
            }
            if (d < bd[0]) { bd[0] = d; id[0] = j; }
            lim = fminf(lim0, bd[KMAX - 1]);
          }
        }
      }
    }
    if (active) {
      int n = 0;
      if (i < nvalid) {
#pragma unroll
        for (int s = 0; s < KMAX; ++s) {
          if (id[s] >= nvalid) id[s] = 0x7fffffff;    // padded particles are dropped AFTER top-k (gnn_dyn.py:238-241)
          n += id[s] != 0x7fffffff;
        }
        sort10(id);
      }
#pragma unroll
      for (int s = 0; s < KMAX; ++s) sel[i * KMAX + s] = (i < nvalid) ? id[s] : 0x7fffffff;
      deg[i] = n;
    }
  }
  __syncthreads();

  // CSR offsets
  const int per = (N + (int)blockDim.x - 1) / (int)blockDim.x;
  {
    const int lo = min(threadIdx.x * per, N), hi = min(lo + per, N);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += deg[i];
    int run = block_exscan(s, warp_sums, &total_s);
    for (int i = lo; i < hi; ++i) { roff[i] = run; run += deg[i]; }
    if (threadIdx.x == 0) roff[N] = total_s;
  }
  __syncthreads();
  int* rp = rowptr + (size_t)b * (N + 1);
  for (int i = threadIdx.x; i <= N; i += blockDim.x) rp[i] = roff[i];
  const size_t ebase = (size_t)b * KMAX * N;
  // one (receiver, slot) pair per thread: consecutive threads write consecutive relations
  const float dn = efeat != nullptr ? dens[b] / 5000.f : 0.f;
  for (int idx = threadIdx.x; idx < N * KMAX; idx += blockDim.x) {
    const int i = idx / KMAX, s = idx - i * KMAX;
    if (s >= deg[i]) continue;
    const int c = sel[idx];
    const size_t e = ebase + roff[i] + s;
    col[e] = c;
    row[e] = i;
    // optional: the relation encoder's input rows (attr_r, attr_s, s_cur_r - s_cur_s, density; gnn_dyn.py:164-172)
    // for the tensor engine, the same values k_edge_features (edge_tc.cu) writes
    if (efeat != nullptr) {
      const float4 pr = cur[i], ps = cur[c];
      float* out = efeat + e * 8;
      *reinterpret_cast<float4*>(out) = make_float4(pr.w, ps.w, pr.x - ps.x, pr.y - ps.y);
      *reinterpret_cast<float4*>(out + 4) = make_float4(pr.z - ps.z, dn, 0.f, 0.f);
    }
  }

  if (trowptr == nullptr) return;
  // sender-major transpose: for sender j the receivers i (ascending) with j in nbr(i), and the edge id.  Built from
  // the receiver lists by a counting sort in shared memory (O(E) shared-memory atomics) instead of a second O(N^2)
  // distance scan per sender; the atomics fill a sender's segment in arbitrary order, so every segment is then sorted
  // by receiver (each receiver appears at most once per sender): the result is deterministic.
  int* tre = reinterpret_cast<int*>(cur + N);      // [KMAX*N] receiver of transposed slot (tape runs only)
  int* ted = tre + KMAX * N;                       // [KMAX*N] edge id
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += blockDim.x) deg[j] = 0;       // reuse as in-degree / fill cursor
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * KMAX; idx += blockDim.x) {
    const int c = sel[idx];
    if (c != 0x7fffffff) atomicAdd(&deg[c], 1);
  }
  __syncthreads();
  {
    const int lo = min(threadIdx.x * per, N), hi = min(lo + per, N);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += deg[i];
    int run = block_exscan(s, warp_sums, &total_s);
    int* trp = trowptr + (size_t)b * (N + 1);
    // cuti is free now (the search is over): keep the segment starts there, deg becomes the running fill cursor
    for (int i = lo; i < hi; ++i) { const int d = deg[i]; trp[i] = run; cuti[i] = run; deg[i] = run; run += d; }
    if (threadIdx.x == 0) trp[N] = total_s;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * KMAX; idx += blockDim.x) {
    const int c = sel[idx];
    if (c != 0x7fffffff) {
      const int i = idx / KMAX, s = idx - i * KMAX;
      const int p = atomicAdd(&deg[c], 1);
      tre[p] = i;
      ted[p] = roff[i] + s;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += blockDim.x) {       // insertion sort of sender j's segment by receiver
    const int lo = cuti[j], hi = deg[j];
    for (int a = lo + 1; a < hi; ++a) {
      const int r = tre[a], e = ted[a];
      int q = a - 1;
      while (q >= lo && tre[q] > r) { tre[q + 1] = tre[q]; ted[q + 1] = ted[q]; --q; }
      tre[q + 1] = r;
      ted[q + 1] = e;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < total_s; idx += blockDim.x) {
    trecv[ebase + idx] = tre[idx];
    tedge[ebase + idx] = ted[idx];
  }
}

// ------------------------------------------------------------------------------------------------
// backward of the pusher model: g_s_delta [B,N,3] -> g_s_cur (+=) and g_action [B,4].
// The hard along-push mask and the push length inside it carry no gradient (planners.py:248).
// ------------------------------------------------------------------------------------------------
constexpr int SDB_THREADS = 256;
__global__ void __launch_bounds__(SDB_THREADS)
k_gen_s_delta_bwd(const float* __restrict__ s_cur, long long s_stride, const float* __restrict__ action,
                  int act_stride, PushCam cam, int N, const float* __restrict__ g_sd, float* __restrict__ g_s_cur,
                  long long g_stride, float* __restrict__ g_action, int g_act_stride) {
  __shared__ float red[9][SDB_THREADS / 32];
  const int b = blockIdx.x;
  const PushFrame f = make_push_frame(cam, action + (size_t)b * act_stride);
  float gu[3] = {0.f, 0.f, 0.f}, ge[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float* p = s_cur + (long long)b * s_stride + i * 3;
    const float x = p[0] - cam.shift_x, y = p[1] - cam.shift_y, z = p[2];
    const float* g = g_sd + ((size_t)b * N + i) * 3;
    const float g0 = g[0], g1 = g[1], g2 = g[2];
    const float rx = x - f.sx, ry = y - f.sy, rz = z - f.sz;
    const float vx = -f.uy, vy = f.ux;
    const float across = rx * vx + ry * vy;
    const float along = rx * f.ux + ry * f.uy + rz * f.uz;
    const float hard = (along < f.len && along > 0.f) ? 1.f : 0.f;
    const float excess = fmaxf(fmaxf(-cam.pusher_w - across, 0.f), fmaxf(across - cam.pusher_w, 0.f));
    const float soft = expf(-excess / cam.decay);
    const float ex = f.ex - x, ey = f.ey - y, ez = f.ez - z;
    const float to_end = ex * f.ux + ey * f.uy + ez * f.uz;
    const float c = hard * soft;
    const float gdotu = g0 * f.ux + g1 * f.uy + g2 * f.uz;
    const float g_te = c * gdotu;
    const float g_soft = to_end * gdotu * hard;
    const float g_excess = -g_soft * soft / cam.decay;
    const float g_across = g_excess * (across > cam.pusher_w ? 1.f : (across < -cam.pusher_w ? -1.f : 0.f));
    // out = to_end * c * u
    gu[0] += to_end * c * g0 + g_te * ex;
    gu[1] += to_end * c * g1 + g_te * ey;
    gu[2] += to_end * c * g2 + g_te * ez;
    // v = (-u.y, u.x, 0): across = rel . v
    gu[1] -= g_across * rx;
    gu[0] += g_across * ry;
    ge[0] += g_te * f.ux; ge[1] += g_te * f.uy; ge[2] += g_te * f.uz;
    const float grx = g_across * vx, gry = g_across * vy;
    gs[0] -= grx; gs[1] -= gry;
    float* o = g_s_cur + (long long)b * g_stride + i * 3;
    o[0] += grx - g_te * f.ux;
    o[1] += gry - g_te * f.uy;
    o[2] += -g_te * f.uz;
  }
  float v[9] = {gu[0], gu[1], gu[2], ge[0], ge[1], ge[2], gs[0], gs[1], gs[2]};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) red[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t[9];
    for (int k = 0; k < 9; ++k) {
      t[k] = 0.f;
      for (int w = 0; w < SDB_THREADS / 32; ++w) t[k] += red[k][w];
    }
    // u = p / |p|
    const float udg = f.ux * t[0] + f.uy * t[1] + f.uz * t[2];
    const float gp[3] = {(t[0] - f.ux * udg) / f.len, (t[1] - f.uy * udg) / f.len, (t[2] - f.uz * udg) / f.len};
    const float gE[3] = {t[3] + gp[0], t[4] + gp[1], t[5] + gp[2]};
    const float gS[3] = {t[6] - gp[0], t[7] - gp[1], t[8] - gp[2]};
    float* ga = g_action + (size_t)b * g_act_stride;
    if (cam.kind == 0) {
      // start = (M[:,0] a0 - M[:,2] a1 + M[:,3]) / gs ; end likewise with a2, a3
      const float inv = 1.f / cam.global_scale;
      ga[0] = (gS[0] * cam.m[0] + gS[1] * cam.m[4] + gS[2] * cam.m[8]) * inv;
      ga[1] = -(gS[0] * cam.m[2] + gS[1] * cam.m[6] + gS[2] * cam.m[10]) * inv;
      ga[2] = (gE[0] * cam.m[0] + gE[1] * cam.m[4] + gE[2] * cam.m[8]) * inv;
      ga[3] = -(gE[0] * cam.m[2] + gE[1] * cam.m[6] + gE[2] * cam.m[10]) * inv;
    } else {
      // start = (a0 / s2r, -a1 / s2r, h) ; end likewise with a2, a3
      const float inv = 1.f / cam.s2r_scale;
      ga[0] = gS[0] * inv;
      ga[1] = -gS[1] * inv;
      ga[2] = gE[0] * inv;
      ga[3] = -gE[1] * inv;
    }
  }
}

int launch_gen_s_delta_bwd(const float* s_cur, long long s_stride, const float* action, int act_stride,
                           const PushCam& cam, int B, int N, const float* g_sd, float* g_s_cur, long long g_stride,
                           float* g_action, int g_act_stride, cudaStream_t st) {
  k_gen_s_delta_bwd<<<B, SDB_THREADS, 0, st>>>(s_cur, s_stride, action, act_stride, cam, N, g_sd, g_s_cur, g_stride,
                                               g_action, g_act_stride);
  PILE_CHECK_LAUNCH();
  return 0;
}

// sender-major transpose of caller-provided relation lists (dense Rr/Rs shim path): one CTA per sample,
// thread per sender, edges visited in ascending id so the result is deterministic
__global__ void k_transpose_relations(const int* __restrict__ rowptr, const int* __restrict__ col,
                                      const int* __restrict__ row, int N, int* __restrict__ trowptr,
                                      int* __restrict__ trecv, int* __restrict__ tedge) {
  extern __shared__ int cnt[];    // [N+1]
  const int b = blockIdx.x;
  const int ne = rowptr[(size_t)b * (N + 1) + N];
  const size_t ebase = (size_t)b * KMAX * N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    int n = 0;
    for (int e = 0; e < ne; ++e) n += col[ebase + e] == j;
    cnt[j] = n;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int j = 0; j < N; ++j) { const int n = cnt[j]; cnt[j] = run; run += n; }
    cnt[N] = run;
  }
  __syncthreads();
  for (int j = threadIdx.x; j <= N; j += blockDim.x) trowptr[(size_t)b * (N + 1) + j] = cnt[j];
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    int o = cnt[j];
    for (int e = 0; e < ne; ++e)
      if (col[ebase + e] == j) { trecv[ebase + o] = row[ebase + e]; tedge[ebase + o] = e; ++o; }
  }
}

int launch_transpose_relations(const Csr& csr, int B, int N, cudaStream_t st) {
  k_transpose_relations<<<B, 256, (N + 1) * sizeof(int), st>>>(csr.rowptr, csr.col, csr.row, N, csr.trowptr,
                                                               csr.trecv, csr.tedge);
  PILE_CHECK_LAUNCH();
  return 0;
}

// dst[b, i, :] += src[b, i, :] with per-sample strides
__global__ void k_add_strided(float* __restrict__ dst, long long d_stride, const float* __restrict__ src,
                              long long s_stride, int per, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / per, k = i % per;
  dst[b * d_stride + k] += src[b * s_stride + k];
}

int launch_add_strided(float* dst, long long d_stride, const float* src, long long s_stride, int B, int per,
                       cudaStream_t st) {
  const long long total = (long long)B * per;
  k_add_strided<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dst, d_stride, src, s_stride, per, total);
  PILE_CHECK_LAUNCH();
  return 0;
}

int launch_gen_s_delta(const float* s_cur, long long s_stride, const float* action, int act_stride,
                       const PushCam& cam, int B, int N, float* s_delta, cudaStream_t st) {
  k_gen_s_delta<<<B, 128, 0, st>>>(s_cur, s_stride, action, act_stride, cam, N, s_delta);
  PILE_CHECK_LAUNCH();
  return 0;
}

static size_t nbr_smem_bytes_mode(int N, bool transpose) {
  const size_t NP = (size_t)(N + NBR_BLOCK - 1) / NBR_BLOCK * NBR_BLOCK;
  size_t words = (size_t)(19 * N + 1 + 3) / 4 * 4 + 3 * NP + 4 * (size_t)N;      // 4N pos + 15N + 1 words, then xs | ys | zs | cur
  if (transpose) words += 2 * (size_t)KMAX * N;                 // + the transposed lists being sorted
  return sizeof(float) * words;
}
size_t nbr_smem_bytes(int N) { return nbr_smem_bytes_mode(N, true); }

int launch_nbr_search(const float* s_cur, long long s_stride, const float* s_delta_in, const float* action,
                      int act_stride, const PushCam& cam, float* s_delta_out, const int* particle_nums, int B,
                      int N, float thr, const Csr& csr, cudaStream_t st, const float* attr, const float* dens,
                      float* efeat) {
  const size_t smem = nbr_smem_bytes_mode(N, csr.trowptr != nullptr);
  if (nbr_smem_bytes(N) > 200 * 1024) return (int)cudaErrorInvalidValue;
  static DeviceOnce once;
  const int dev = once.pending();
  if (dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(k_nbr_search, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    once.done(dev);
  }
  int threads = (N + 31) / 32 * 32;
  threads = threads < 64 ? 64 : (threads > NBR_THREADS ? NBR_THREADS : threads);
  k_nbr_search<<<B, threads, smem, st>>>(s_cur, s_stride, s_delta_in, action, act_stride, cam, s_delta_out,
                                         particle_nums, N, thr, csr.rowptr, csr.col, csr.row, csr.trowptr,
                                         csr.trecv, csr.tedge, attr, dens, efeat);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
