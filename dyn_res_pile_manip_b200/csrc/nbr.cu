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

// One CTA per sample.  Optional fused s_delta (action != nullptr).
// SPLIT > 1 (small batches, where one 10-warp CTA per SM is latency-bound): the candidate range of every receiver is
// cut into SPLIT contiguous pieces scanned by SPLIT different warps (thread = piece * T0 + receiver slot, T0 =
// blockDim / SPLIT); the pieces share their admission bound through shared memory (a candidate farther than ANY
// piece's current 10th best cannot be in the union's top 10) and the SPLIT sorted lists are merged by (distance,
// index) -- the same order a single ascending scan produces, so the relation set is identical.
// dynamic smem: pos[N] (float4) | cutd[N] | cuti[N] | deg[N] | roff[N+1] | sel[N*KMAX] | perm[N]
//               | SPLIT > 1: bound[N] | list_d[SPLIT*N*KMAX] | list_j[SPLIT*N*KMAX]
template <int SPLIT>
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
  int* bound = perm + N;         // SPLIT > 1: per receiver, min over the pieces of the 10th best distance (float bits)
  float* list_d = reinterpret_cast<float*>(bound + N);     // [SPLIT][N][KMAX]
  int* list_j = reinterpret_cast<int*>(list_d + SPLIT * N * KMAX);
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
    pos[i] = make_float4(__fadd_rn(x, dx), __fadd_rn(y, dy), __fadd_rn(z, dz), 0.f);
  }
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
  const int T0 = (int)blockDim.x / SPLIT;                // receiver slots per pass; a multiple of 32
  const int piece = (int)threadIdx.x / T0, slot = (int)threadIdx.x - piece * T0;      // piece is warp-uniform
  const int jper = (N + SPLIT - 1) / SPLIT;
  const int jlo = min(piece * jper, N), jhi = min(jlo + jper, N);
  for (int base_i = 0; base_i < N; base_i += T0) {     // every thread takes part in the warp votes
    const bool active = base_i + slot < N;
    const int i = active ? perm[base_i + slot] : N;
    if (SPLIT > 1) {
      if (piece == 0 && active) bound[i] = 0x7f800000;
      __syncthreads();
    }
    float bd[KMAX];
    int id[KMAX];
#pragma unroll
    for (int s = 0; s < KMAX; ++s) { bd[s] = __int_as_float(0x7f800000); id[s] = 0x7fffffff; }
    float qd0 = 0.f, qd1 = 0.f, qd2 = 0.f;
    int qj0 = 0, qj1 = 0, qj2 = 0, qn = 0;
    const int ic = active ? i : 0;
    const float4 pi = pos[ic];
    const float xi = pi.x, yi = pi.y, zi = pi.z;
    // admission bound min(radius^2, current 10th best); -1 keeps lanes without a receiver out (d >= 0)
    float lim = active ? thr : -1.f;
    // pop the oldest parked candidate (if any) and insert it keeping (d, index) ascending; candidates of a
    // lane are inserted in ascending j, so equal distances keep the lower index first
    auto drain_one = [&]() {
      if (qn > 0) {
        const float d = qd0;
        const int j = qj0;
        qd0 = qd1; qj0 = qj1; qd1 = qd2; qj1 = qj2;
        --qn;
#pragma unroll
        for (int s = KMAX - 1; s > 0; --s) {
          const bool shift = d < bd[s - 1];
          const bool here = !shift && d < bd[s];
          const float nd = shift ? bd[s - 1] : (here ? d : bd[s]);
          const int ni = shift ? id[s - 1] : (here ? j : id[s]);
          bd[s] = nd; id[s] = ni;
        }
        if (d < bd[0]) { bd[0] = d; id[0] = j; }
        if (SPLIT == 1) {
          lim = fminf(thr, bd[KMAX - 1]);
        } else {
          // publish this piece's 10th best, take the tightest one of all pieces.  A candidate AT another piece's
          // bound may still win by its lower index, so foreign bounds admit d <= bound: compare against the next
          // float up (the insertion and the final merge re-check exactly)
          const int mine = __float_as_int(bd[KMAX - 1]);          // non-negative floats order like their bit patterns
          if (mine < 0x7f800000) atomicMin(&bound[ic], mine);
          const int seen = bound[ic];
          const float up = seen < 0x7f800000 ? __int_as_float(seen + 1) : __int_as_float(0x7f800000);
          lim = fminf(thr, up);
        }
      }
    };
#pragma unroll 4
    for (int j = jlo; j < jhi; ++j) {
      const float4 pj = pos[j];
      const float d = sqdist_rn(xi, yi, zi, pj.x, pj.y, pj.z);
      if (d < lim) {
        if (qn == 0) { qd0 = d; qj0 = j; } else if (qn == 1) { qd1 = d; qj1 = j; } else { qd2 = d; qj2 = j; }
        ++qn;
      }
      if (__any_sync(0xffffffffu, qn == 3)) drain_one();
    }
    while (__any_sync(0xffffffffu, qn > 0)) drain_one();
    if (SPLIT > 1) {
      // every piece leaves its sorted list in shared memory; piece 0 merges them by (distance, index)
      if (active) {
#pragma unroll
        for (int s = 0; s < KMAX; ++s) {
          list_d[(piece * N + i) * KMAX + s] = bd[s];
          list_j[(piece * N + i) * KMAX + s] = id[s];
        }
      }
      __syncthreads();
      if (piece == 0 && active) {
        int head[SPLIT];
#pragma unroll
        for (int q = 0; q < SPLIT; ++q) head[q] = 0;
#pragma unroll
        for (int s = 0; s < KMAX; ++s) {
          float best_d = __int_as_float(0x7f800000);
          int best_j = 0x7fffffff, best_q = 0;
#pragma unroll
          for (int q = 0; q < SPLIT; ++q) {
            // heads past the end read as (+inf, INT_MAX); pieces hold ascending index ranges, so on equal distances
            // the lower piece has the lower index: strict '<' keeps it
            const float dq = head[q] < KMAX ? list_d[(q * N + i) * KMAX + head[q]] : __int_as_float(0x7f800000);
            const int jq = head[q] < KMAX ? list_j[(q * N + i) * KMAX + head[q]] : 0x7fffffff;
            if (dq < best_d || (dq == best_d && jq < best_j)) { best_d = dq; best_j = jq; best_q = q; }
          }
#pragma unroll
          for (int q = 0; q < SPLIT; ++q) head[q] += (q == best_q && best_j != 0x7fffffff) ? 1 : 0;
          bd[s] = best_d;
          id[s] = best_j;
        }
      }
    }
    if (active && piece == 0) {
      cutd[i] = bd[KMAX - 1];          // +inf when fewer than 10 in radius
      cuti[i] = id[KMAX - 1];
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
    if (SPLIT > 1) __syncthreads();     // the lists and bounds are rewritten by the next pass
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
      const float* pr = s_cur + (long long)b * s_stride + i * 3;
      const float* ps = s_cur + (long long)b * s_stride + c * 3;
      float* out = efeat + e * 8;
      *reinterpret_cast<float4*>(out) = make_float4(attr[base + i], attr[base + c], pr[0] - ps[0], pr[1] - ps[1]);
      *reinterpret_cast<float4*>(out + 4) = make_float4(pr[2] - ps[2], dn, 0.f, 0.f);
    }
  }

  if (trowptr == nullptr) return;
  // sender-major transpose: for sender j the receivers i (ascending) with j in nbr(i), and the edge id
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    int n = 0;
    if (j < nvalid) {
      const float4 pj = pos[j];
      for (int i = 0; i < nvalid; ++i) {
        const float4 pi = pos[i];
        const float d = sqdist_rn(pi.x, pi.y, pi.z, pj.x, pj.y, pj.z);
        const float cd = cutd[i];
        n += (d < thr) && (d < cd || (d == cd && j <= cuti[i]));
      }
    }
    deg[j] = n;     // reuse as in-degree
  }
  __syncthreads();
  {
    const int lo = min(threadIdx.x * per, N), hi = min(lo + per, N);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += deg[i];
    int run = block_exscan(s, warp_sums, &total_s);
    int* trp = trowptr + (size_t)b * (N + 1);
    for (int i = lo; i < hi; ++i) { const int d = deg[i]; trp[i] = run; deg[i] = run; run += d; }
    if (threadIdx.x == 0) trp[N] = total_s;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nvalid; j += blockDim.x) {
    const float4 pj = pos[j];
    int o = deg[j];
    for (int i = 0; i < nvalid; ++i) {
      const float4 pi = pos[i];
      const float d = sqdist_rn(pi.x, pi.y, pi.z, pj.x, pj.y, pj.z);
      const float cd = cutd[i];
      if ((d < thr) && (d < cd || (d == cd && j <= cuti[i]))) {
        int pos = 0;
#pragma unroll
        for (int s = 0; s < KMAX; ++s) pos += sel[i * KMAX + s] < j;
        trecv[ebase + o] = i;
        tedge[ebase + o] = roff[i] + pos;
        ++o;
      }
    }
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

static size_t nbr_smem_bytes_split(int N, int split) {
  size_t b = sizeof(float) * (size_t)(4 * N + N) + sizeof(int) * (size_t)(N + N + N + 1 + N * KMAX + N);
  if (split > 1) b += sizeof(int) * (size_t)N + 2 * sizeof(float) * (size_t)split * N * KMAX;
  return b;
}
size_t nbr_smem_bytes(int N) { return nbr_smem_bytes_split(N, 1); }

// pieces per receiver: 1 when the batch alone fills the SMs with warps, up to 3 for small batches
static std::atomic<int> g_nbr_split_override{0};
int set_nbr_split(int split) { return g_nbr_split_override.exchange(split < 0 ? 0 : (split > 3 ? 3 : split)); }

static int nbr_split(int B, int T0, int N) {
  const int forced = g_nbr_split_override.load();
  int split = 1;
  const double fill = (double)B * T0 / (148.0 * 1024.0);       // resident threads per SM / 1024
  if (fill < 0.6) split = 3;
  else if (fill < 1.3) split = 2;
  if (forced >= 1 && forced <= 3) split = forced;
  while (split > 1 && (split * T0 > NBR_THREADS || nbr_smem_bytes_split(N, split) > 200 * 1024)) --split;
  return split;
}

int launch_nbr_search(const float* s_cur, long long s_stride, const float* s_delta_in, const float* action,
                      int act_stride, const PushCam& cam, float* s_delta_out, const int* particle_nums, int B,
                      int N, float thr, const Csr& csr, cudaStream_t st, const float* attr, const float* dens,
                      float* efeat) {
  if (nbr_smem_bytes(N) > 200 * 1024) return (int)cudaErrorInvalidValue;
  static DeviceOnce once;
  const int dev = once.pending();
  if (dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(k_nbr_search<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_nbr_search<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_nbr_search<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    once.done(dev);
  }
  int threads = (N + 31) / 32 * 32;
  threads = threads < 64 ? 64 : (threads > NBR_THREADS ? NBR_THREADS : threads);
  const int split = nbr_split(B, threads, N);
  const size_t smem = nbr_smem_bytes_split(N, split);
  auto kern = split == 3 ? k_nbr_search<3> : (split == 2 ? k_nbr_search<2> : k_nbr_search<1>);
  kern<<<B, threads * split, smem, st>>>(s_cur, s_stride, s_delta_in, action, act_stride, cam, s_delta_out,
                                         particle_nums, N, thr, csr.rowptr, csr.col, csr.row, csr.trowptr,
                                         csr.trecv, csr.tedge, attr, dens, efeat);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
