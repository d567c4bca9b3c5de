// General-width engine of the propagation network: any nf_effect (reference model/gnn_dyn.py:119 reads it from the
// config; the planner engines in fwd.cu / *_tc.cu are compiled for 64), forward that keeps every layer input and a
// backward that returns d/ds_cur, d/ds_delta and -- optionally -- the weight gradients of the nine linear layers
// (train/train_gnn_dyn.py:150-199).  Same un-hoisted formulation as train.cu:
//
//   forward   X0 -PE0-> H0 -PE1-> P = eff_0          Y0 -RE0-> R1 -RE1-> R2 -RE2-> R3
//             p = 0..2:  M_p = ReLU([R3 | eff_p[recv] | eff_p[send] | d] W_rp^T + b)      (relation rows)
//                        agg_p = segment-sum of M_p over receivers
//                        eff_{p+1} = ReLU([P | agg_p | d] W_pp^T + b + eff_p)              (particle rows)
//             Q = ReLU(eff_3 V0^T + b),  s_pred = Q V1^T + b + s_cur
//
// Width: the hidden width H is padded to Hp = 64 * nb (nb <= 4): every feature array is [rows, Hp] row-major and
// every weight matrix is stored zero-padded to Hp x Hp, so padded channels are exactly 0 after every ReLU and the
// real channels see the same sums in the same order as without padding.  A linear layer is a block GEMM over 64-wide
// column blocks: one CTA owns an output block `ob` and walks (source, source block) pairs, 128-row tiles, FP32 on
// the CUDA cores (gemm_rows4x8, common.cuh).  The backward of a layer is three generic steps:
//   Gm = upstream gradient (gathered by receiver for relation rows) masked by the layer's output > 0   (k_g_mask)
//   dX_s = Gm W_s                  the forward kernel again, with W in the checkpoint's [out][in] layout
//   dW_s[ob][ib] += Gm[:, ob]^T X_s[:, ib]   per-CTA partial sums in registers over all its tiles, written once and
//                  added in a fixed order by k_g_finish: deterministic, no float atomics.
#include "common.cuh"
#include "kernels.h"

namespace pile {
namespace general {

constexpr int MAXSRC = 3;
constexpr int MAXNB = 4;
constexpr int BW = 64;                    // block width

struct Src {
  const float* x;      // [*, Hp] rows (particle rows when gather != 0)
  const float* w;      // [Hp in][Hp out] (forward) / [Hp out][Hp in] (dX)
  int gather;          // relation kernels: 0 = the relation's own row, 1 = its receiver particle's row, 2 = its sender's
};

struct LinArgs {
  int B, N, nb;
  const int* rowptr;
  const int* col;
  const int* row;
  int nsrc;
  Src src[MAXSRC];
  const float* x8;     // [rows, 8] narrow input block (nullptr: none)
  const float* w8;     // [8][Hp]
  const float* dens;   // [B] (nullptr: no density column)
  const float* wd;     // [Hp]
  const float* bias;   // [Hp] (nullptr: 0)
  const float* res;    // residual [rows, Hp] added before the ReLU (nullptr: none)
  float* y;            // [rows, Hp]
  int relu;
  int accumulate;      // y += instead of y =
};

struct Tile {
  long long row0;
  int nrows;
  int b;
};

template <bool EDGE>
__device__ __forceinline__ int num_tiles(int B, int N) {
  if (EDGE) return B * ((KMAX * N + TILE - 1) / TILE);
  return (int)(((long long)B * N + TILE - 1) / TILE);
}

template <bool EDGE>
__device__ __forceinline__ Tile get_tile(int B, int N, const int* __restrict__ rowptr, int t) {
  Tile q;
  if (EDGE) {
    const int tps = (KMAX * N + TILE - 1) / TILE;
    q.b = t / tps;
    const int e0 = (t - q.b * tps) * TILE;
    const int ne = rowptr[(long long)q.b * (N + 1) + N];
    q.nrows = min(TILE, ne - e0);
    q.row0 = (long long)q.b * KMAX * N + e0;
  } else {
    q.b = 0;
    q.row0 = (long long)t * TILE;
    q.nrows = (int)min((long long)TILE, (long long)B * N - q.row0);
  }
  return q;
}

template <bool EDGE>
__device__ __forceinline__ long long src_row(int N, const int* __restrict__ row, const int* __restrict__ col,
                                             const Tile& q, int r, int gather) {
  if (!EDGE || gather == 0) return q.row0 + r;
  const int* ix = gather == 1 ? row : col;
  return (long long)q.b * N + ix[q.row0 + r];
}

// column block `cb` of the rows of a [*, ld] array -> shared tile [TILE][LDA] (half-warp per row, zero fill)
template <bool EDGE>
__device__ __forceinline__ void load_rows(int N, const int* __restrict__ row, const int* __restrict__ col, const Tile& q,
                                          const float* __restrict__ x, int ld, int cb, int gather, float* __restrict__ A) {
  const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
  for (int r = hw; r < TILE; r += NT / 16) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < q.nrows) v = ld4(x + src_row<EDGE>(N, row, col, q, r, gather) * ld + cb * BW + 4 * l16);
    st4(A + r * LDA + 4 * l16, v);
  }
}

__device__ __forceinline__ void load_rows8(const Tile& q, const float* __restrict__ x8, float* __restrict__ A8) {
  for (int idx = threadIdx.x; idx < TILE * 2; idx += NT) {
    const int r = idx >> 1, h = idx & 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < q.nrows) v = ld4(x8 + (q.row0 + r) * 8 + 4 * h);
    st4(A8 + r * LDX + 4 * h, v);
  }
}

// [rows x 64] block of a matrix with row stride ld -> dense shared [rows][64]
__device__ __forceinline__ void load_wblock(float* __restrict__ dst, const float* __restrict__ src, int ld, int rows) {
  for (int i = threadIdx.x * 4; i < rows * BW; i += NT * 4) {
    const int r = i >> 6, c = i & 63;
    st4(dst + i, ld4(src + (long long)r * ld + c));
  }
}

template <bool EDGE>
__device__ __forceinline__ float dens_of(const float* __restrict__ dens, int N, const Tile& q, int r) {
  if (dens == nullptr || r >= q.nrows) return 0.f;
  const int b = EDGE ? q.b : (int)((q.row0 + r) / N);
  return dens[b] / 5000.f;           // gnn_dyn.py:158
}

// ------------------------------------------------------------------------------------------------
// y[:, ob] (+)= act(sum_s x_s W_s + x8 W8 + d wd + bias + res)
// ------------------------------------------------------------------------------------------------
struct LinSmem {
  float w[MAXSRC][BW * BW];
  float w8[8 * BW];
  float bias[BW], wd[BW];
  float A[TILE * LDA];
  float A8[TILE * LDX];
};

template <bool EDGE>
__global__ void __launch_bounds__(NT, 2) k_g_lin(LinArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LinSmem& S = *reinterpret_cast<LinSmem*>(smem_raw);
  const int nb = a.nb, Hp = nb * BW, ob = blockIdx.y;
  // all weight blocks of this output block stay in shared memory when there are at most three of them
  const bool resident = a.nsrc * nb <= MAXSRC;
  if (resident)
    for (int s = 0; s < a.nsrc; ++s)
      for (int sb = 0; sb < nb; ++sb) load_wblock(S.w[s * nb + sb], a.src[s].w + (long long)sb * BW * Hp + ob * BW, Hp, BW);
  if (a.x8) load_wblock(S.w8, a.w8 + ob * BW, Hp, 8);
  if (threadIdx.x < BW) {
    S.bias[threadIdx.x] = a.bias ? a.bias[ob * BW + threadIdx.x] : 0.f;
    S.wd[threadIdx.x] = a.wd ? a.wd[ob * BW + threadIdx.x] : 0.f;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c0 = warp * 8;
  const int ntiles = num_tiles<EDGE>(a.B, a.N);
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const Tile q = get_tile<EDGE>(a.B, a.N, a.rowptr, t);
    if (q.nrows <= 0) continue;          // CTA-uniform
    __syncthreads();
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float d = dens_of<EDGE>(a.dens, a.N, q, lane + 32 * i);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(d, S.wd[c0 + j], S.bias[c0 + j]);
    }
    if (a.res) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = lane + 32 * i;
        if (r < q.nrows) {
          const float* p = a.res + (q.row0 + r) * Hp + ob * BW + c0;
          const float4 u = ld4(p), v = ld4(p + 4);
          acc[i][0] += u.x; acc[i][1] += u.y; acc[i][2] += u.z; acc[i][3] += u.w;
          acc[i][4] += v.x; acc[i][5] += v.y; acc[i][6] += v.z; acc[i][7] += v.w;
        }
      }
    }
    bool first = true;
    for (int s = 0; s < a.nsrc; ++s)
      for (int sb = 0; sb < nb; ++sb) {
        if (!first) __syncthreads();
        first = false;
        load_rows<EDGE>(a.N, a.row, a.col, q, a.src[s].x, Hp, sb, a.src[s].gather, S.A);
        if (!resident) load_wblock(S.w[0], a.src[s].w + (long long)sb * BW * Hp + ob * BW, Hp, BW);
        __syncthreads();
        gemm_rows4x8<BW>(S.A, LDA, resident ? S.w[s * nb + sb] : S.w[0], lane, c0, acc);
      }
    if (a.x8) {
      load_rows8(q, a.x8, S.A8);
      __syncthreads();
      gemm_rows4x8<8>(S.A8, LDX, S.w8, lane, c0, acc);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = lane + 32 * i;
      if (r < q.nrows) {
        float* p = a.y + (q.row0 + r) * Hp + ob * BW + c0;
        if (a.accumulate) {
          const float4 u = ld4(p), v = ld4(p + 4);
          acc[i][0] += u.x; acc[i][1] += u.y; acc[i][2] += u.z; acc[i][3] += u.w;
          acc[i][4] += v.x; acc[i][5] += v.y; acc[i][6] += v.z; acc[i][7] += v.w;
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaxf(acc[i][j], 0.f);
        }
        st4(p, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        st4(p + 4, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Gm[r] = g[src row of r] * (y[r] > 0)       (one float4 per thread)
// ------------------------------------------------------------------------------------------------
template <bool EDGE>
__global__ void k_g_mask(const float* __restrict__ g, int g_gather, const float* __restrict__ ymask,
                         const int* __restrict__ rowptr, const int* __restrict__ row, float* __restrict__ Gm, int B,
                         int N, int Hp) {
  const int q4 = Hp / 4;
  if (EDGE) {
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
      const int ne = rowptr[(long long)b * (N + 1) + N];
      const long long slot0 = (long long)b * KMAX * N;
      for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)ne * q4;
           idx += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(idx / q4), c = (int)(idx - (long long)e * q4) * 4;
        const long long gr = g_gather ? (long long)b * N + row[slot0 + e] : slot0 + e;
        float4 v = ld4(g + gr * Hp + c);
        if (ymask) {
          const float4 y = ld4(ymask + (slot0 + e) * Hp + c);
          v.x = y.x > 0.f ? v.x : 0.f; v.y = y.y > 0.f ? v.y : 0.f; v.z = y.z > 0.f ? v.z : 0.f; v.w = y.w > 0.f ? v.w : 0.f;
        }
        st4(Gm + (slot0 + e) * Hp + c, v);
      }
    }
  } else {
    const long long total = (long long)B * N * q4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
      float4 v = ld4(g + idx * 4);
      if (ymask) {
        const float4 y = ld4(ymask + idx * 4);
        v.x = y.x > 0.f ? v.x : 0.f; v.y = y.y > 0.f ? v.y : 0.f; v.z = y.z > 0.f ? v.z : 0.f; v.w = y.w > 0.f ? v.w : 0.f;
      }
      st4(Gm + idx * 4, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight-gradient partial sums of one layer; grid (gx, nb * nb): CTA (x, ob * nb + ib)
// ------------------------------------------------------------------------------------------------
struct DwArgs {
  int B, N, nb;
  const int* rowptr;
  const int* col;
  const int* row;
  const float* G;      // masked upstream gradient [rows, Hp]
  int nsrc;
  Src src[MAXSRC];     // x, gather (w unused)
  const float* x8;
  const float* dens;
  float* partial;      // [gridDim.y][gridDim.x][PARTIAL]
};
constexpr int PARTIAL = MAXSRC * BW * BW + BW * 8 + BW + BW;     // dW blocks | dW8 | dbias | dwd
constexpr int OFF8 = MAXSRC * BW * BW, OFFB = OFF8 + BW * 8, OFFD = OFFB + BW;

struct DwSmem {
  float G[TILE * LDA];
  float X[TILE * LDA];
  float X8[TILE * LDX];
  float dn[TILE];
};

template <bool EDGE>
__global__ void __launch_bounds__(NT, 2) k_g_dw(DwArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DwSmem& S = *reinterpret_cast<DwSmem*>(smem_raw);
  const int nb = a.nb, Hp = nb * BW;
  const int ob = blockIdx.y / nb, ib = blockIdx.y - ob * nb;
  const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
  const int o0 = hw * 4, k0 = l16 * 4;           // this thread's 4 x 4 block of every dW block
  float dw[MAXSRC][4][4];
#pragma unroll
  for (int s = 0; s < MAXSRC; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dw[s][i][j] = 0.f;
  float dw8[2] = {0.f, 0.f};                     // thread t: out = t / 4, in = 2 * (t % 4) + {0, 1}
  float dbias = 0.f, dwd = 0.f;                  // threads < 64: column threadIdx.x
  const int ntiles = num_tiles<EDGE>(a.B, a.N);
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const Tile q = get_tile<EDGE>(a.B, a.N, a.rowptr, t);
    if (q.nrows <= 0) continue;          // CTA-uniform
    __syncthreads();
    load_rows<EDGE>(a.N, a.row, a.col, q, a.G, Hp, ob, 0, S.G);
    if (ib == 0 && threadIdx.x < TILE) S.dn[threadIdx.x] = dens_of<EDGE>(a.dens, a.N, q, threadIdx.x);
    __syncthreads();
    if (ib == 0 && threadIdx.x < BW) {
      float sb = 0.f, sd = 0.f;
      for (int r = 0; r < q.nrows; ++r) {
        const float gv = S.G[r * LDA + threadIdx.x];
        sb += gv;
        sd = fmaf(gv, S.dn[r], sd);
      }
      dbias += sb;
      dwd += sd;
    }
#pragma unroll
    for (int s = 0; s < MAXSRC; ++s) {
      if (s >= a.nsrc) break;                  // CTA-uniform
      if (s) __syncthreads();
      load_rows<EDGE>(a.N, a.row, a.col, q, a.src[s].x, Hp, ib, a.src[s].gather, S.X);
      __syncthreads();
#pragma unroll 4
      for (int r = 0; r < TILE; ++r) {
        const float4 g4 = ld4(S.G + r * LDA + o0), x4 = ld4(S.X + r * LDA + k0);
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dw[s][i][j] = fmaf(gv[i], xv[j], dw[s][i][j]);
      }
    }
    if (a.x8 && ib == 0) {
      load_rows8(q, a.x8, S.X8);
      __syncthreads();
      const int o = threadIdx.x >> 2, kk = (threadIdx.x & 3) * 2;
      float s0 = 0.f, s1 = 0.f;
      for (int r = 0; r < q.nrows; ++r) {
        const float gv = S.G[r * LDA + o];
        s0 = fmaf(gv, S.X8[r * LDX + kk], s0);
        s1 = fmaf(gv, S.X8[r * LDX + kk + 1], s1);
      }
      dw8[0] += s0;
      dw8[1] += s1;
    }
  }
  float* P = a.partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * PARTIAL;
#pragma unroll
  for (int s = 0; s < MAXSRC; ++s) {
    if (s >= a.nsrc) break;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      st4(P + s * BW * BW + (o0 + i) * BW + k0, make_float4(dw[s][i][0], dw[s][i][1], dw[s][i][2], dw[s][i][3]));
  }
  {
    const int o = threadIdx.x >> 2, kk = (threadIdx.x & 3) * 2;
    P[OFF8 + o * 8 + kk] = dw8[0];
    P[OFF8 + o * 8 + kk + 1] = dw8[1];
  }
  if (threadIdx.x < BW) {
    P[OFFB + threadIdx.x] = dbias;
    P[OFFD + threadIdx.x] = dwd;
  }
}

// fixed-order sum of the CTAs' partial sums into the gradient tensors (reference shapes, unpadded)
struct FinishArgs {
  const float* partial;
  int gx, nb, H;
  int nsrc;
  struct { float* dest; int ld, c0; } src[MAXSRC];     // dW_s[o][k] -> dest[o * ld + c0 + k]
  float* dest8; int ld8, k8;                            // dW8[o][k < k8]
  float* dbias;
  float* dwd; int ld_wd;                                // dwd[o] -> dwd[o * ld_wd]
};

__device__ __forceinline__ float sum_parts(const float* __restrict__ src, int n, long long stride) {
  float s = 0.f;
  int c = 0;
  for (; c + 8 <= n; c += 8) {          // eight loads in flight, summed in the fixed order c = 0, 1, 2, ...
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = src[(long long)(c + u) * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; c < n; ++c) s += src[(long long)c * stride];
  return s;
}

__global__ void k_g_finish(FinishArgs f) {
  const int H = f.H, nb = f.nb;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int s = 0; s < f.nsrc; ++s)
    for (int idx = tid; idx < H * H; idx += nth) {
      const int o = idx / H, k = idx - o * H;
      const int plane = (o / BW) * nb + k / BW;
      const float* src = f.partial + (long long)plane * f.gx * PARTIAL + s * BW * BW + (o % BW) * BW + (k % BW);
      f.src[s].dest[(long long)o * f.src[s].ld + f.src[s].c0 + k] += sum_parts(src, f.gx, PARTIAL);
    }
  if (f.dest8)
    for (int idx = tid; idx < H * f.k8; idx += nth) {
      const int o = idx / f.k8, k = idx - o * f.k8;
      const float* src = f.partial + (long long)((o / BW) * nb) * f.gx * PARTIAL + OFF8 + (o % BW) * 8 + k;
      f.dest8[(long long)o * f.ld8 + k] += sum_parts(src, f.gx, PARTIAL);
    }
  for (int o = tid; o < H; o += nth) {
    const float* base = f.partial + (long long)((o / BW) * nb) * f.gx * PARTIAL;
    if (f.dbias) f.dbias[o] += sum_parts(base + OFFB + (o % BW), f.gx, PARTIAL);
    if (f.dwd) f.dwd[(long long)o * f.ld_wd] += sum_parts(base + OFFD + (o % BW), f.gx, PARTIAL);
  }
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
// X0[r] = (s_delta xyz, attr, dens / 5000, 0, 0, 0)   (gnn_dyn.py:174-175)
__global__ void k_g_node_in(const float* __restrict__ s_delta, const float* __restrict__ attr,
                            const float* __restrict__ dens, float* __restrict__ X0, int B, int N) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (long long)B * N) return;
  const float* sd = s_delta + r * 3;
  st4(X0 + r * 8, make_float4(sd[0], sd[1], sd[2], attr[r]));
  st4(X0 + r * 8 + 4, make_float4(dens[r / N] / 5000.f, 0.f, 0.f, 0.f));
}

// Y0[e] = (attr_r, attr_s, s_r - s_s, dens / 5000, 0, 0)   (gnn_dyn.py:164-172, 179-180)
__global__ void k_g_edge_in(const float* __restrict__ attr, const float* __restrict__ dens, const float* __restrict__ s_cur,
                            const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ row,
                            float* __restrict__ Y0, int B, int N) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    if (e >= rowptr[(long long)b * (N + 1) + N]) continue;
    const long long slot = (long long)b * KMAX * N + e;
    const int r = row[slot], c = col[slot];
    const float* pr = s_cur + ((long long)b * N + r) * 3;
    const float* ps = s_cur + ((long long)b * N + c) * 3;
    st4(Y0 + slot * 8, make_float4(attr[(long long)b * N + r], attr[(long long)b * N + c], pr[0] - ps[0], pr[1] - ps[1]));
    st4(Y0 + slot * 8 + 4, make_float4(pr[2] - ps[2], dens[b] / 5000.f, 0.f, 0.f));
  }
}

// out[i] (+)= base[i] + sum_{e in row i} A[e] + sum_{k in trow i} Bs[tedge k]   ([*, Hp] rows; one thread per float4)
// (base / Bs / trowptr nullable: the plain receiver-segment sum of the forward, gnn_dyn.py:189)
__global__ void k_g_gather_nodes(const int* __restrict__ rowptr, const int* __restrict__ trowptr,
                                 const int* __restrict__ tedge, const float* __restrict__ base,
                                 const float* __restrict__ A, const float* __restrict__ Bs, float* __restrict__ out,
                                 int accumulate, int B, int N, int Hp) {
  const int q4 = Hp / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * N * q4) return;
  const long long node = idx / q4;
  const int c = (int)(idx - node * q4) * 4;
  const int b = (int)(node / N), i = (int)(node - (long long)b * N);
  const long long slot = (long long)b * KMAX * N;
  float4 s = base ? ld4(base + node * Hp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int* rp = rowptr + (long long)b * (N + 1) + i;
  for (int e = rp[0]; e < rp[1]; ++e) {
    const float4 v = ld4(A + (slot + e) * Hp + c);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  if (Bs) {
    const int* tp = trowptr + (long long)b * (N + 1) + i;
    for (int k = tp[0]; k < tp[1]; ++k) {
      const float4 v = ld4(Bs + (slot + tedge[slot + k]) * Hp + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  float* o = out + node * Hp + c;
  if (accumulate) { const float4 u = ld4(o); s.x += u.x; s.y += u.y; s.z += u.z; s.w += u.w; }
  st4(o, s);
}

// hoisted relation propagator (inference): agg[i] = sum_{e in row i} ReLU(Ce[e] + Pr[i] + Ps[col e])   (one thread per float4)
// with Ce = W_e r3 + w_d d + b per relation (once per model step) and (Pr, Ps) = (W_r, W_s) eff per PARTICLE -- the same
// regrouping of Linear([r3, eff_r, eff_s, d]) the width-64 planner engines use
__global__ void k_g_agg_hoisted(const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ Ce,
                                const float* __restrict__ Pr, const float* __restrict__ Ps, float* __restrict__ out, int B,
                                int N, int Hp) {
  const int q4 = Hp / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * N * q4) return;
  const long long node = idx / q4;
  const int c = (int)(idx - node * q4) * 4;
  const int b = (int)(node / N), i = (int)(node - (long long)b * N);
  const long long slot = (long long)b * KMAX * N;
  const float4 pr = ld4(Pr + node * Hp + c);
  const float* ps_b = Ps + (long long)b * N * Hp + c;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  const int* rp = rowptr + (long long)b * (N + 1) + i;
  for (int e = rp[0]; e < rp[1]; ++e) {
    const float4 ce = ld4(Ce + (slot + e) * Hp + c);
    const float4 ps = ld4(ps_b + (long long)col[slot + e] * Hp);
    s.x += fmaxf(ce.x + pr.x + ps.x, 0.f);
    s.y += fmaxf(ce.y + pr.y + ps.y, 0.f);
    s.z += fmaxf(ce.z + pr.z + ps.z, 0.f);
    s.w += fmaxf(ce.w + pr.w + ps.w, 0.f);
  }
  st4(out + node * Hp + c, s);
}

// s_pred = Q V1^T + b + s_cur   (gnn_dyn.py:196-198); V1T [Hp][4], b [4]
__global__ void k_g_predict(const float* __restrict__ Q, const float* __restrict__ v1t, const float* __restrict__ b1,
                            const float* __restrict__ s_cur, float* __restrict__ s_out, int B, int N, int Hp) {
  extern __shared__ __align__(16) float w_sh[];
  for (int i = threadIdx.x; i < Hp * 4; i += blockDim.x) w_sh[i] = v1t[i];
  __syncthreads();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (long long)B * N) return;
  float o0 = b1[0], o1 = b1[1], o2 = b1[2];
  for (int k = 0; k < Hp; k += 4) {
    const float4 q = ld4(Q + r * Hp + k);
    o0 = fmaf(q.x, w_sh[(k + 0) * 4 + 0], o0); o1 = fmaf(q.x, w_sh[(k + 0) * 4 + 1], o1); o2 = fmaf(q.x, w_sh[(k + 0) * 4 + 2], o2);
    o0 = fmaf(q.y, w_sh[(k + 1) * 4 + 0], o0); o1 = fmaf(q.y, w_sh[(k + 1) * 4 + 1], o1); o2 = fmaf(q.y, w_sh[(k + 1) * 4 + 2], o2);
    o0 = fmaf(q.z, w_sh[(k + 2) * 4 + 0], o0); o1 = fmaf(q.z, w_sh[(k + 2) * 4 + 1], o1); o2 = fmaf(q.z, w_sh[(k + 2) * 4 + 2], o2);
    o0 = fmaf(q.w, w_sh[(k + 3) * 4 + 0], o0); o1 = fmaf(q.w, w_sh[(k + 3) * 4 + 1], o1); o2 = fmaf(q.w, w_sh[(k + 3) * 4 + 2], o2);
  }
  s_out[r * 3 + 0] = o0 + s_cur[r * 3 + 0];
  s_out[r * 3 + 1] = o1 + s_cur[r * 3 + 1];
  s_out[r * 3 + 2] = o2 + s_cur[r * 3 + 2];
}

// gQ[r] = g[r] V1 (the V0 layer's backward masks it with Q > 0);  ppart[cta] = (dV1 [3][Hp] | db [3] ...) as [4][Hp]
constexpr int PB_ROWS = 4;
__global__ void k_g_predict_bwd(const float* __restrict__ g, const float* __restrict__ Q, const float* __restrict__ v1,
                                float* __restrict__ gQ, float* __restrict__ ppart, int B, int N, int Hp) {
  extern __shared__ __align__(16) float pb_sh[];      // v1 [4][Hp] | red [PB_ROWS][4][Hp]
  float* w = pb_sh;
  float* red = pb_sh + 4 * Hp;
  for (int i = threadIdx.x; i < 4 * Hp; i += blockDim.x) w[i] = v1[i];
  __syncthreads();
  const long long R = (long long)B * N;
  const int k = threadIdx.x % Hp, sub = threadIdx.x / Hp;        // blockDim = PB_ROWS * Hp
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (long long r = (long long)blockIdx.x * PB_ROWS + sub; r < R; r += (long long)gridDim.x * PB_ROWS) {
    const float g0 = g[r * 3], g1 = g[r * 3 + 1], g2 = g[r * 3 + 2];
    gQ[r * Hp + k] = g0 * w[k] + g1 * w[Hp + k] + g2 * w[2 * Hp + k];
    const float q = Q[r * Hp + k];
    a0 = fmaf(g0, q, a0); a1 = fmaf(g1, q, a1); a2 = fmaf(g2, q, a2);
    s0 += g0; s1 += g1; s2 += g2;
  }
  float* rd = red + (long long)sub * 4 * Hp;
  rd[k] = a0; rd[Hp + k] = a1; rd[2 * Hp + k] = a2;
  rd[3 * Hp + k] = k == 0 ? s0 : (k == 1 ? s1 : (k == 2 ? s2 : 0.f));
  __syncthreads();
  if (sub == 0) {
    float* P = ppart + (long long)blockIdx.x * 4 * Hp;
    for (int c = 0; c < 4; ++c) {
      float t = 0.f;
      for (int u = 0; u < PB_ROWS; ++u) t += red[((long long)u * 4 + c) * Hp + k];
      P[c * Hp + k] = t;
    }
  }
}

__global__ void k_g_predict_finish(const float* __restrict__ ppart, int nparts, int Hp, int H, float* __restrict__ dV1,
                                   float* __restrict__ db) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int idx = tid; idx < 3 * H; idx += nth) {
    const int j = idx / H, k = idx - j * H;
    dV1[idx] += sum_parts(ppart + j * Hp + k, nparts, 4LL * Hp);
  }
  for (int j = tid; j < 3; j += nth) db[j] += sum_parts(ppart + 3 * Hp + j, nparts, 4LL * Hp);
}

// dX8[r] = Gm[r] W8   (W8 [Hp][8]); one row per thread
__global__ void k_g_dx8(const float* __restrict__ Gm, const float* __restrict__ w8, const int* __restrict__ rowptr,
                        float* __restrict__ dx8, int B, int N, int Hp, int edge) {
  extern __shared__ __align__(16) float w8_sh[];
  for (int i = threadIdx.x; i < Hp * 8; i += blockDim.x) w8_sh[i] = w8[i];
  __syncthreads();
  for (int b = blockIdx.y; b < (edge ? B : 1); b += gridDim.y) {
    long long r;
    if (edge) {
      const int e = blockIdx.x * blockDim.x + threadIdx.x;
      if (e >= rowptr[(long long)b * (N + 1) + N]) continue;
      r = (long long)b * KMAX * N + e;
    } else {
      r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
      if (r >= (long long)B * N) continue;
    }
    float o8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < Hp; k += 4) {
      const float4 gv = ld4(Gm + r * Hp + k);
      const float g4[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) o8[j] = fmaf(g4[u], w8_sh[(k + u) * 8 + j], o8[j]);
    }
    st4(dx8 + r * 8, make_float4(o8[0], o8[1], o8[2], o8[3]));
    st4(dx8 + r * 8 + 4, make_float4(o8[4], o8[5], o8[6], o8[7]));
  }
}

// g_s_cur[i] = g_pred[i] + sum_{e in row i} dY0[e][2:5] - sum_{e: sender = i} dY0[e][2:5];  g_s_delta[i] = dX0[i][0:3]
__global__ void k_g_positions(const int* __restrict__ rowptr, const int* __restrict__ trowptr,
                              const int* __restrict__ tedge, const float* __restrict__ dY0,
                              const float* __restrict__ dX0, const float* __restrict__ g_pred,
                              float* __restrict__ g_s_cur, float* __restrict__ g_s_delta, int B, int N) {
  const long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= (long long)B * N) return;
  const int b = (int)(node / N), i = (int)(node - (long long)b * N);
  const long long slot = (long long)b * KMAX * N;
  float x = g_pred[node * 3], y = g_pred[node * 3 + 1], z = g_pred[node * 3 + 2];
  const int* rp = rowptr + (long long)b * (N + 1) + i;
  for (int e = rp[0]; e < rp[1]; ++e) {
    const float* d = dY0 + (slot + e) * 8;
    x += d[2]; y += d[3]; z += d[4];
  }
  const int* tp = trowptr + (long long)b * (N + 1) + i;
  for (int k = tp[0]; k < tp[1]; ++k) {
    const float* d = dY0 + (slot + tedge[slot + k]) * 8;
    x -= d[2]; y -= d[3]; z -= d[4];
  }
  g_s_cur[node * 3] = x; g_s_cur[node * 3 + 1] = y; g_s_cur[node * 3 + 2] = z;
  g_s_delta[node * 3] = dX0[node * 8]; g_s_delta[node * 3 + 1] = dX0[node * 8 + 1]; g_s_delta[node * 3 + 2] = dX0[node * 8 + 2];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
size_t up256(size_t x) { return (x + 255) / 256 * 256; }

struct Carve {
  char* base;
  size_t off = 0;
  explicit Carve(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    T* q = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += up256(n * sizeof(T));
    return q;
  }
};

inline int nblocks(int H) { return (H + BW - 1) / BW; }

// packed weights for width H (floats): the non-tensor-core slots of enum WSlot with every H replaced by Hp
long long slot_size(int s, int Hp) {
  switch (s) {
    case W_PE0T: case W_RE0T: case W_PE0: case W_RE0: return 8LL * Hp;
    case B_PE0: case B_PE1: case B_RE0: case B_RE1: case B_RE2: case WD_RP: case B_RP:
    case WD_PP: case B_PP: case B_V0: return Hp;
    case W_V1T: case W_V1: return 4LL * Hp;
    case B_V1: return 4;
    default: return (long long)Hp * Hp;
  }
}
constexpr int NSLOTS = W_V1 + 1;      // slots W_PE0T .. W_V1
long long slot_offset(int s, int Hp) {
  long long o = 0;
  for (int i = 0; i < s; ++i) o += slot_size(i, Hp);
  return o;
}

struct Tape {
  Csr csr;
  float *X0, *H0, *P, *eff[PSTEP], *agg[PSTEP], *Q;      // particle rows
  float *Y0, *R1, *R2, *R3, *M[PSTEP];                   // relation slots
  size_t bytes;
};

Tape carve_tape(void* p, int B, int N, int Hp) {
  Carve c(p);
  const size_t R = (size_t)B * N, E = (size_t)B * KMAX * N;
  Tape t;
  t.csr.rowptr = c.take<int>((size_t)B * (N + 1));
  t.csr.col = c.take<int>(E);
  t.csr.row = c.take<int>(E);
  t.csr.trowptr = c.take<int>((size_t)B * (N + 1));
  t.csr.trecv = c.take<int>(E);
  t.csr.tedge = c.take<int>(E);
  t.X0 = c.take<float>(R * 8);
  t.H0 = c.take<float>(R * Hp);
  t.P = c.take<float>(R * Hp);
  for (int p2 = 0; p2 < PSTEP; ++p2) t.eff[p2] = c.take<float>(R * Hp);
  for (int p2 = 0; p2 < PSTEP; ++p2) t.agg[p2] = c.take<float>(R * Hp);
  t.Q = c.take<float>(R * Hp);
  t.Y0 = c.take<float>((E + TILE) * 8);
  t.R1 = c.take<float>(E * Hp);
  t.R2 = c.take<float>(E * Hp);
  t.R3 = c.take<float>(E * Hp);
  for (int p2 = 0; p2 < PSTEP; ++p2) t.M[p2] = c.take<float>(E * Hp);
  t.bytes = c.off;
  return t;
}

inline int dw_gx(int nb) { const int g = 2 * NSM / (nb * nb); return g < 8 ? 8 : g; }

struct BwdScratch {
  float *gA, *gEff, *gmN, *gP, *gAgg, *gH0, *dX0;     // particle rows
  float *gmE, *gR3, *dZr, *dZs, *gR2, *dY0;           // relation slots (gR1 reuses dZr)
  float* partial;
  float* ppart;
  size_t bytes;
};

BwdScratch carve_bwd(void* p, int B, int N, int Hp) {
  Carve c(p);
  const size_t R = (size_t)B * N, E = (size_t)B * KMAX * N;
  const int nb = Hp / BW;
  BwdScratch s;
  s.gA = c.take<float>(R * Hp);
  s.gEff = c.take<float>(R * Hp);
  s.gmN = c.take<float>(R * Hp);
  s.gP = c.take<float>(R * Hp);
  s.gAgg = c.take<float>(R * Hp);
  s.gH0 = c.take<float>(R * Hp);
  s.dX0 = c.take<float>(R * 8);
  s.gmE = c.take<float>(E * Hp);
  s.gR3 = c.take<float>(E * Hp);
  s.dZr = c.take<float>(E * Hp);
  s.dZs = c.take<float>(E * Hp);
  s.gR2 = c.take<float>(E * Hp);
  s.dY0 = c.take<float>(E * 8);
  s.partial = c.take<float>((size_t)dw_gx(nb) * nb * nb * PARTIAL);
  s.ppart = c.take<float>((size_t)2 * NSM * 4 * Hp);
  s.bytes = c.off;
  return s;
}

template <bool EDGE>
int tile_grid(int B, int N, int cap) {
  const long long tiles = EDGE ? (long long)B * ((KMAX * N + TILE - 1) / TILE) : ((long long)B * N + TILE - 1) / TILE;
  return (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
}

int configure() {
  static DeviceOnce once;
  const int dev = once.pending();
  if (dev < 0) return 0;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k_g_lin<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LinSmem)))) return (int)e;
  if ((e = cudaFuncSetAttribute(k_g_lin<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LinSmem)))) return (int)e;
  if ((e = cudaFuncSetAttribute(k_g_dw<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DwSmem)))) return (int)e;
  if ((e = cudaFuncSetAttribute(k_g_dw<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DwSmem)))) return (int)e;
  once.done(dev);
  return 0;
}

template <bool EDGE>
int lin(LinArgs a, cudaStream_t st) {
  const dim3 grid(tile_grid<EDGE>(a.B, a.N, 2 * NSM / a.nb > 16 ? 2 * NSM / a.nb : 16), a.nb);
  k_g_lin<EDGE><<<grid, NT, sizeof(LinSmem), st>>>(a);
  PILE_CHECK_LAUNCH();
  return 0;
}

template <bool EDGE>
int mask(const float* g, int g_gather, const float* ymask, const Csr& csr, float* Gm, int B, int N, int Hp, cudaStream_t st) {
  if (EDGE) {
    const dim3 grid((unsigned)(((long long)KMAX * N * (Hp / 4) + 255) / 256), B < 65535 ? B : 65535);
    k_g_mask<true><<<grid, 256, 0, st>>>(g, g_gather, ymask, csr.rowptr, csr.row, Gm, B, N, Hp);
  } else {
    const long long total = (long long)B * N * (Hp / 4);
    k_g_mask<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, 0, ymask, nullptr, nullptr, Gm, B, N, Hp);
  }
  PILE_CHECK_LAUNCH();
  return 0;
}

LinArgs lin_base(int B, int N, int nb, const Csr& csr) {
  LinArgs a{};
  a.B = B; a.N = N; a.nb = nb;
  a.rowptr = csr.rowptr; a.col = csr.col; a.row = csr.row;
  return a;
}

// weight gradients of one layer from the masked upstream gradient Gm: partial sums + fixed-order reduction
template <bool EDGE>
int wgrad(int B, int N, int nb, int H, const Csr& csr, const float* Gm, int nsrc, const Src* src, const float* x8,
          const float* dens, float* partial, FinishArgs f, cudaStream_t st) {
  DwArgs a{};
  a.B = B; a.N = N; a.nb = nb;
  a.rowptr = csr.rowptr; a.col = csr.col; a.row = csr.row;
  a.G = Gm; a.nsrc = nsrc;
  for (int s = 0; s < nsrc; ++s) a.src[s] = src[s];
  a.x8 = x8; a.dens = dens; a.partial = partial;
  const int gx = tile_grid<EDGE>(B, N, dw_gx(nb));
  k_g_dw<EDGE><<<dim3(gx, nb * nb), NT, sizeof(DwSmem), st>>>(a);
  PILE_CHECK_LAUNCH();
  f.partial = partial; f.gx = gx; f.nb = nb; f.H = H; f.nsrc = nsrc;
  k_g_finish<<<64, 256, 0, st>>>(f);
  PILE_CHECK_LAUNCH();
  return 0;
}

// offsets (floats) of the 18 gradient tensors inside the gradient buffer = reference state_dict order and shapes
struct GradOff {
  long long pe0_w, pe0_b, pe1_w, pe1_b, re0_w, re0_b, re1_w, re1_b, re2_w, re2_b, pp_w, pp_b, rp_w, rp_b, v0_w, v0_b, v1_w, v1_b,
      total;
};
GradOff grad_offsets(int H) {
  GradOff g;
  long long o = 0;
  auto take = [&](long long n) { const long long r = o; o += n; return r; };
  g.pe0_w = take(H * 5LL); g.pe0_b = take(H);
  g.pe1_w = take((long long)H * H); g.pe1_b = take(H);
  g.re0_w = take(H * 6LL); g.re0_b = take(H);
  g.re1_w = take((long long)H * H); g.re1_b = take(H);
  g.re2_w = take((long long)H * H); g.re2_b = take(H);
  g.pp_w = take((long long)H * (2 * H + 1)); g.pp_b = take(H);
  g.rp_w = take((long long)H * (3 * H + 1)); g.rp_b = take(H);
  g.v0_w = take((long long)H * H); g.v0_b = take(H);
  g.v1_w = take(3LL * H); g.v1_b = take(3);
  g.total = o;
  return g;
}

// hoisted: inference only -- the relation propagator in its regrouped form (k_g_agg_hoisted); the tape then holds the
// relation lists and the particle-side arrays but NOT the relation-side layer inputs M[p] a backward pass would need
// (M[0] = Ce, the head of M[1] / M[2] = Pr / Ps)
int forward_body(const float* wpack, int Hp, const float* attr, const float* dens, const float* s_cur,
                 const float* s_delta, int B, int N, const Tape& t, float* s_pred, cudaStream_t st, bool hoisted = false) {
  int e = 0;
  const int nb = Hp / BW;
  const long long R = (long long)B * N;
  k_g_node_in<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(s_delta, attr, dens, t.X0, B, N);
  PILE_CHECK_LAUNCH();
  auto W = [&](int slot) { return wpack + slot_offset(slot, Hp); };
  // Inference with the tensor-core GEMM engine selected (pile_set_tensor_cores != 0) and enough relation rows to fill the
  // GPU: the three wide relation-side layers run on tcgen05 (general_tc.cu).  Their bf16 hi / lo weight images are built here,
  // into the tape space behind P_r (the head of M[1]; M[1] / M[2] hold nothing else in an inference step).
  const size_t E_cap = (size_t)B * KMAX * N;
  enum { I_RE1, I_RE2, I_E, I_PE1, I_R, I_S, I_P, I_A, I_V0, NIMG };
  const bool tc = hoisted && g_use_tensor_cores != 0 && E_cap >= 4096 &&
                  E_cap * Hp >= (size_t)R * Hp + NIMG * tc_image_floats(Hp);
  float* img[NIMG] = {};
  if (tc) {
    const int wslot[NIMG] = {W_RE1T, W_RE2T, W_ET, W_PE1T, W_RT, W_ST, W_PT, W_AT, W_V0T};
    long long off[NIMG];
    for (int i = 0; i < NIMG; ++i) {
      img[i] = t.M[1] + (size_t)R * Hp + i * tc_image_floats(Hp);
      off[i] = slot_offset(wslot[i], Hp);
    }
    if ((e = launch_tc_images(wpack, off, NIMG, Hp, img[0], st))) return e;
  }
  {  // particle encoder
    LinArgs a = lin_base(B, N, nb, t.csr);
    a.x8 = t.X0; a.w8 = W(W_PE0T); a.bias = W(B_PE0); a.y = t.H0; a.relu = 1;
    if ((e = lin<false>(a, st))) return e;
    if (tc) {
      if ((e = launch_lin_tc_node(t.H0, img[I_PE1], nullptr, nullptr, W(B_PE1), nullptr, nullptr, nullptr, 1, t.P, B, N, Hp, st)))
        return e;
    } else {
      a = lin_base(B, N, nb, t.csr);
      a.nsrc = 1; a.src[0] = {t.H0, W(W_PE1T), 0}; a.bias = W(B_PE1); a.y = t.P; a.relu = 1;
      if ((e = lin<false>(a, st))) return e;
    }
  }
  {  // relation encoder
    LinArgs a = lin_base(B, N, nb, t.csr);
    if (tc) {
      // RE0 (K = 8) is computed inside the RE1 kernel while it builds its A tiles: R1 never exists in memory
      if ((e = launch_lin_tc_edge(nullptr, img[I_RE1], W(B_RE1), nullptr, nullptr, 1, t.R2, t.csr.rowptr, B, N, Hp, st, t.Y0,
                                  W(W_RE0T), W(B_RE0))))
        return e;
      if ((e = launch_lin_tc_edge(t.R2, img[I_RE2], W(B_RE2), nullptr, nullptr, 1, t.R3, t.csr.rowptr, B, N, Hp, st))) return e;
    } else {
      a.x8 = t.Y0; a.w8 = W(W_RE0T); a.bias = W(B_RE0); a.y = t.R1; a.relu = 1;
      if ((e = lin<true>(a, st))) return e;
      a = lin_base(B, N, nb, t.csr);
      a.nsrc = 1; a.src[0] = {t.R1, W(W_RE1T), 0}; a.bias = W(B_RE1); a.y = t.R2; a.relu = 1;
      if ((e = lin<true>(a, st))) return e;
      a.src[0] = {t.R2, W(W_RE2T), 0}; a.bias = W(B_RE2); a.y = t.R3;
      if ((e = lin<true>(a, st))) return e;
    }
  }
  const unsigned node4 = (unsigned)((R * (Hp / 4) + 255) / 256);
  if (hoisted) {          // Ce = W_e r3 + w_d d + b
    if (tc) {
      if ((e = launch_lin_tc_edge(t.R3, img[I_E], W(B_RP), W(WD_RP), dens, 0, t.M[0], t.csr.rowptr, B, N, Hp, st))) return e;
    } else {
      LinArgs a = lin_base(B, N, nb, t.csr);
      a.nsrc = 1; a.src[0] = {t.R3, W(W_ET), 0}; a.dens = dens; a.wd = W(WD_RP); a.bias = W(B_RP); a.y = t.M[0]; a.relu = 0;
      if ((e = lin<true>(a, st))) return e;
    }
  }
  for (int p = 0; p < PSTEP; ++p) {
    const float* eff_in = p == 0 ? t.P : t.eff[p - 1];
    LinArgs a = lin_base(B, N, nb, t.csr);
    if (hoisted) {
      if (tc) {
        if ((e = launch_lin_tc_node(eff_in, img[I_R], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, t.M[1], B, N, Hp, st)))
          return e;
        if ((e = launch_lin_tc_node(eff_in, img[I_S], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, t.M[2], B, N, Hp, st)))
          return e;
      } else {
        a.nsrc = 1; a.src[0] = {eff_in, W(W_RT), 0}; a.y = t.M[1]; a.relu = 0;
        if ((e = lin<false>(a, st))) return e;
        a.src[0] = {eff_in, W(W_ST), 0}; a.y = t.M[2];
        if ((e = lin<false>(a, st))) return e;
      }
      k_g_agg_hoisted<<<node4, 256, 0, st>>>(t.csr.rowptr, t.csr.col, t.M[0], t.M[1], t.M[2], t.agg[p], B, N, Hp);
      PILE_CHECK_LAUNCH();
      if (tc) {
        if ((e = launch_lin_tc_node(t.P, img[I_P], t.agg[p], img[I_A], W(B_PP), W(WD_PP), dens, eff_in, 1, t.eff[p], B, N, Hp, st)))
          return e;
        continue;
      }
      a = lin_base(B, N, nb, t.csr);
      a.nsrc = 2;
      a.src[0] = {t.P, W(W_PT), 0};
      a.src[1] = {t.agg[p], W(W_AT), 0};
      a.dens = dens; a.wd = W(WD_PP); a.bias = W(B_PP); a.res = eff_in; a.y = t.eff[p]; a.relu = 1;
      if ((e = lin<false>(a, st))) return e;
      continue;
    }
    a.nsrc = 3;
    a.src[0] = {t.R3, W(W_ET), 0};
    a.src[1] = {eff_in, W(W_RT), 1};
    a.src[2] = {eff_in, W(W_ST), 2};
    a.dens = dens; a.wd = W(WD_RP); a.bias = W(B_RP); a.y = t.M[p]; a.relu = 1;
    if ((e = lin<true>(a, st))) return e;
    k_g_gather_nodes<<<node4, 256, 0, st>>>(t.csr.rowptr, nullptr, nullptr, nullptr, t.M[p], nullptr, t.agg[p], 0, B, N, Hp);
    PILE_CHECK_LAUNCH();
    a = lin_base(B, N, nb, t.csr);
    a.nsrc = 2;
    a.src[0] = {t.P, W(W_PT), 0};
    a.src[1] = {t.agg[p], W(W_AT), 0};
    a.dens = dens; a.wd = W(WD_PP); a.bias = W(B_PP); a.res = eff_in; a.y = t.eff[p]; a.relu = 1;
    if ((e = lin<false>(a, st))) return e;
  }
  if (tc) {
    if ((e = launch_lin_tc_node(t.eff[PSTEP - 1], img[I_V0], nullptr, nullptr, W(B_V0), nullptr, nullptr, nullptr, 1, t.Q, B, N, Hp, st)))
      return e;
  } else {
    LinArgs a = lin_base(B, N, nb, t.csr);
    a.nsrc = 1; a.src[0] = {t.eff[PSTEP - 1], W(W_V0T), 0}; a.bias = W(B_V0); a.y = t.Q; a.relu = 1;
    if ((e = lin<false>(a, st))) return e;
  }
  k_g_predict<<<(unsigned)((R + 127) / 128), 128, Hp * 4 * sizeof(float), st>>>(t.Q, W(W_V1T), W(B_V1), s_cur, s_pred, B, N, Hp);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace general

using namespace general;

static inline bool bad_width(int H) { return H < 1 || H > MAXNB * BW; }

long long general_wpack_slot_offset(int slot, int H) {
  if (bad_width(H) || slot < 0 || slot > NSLOTS) return -1;
  return slot_offset(slot, nblocks(H) * BW);
}
long long general_tape_bytes(int B, int N, int H) {
  return bad_width(H) ? -1 : (long long)carve_tape(nullptr, B, N, nblocks(H) * BW).bytes;
}
long long general_bwd_scratch_bytes(int B, int N, int H) {
  return bad_width(H) ? -1 : (long long)carve_bwd(nullptr, B, N, nblocks(H) * BW).bytes;
}
long long general_grad_offset(int tensor_index, int H) {
  if (bad_width(H)) return -1;
  const GradOff g = grad_offsets(H);
  const long long t[19] = {g.pe0_w, g.pe0_b, g.pe1_w, g.pe1_b, g.re0_w, g.re0_b, g.re1_w, g.re1_b, g.re2_w, g.re2_b,
                           g.pp_w, g.pp_b, g.rp_w, g.rp_b, g.v0_w, g.v0_b, g.v1_w, g.v1_b, g.total};
  return (tensor_index < 0 || tensor_index > 18) ? -1 : t[tensor_index];
}

int general_relations_view(void* tape, int B, int N, int H, int** rowptr, int** col, int** row) {
  if (bad_width(H)) return (int)cudaErrorInvalidValue;
  const Tape t = carve_tape(tape, B, N, nblocks(H) * BW);
  *rowptr = t.csr.rowptr; *col = t.csr.col; *row = t.csr.row;
  return 0;
}

int launch_general_forward(const float* wpack, int H, const float* attr, const float* dens, const int* particle_nums,
                           const float* s_cur, const float* s_delta, float adj_thresh, int B, int N, void* tape,
                           float* s_pred, cudaStream_t st, bool hoisted) {
  if (bad_width(H)) return (int)cudaErrorInvalidValue;
  int e = configure();
  if (e) return e;
  const int Hp = nblocks(H) * BW;
  const Tape t = carve_tape(tape, B, N, Hp);
  PushCam none{};
  e = launch_nbr_search(s_cur, (long long)N * 3, s_delta, nullptr, 0, none, nullptr, particle_nums, B, N,
                        adj_thresh * adj_thresh, t.csr, st, attr, dens, t.Y0);
  if (e) return e;
  return forward_body(wpack, Hp, attr, dens, s_cur, s_delta, B, N, t, s_pred, st, hoisted);
}

// the same step on caller-provided relation lists (receiver-grouped CSR, any number of relations per receiver as long
// as a sample has at most KMAX * N in total): the "Rr / Rs" entry of PropModuleDiffDen.forward (gnn_dyn.py:147)
int launch_general_forward_relations(const float* wpack, int H, const float* attr, const float* dens, const float* s_cur,
                                     const float* s_delta, const int* rowptr, const int* col, const int* row, int B,
                                     int N, void* tape, float* s_pred, cudaStream_t st) {
  if (bad_width(H)) return (int)cudaErrorInvalidValue;
  int e = configure();
  if (e) return e;
  const int Hp = nblocks(H) * BW;
  const Tape t = carve_tape(tape, B, N, Hp);
  const size_t E = (size_t)B * KMAX * N;
  cudaError_t ce;
  if ((ce = cudaMemcpyAsync(t.csr.rowptr, rowptr, sizeof(int) * (size_t)B * (N + 1), cudaMemcpyDeviceToDevice, st))) return (int)ce;
  if ((ce = cudaMemcpyAsync(t.csr.col, col, sizeof(int) * E, cudaMemcpyDeviceToDevice, st))) return (int)ce;
  if ((ce = cudaMemcpyAsync(t.csr.row, row, sizeof(int) * E, cudaMemcpyDeviceToDevice, st))) return (int)ce;
  if ((e = launch_transpose_relations(t.csr, B, N, st))) return e;
  const dim3 fgrid((KMAX * N + 255) / 256, B < 65535 ? B : 65535);
  k_g_edge_in<<<fgrid, 256, 0, st>>>(attr, dens, s_cur, t.csr.rowptr, t.csr.col, t.csr.row, t.Y0, B, N);
  PILE_CHECK_LAUNCH();
  return forward_body(wpack, Hp, attr, dens, s_cur, s_delta, B, N, t, s_pred, st);
}

// grads == nullptr: input gradients only (the planner's action refinement, planners.py:674)
int launch_general_backward(const float* wpack, int H, const float* dens, void* tape, int B, int N, const float* g_pred,
                            float* g_s_cur, float* g_s_delta, float* grads, void* scratch, cudaStream_t st) {
  if (bad_width(H)) return (int)cudaErrorInvalidValue;
  int e = configure();
  if (e) return e;
  const int nb = nblocks(H), Hp = nb * BW;
  const Tape t = carve_tape(tape, B, N, Hp);
  const BwdScratch s = carve_bwd(scratch, B, N, Hp);
  const GradOff go = grad_offsets(H);
  const long long R = (long long)B * N;
  auto W = [&](int slot) { return wpack + slot_offset(slot, Hp); };
  const bool wg = grads != nullptr;
  const unsigned node4 = (unsigned)((R * (Hp / 4) + 255) / 256);

  // dX = Gm W (W in [out][in] layout = the [in' = out][out' = in] operand of the forward kernel)
  auto dx = [&](bool edge, const float* Gm, const float* w, float* out, int accumulate) {
    LinArgs a = lin_base(B, N, nb, t.csr);
    a.nsrc = 1; a.src[0] = {Gm, w, 0}; a.y = out; a.accumulate = accumulate;
    return edge ? lin<true>(a, st) : lin<false>(a, st);
  };

  // predictor output layer
  {
    const int grid = (int)((R + PB_ROWS - 1) / PB_ROWS < 2 * NSM ? (R + PB_ROWS - 1) / PB_ROWS : 2 * NSM);
    k_g_predict_bwd<<<grid, PB_ROWS * Hp, (4 + PB_ROWS * 4) * Hp * sizeof(float), st>>>(g_pred, t.Q, W(W_V1), s.gA, s.ppart, B, N, Hp);
    PILE_CHECK_LAUNCH();
    if (wg) {
      k_g_predict_finish<<<4, 256, 0, st>>>(s.ppart, grid, Hp, H, grads + go.v1_w, grads + go.v1_b);
      PILE_CHECK_LAUNCH();
    }
  }
  {  // V0
    if ((e = mask<false>(s.gA, 0, t.Q, t.csr, s.gmN, B, N, Hp, st))) return e;
    if ((e = dx(false, s.gmN, W(W_V0), s.gEff, 0))) return e;
    if (wg) {
      const Src src[1] = {{t.eff[PSTEP - 1], nullptr, 0}};
      FinishArgs f{};
      f.src[0] = {grads + go.v0_w, H, 0}; f.dbias = grads + go.v0_b;
      if ((e = wgrad<false>(B, N, nb, H, t.csr, s.gmN, 1, src, nullptr, nullptr, s.partial, f, st))) return e;
    }
  }
  const int PPLD = 2 * H + 1, RPLD = 3 * H + 1;
  for (int p = PSTEP - 1; p >= 0; --p) {
    const float* eff_in = p == 0 ? t.P : t.eff[p - 1];
    {  // particle propagator: Gm = gEff masked by eff_{p+1} (also the gradient of the residual path)
      if ((e = mask<false>(s.gEff, 0, t.eff[p], t.csr, s.gmN, B, N, Hp, st))) return e;
      if ((e = dx(false, s.gmN, W(W_P), s.gP, p == PSTEP - 1 ? 0 : 1))) return e;
      if ((e = dx(false, s.gmN, W(W_A), s.gAgg, 0))) return e;
      if (wg) {
        const Src src[2] = {{t.P, nullptr, 0}, {t.agg[p], nullptr, 0}};
        FinishArgs f{};
        f.src[0] = {grads + go.pp_w, PPLD, 0}; f.src[1] = {grads + go.pp_w, PPLD, H};
        f.dwd = grads + go.pp_w + 2 * H; f.ld_wd = PPLD; f.dbias = grads + go.pp_b;
        if ((e = wgrad<false>(B, N, nb, H, t.csr, s.gmN, 2, src, nullptr, dens, s.partial, f, st))) return e;
      }
    }
    {  // relation propagator: Gm(e) = gAgg[recv e] masked by M_p(e)
      if ((e = mask<true>(s.gAgg, 1, t.M[p], t.csr, s.gmE, B, N, Hp, st))) return e;
      if ((e = dx(true, s.gmE, W(W_E), s.gR3, p == PSTEP - 1 ? 0 : 1))) return e;
      if ((e = dx(true, s.gmE, W(W_R), s.dZr, 0))) return e;
      if ((e = dx(true, s.gmE, W(W_S), s.dZs, 0))) return e;
      if (wg) {
        const Src src[3] = {{t.R3, nullptr, 0}, {eff_in, nullptr, 1}, {eff_in, nullptr, 2}};
        FinishArgs f{};
        f.src[0] = {grads + go.rp_w, RPLD, 0}; f.src[1] = {grads + go.rp_w, RPLD, H}; f.src[2] = {grads + go.rp_w, RPLD, 2 * H};
        f.dwd = grads + go.rp_w + 3 * H; f.ld_wd = RPLD; f.dbias = grads + go.rp_b;
        if ((e = wgrad<true>(B, N, nb, H, t.csr, s.gmE, 3, src, nullptr, dens, s.partial, f, st))) return e;
      }
    }
    // d/d eff_p = masked rows (residual) + receiver-side + sender-side relation terms; p == 0: eff_0 is P
    k_g_gather_nodes<<<node4, 256, 0, st>>>(t.csr.rowptr, t.csr.trowptr, t.csr.tedge, s.gmN, s.dZr, s.dZs,
                                            p == 0 ? s.gP : s.gEff, p == 0 ? 1 : 0, B, N, Hp);
    PILE_CHECK_LAUNCH();
  }
  {  // particle encoder
    if ((e = mask<false>(s.gP, 0, t.P, t.csr, s.gmN, B, N, Hp, st))) return e;
    if ((e = dx(false, s.gmN, W(W_PE1), s.gH0, 0))) return e;
    if (wg) {
      const Src src[1] = {{t.H0, nullptr, 0}};
      FinishArgs f{};
      f.src[0] = {grads + go.pe1_w, H, 0}; f.dbias = grads + go.pe1_b;
      if ((e = wgrad<false>(B, N, nb, H, t.csr, s.gmN, 1, src, nullptr, nullptr, s.partial, f, st))) return e;
    }
    if ((e = mask<false>(s.gH0, 0, t.H0, t.csr, s.gmN, B, N, Hp, st))) return e;
    k_g_dx8<<<(unsigned)((R + 127) / 128), 128, Hp * 8 * sizeof(float), st>>>(s.gmN, W(W_PE0), nullptr, s.dX0, B, N, Hp, 0);
    PILE_CHECK_LAUNCH();
    if (wg) {
      FinishArgs f{};
      f.dest8 = grads + go.pe0_w; f.ld8 = 5; f.k8 = 5; f.dbias = grads + go.pe0_b;
      if ((e = wgrad<false>(B, N, nb, H, t.csr, s.gmN, 0, nullptr, t.X0, nullptr, s.partial, f, st))) return e;
    }
  }
  {  // relation encoder
    if ((e = mask<true>(s.gR3, 0, t.R3, t.csr, s.gmE, B, N, Hp, st))) return e;
    if ((e = dx(true, s.gmE, W(W_RE2), s.gR2, 0))) return e;
    if (wg) {
      const Src src[1] = {{t.R2, nullptr, 0}};
      FinishArgs f{};
      f.src[0] = {grads + go.re2_w, H, 0}; f.dbias = grads + go.re2_b;
      if ((e = wgrad<true>(B, N, nb, H, t.csr, s.gmE, 1, src, nullptr, nullptr, s.partial, f, st))) return e;
    }
    if ((e = mask<true>(s.gR2, 0, t.R2, t.csr, s.gmE, B, N, Hp, st))) return e;
    if ((e = dx(true, s.gmE, W(W_RE1), s.dZr, 0))) return e;          // gR1 reuses the dZr buffer
    if (wg) {
      const Src src[1] = {{t.R1, nullptr, 0}};
      FinishArgs f{};
      f.src[0] = {grads + go.re1_w, H, 0}; f.dbias = grads + go.re1_b;
      if ((e = wgrad<true>(B, N, nb, H, t.csr, s.gmE, 1, src, nullptr, nullptr, s.partial, f, st))) return e;
    }
    if ((e = mask<true>(s.dZr, 0, t.R1, t.csr, s.gmE, B, N, Hp, st))) return e;
    const dim3 egrid((KMAX * N + 127) / 128, B < 65535 ? B : 65535);
    k_g_dx8<<<egrid, 128, Hp * 8 * sizeof(float), st>>>(s.gmE, W(W_RE0), t.csr.rowptr, s.dY0, B, N, Hp, 1);
    PILE_CHECK_LAUNCH();
    if (wg) {
      FinishArgs f{};
      f.dest8 = grads + go.re0_w; f.ld8 = 6; f.k8 = 6; f.dbias = grads + go.re0_b;
      if ((e = wgrad<true>(B, N, nb, H, t.csr, s.gmE, 0, nullptr, t.Y0, nullptr, s.partial, f, st))) return e;
    }
  }
  k_g_positions<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(t.csr.rowptr, t.csr.trowptr, t.csr.tedge, s.dY0, s.dX0, g_pred,
                                                              g_s_cur, g_s_delta, B, N);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
