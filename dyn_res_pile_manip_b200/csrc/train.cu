// Training path of the propagation network (reference train/train_gnn_dyn.py:150-199 drives
// PropNetDiffDenModel.predict_one_step under autograd and steps Adam on all 18 tensors): forward that keeps every
// layer input, backward that returns d/ds_cur, d/ds_delta AND the weight gradients of the nine linear layers.
//
// The planner path (fwd.cu / *_tc.cu, bwd*.cu) is dgrad-only and replays ReLUs from sign bits; training also needs
// dW = G^T X for every layer, i.e. the layer inputs X.  This file therefore runs the network in the reference's
// un-hoisted form with every activation materialised (fp32, row-major [rows, 64]):
//
//   forward   X0 -PE0-> H0 -PE1-> P = eff_0          Y0 -RE0-> R1 -RE1-> R2 -RE2-> R3
//             p = 0..2:  M_p = ReLU([R3 | eff_p[recv] | eff_p[send] | d] W_rp^T + b)      (edge rows)
//                        agg_p = segment-sum of M_p over receivers
//                        eff_{p+1} = ReLU([P | agg_p | d] W_pp^T + b + eff_p)              (node rows)
//             Q = ReLU(eff_3 V0^T + b),  s_pred = Q V1^T + b + s_cur
//   backward  one generic tile kernel per linear layer: G = upstream gradient (optionally gathered by receiver)
//             masked by the layer's own output > 0; per source block  dX = G W  and  dW += G^T X ; bias / density
//             column sums.  Every CTA keeps its dW partial sums in registers over all its tiles and writes them
//             once; k_tl_finish adds the CTAs' partials in a fixed order (deterministic, no float atomics).
//             Scatters are gathers over the CSR and its sender-major transpose.
//
// All GEMMs are FP32 on the CUDA cores (128-row tiles, 8 warps x (4 rows x 8 columns) per lane, common.cuh): the
// training batches of the reference are a few thousand particle rows, far from the regime the tcgen05 engine of the
// planner path is built for, and fp32 keeps the weight gradients within 1e-4 of autograd.
#include "common.cuh"
#include "kernels.h"

namespace pile {

namespace {

constexpr int MAXSRC = 3;

// one 64-wide input block of a linear layer
struct TlSrc {
  const float* x;      // [*, 64] rows (node rows when gather != 0)
  const float* w;      // forward: W^T [in 64][out 64];  backward: W [out 64][in 64]
  int gather;          // edge kernels: 0 = the edge's own row, 1 = row of its receiver particle, 2 = of its sender
  float* dx;           // backward: [rows, 64] d/dx of this block (nullptr: not needed)
  int dx_accumulate;   // backward: += instead of =
};

struct TlArgs {
  int B, N;
  const int* rowptr;   // edge kernels
  const int* col;      // sender of an edge
  const int* row;      // receiver of an edge
  int nsrc;
  TlSrc src[MAXSRC];
  const float* x8;     // [rows, 8] narrow input block (nullptr: none)
  const float* w8;     // forward: W8^T [8][64]; backward: W8 [64][8]
  float* dx8;          // backward: [rows, 8]
  const float* dens;   // [B] (nullptr: no density column)
  const float* wd;     // [64] density column of the layer
  const float* bias;   // [64]
  const float* res;    // forward: residual [rows, 64] added before the ReLU (nullptr: none)
  float* y;            // forward: output [rows, 64]
  int relu;
  // backward
  const float* g;      // upstream gradient [rows, 64] (node rows when g_gather)
  int g_gather;        // edge kernels: 1 = take the receiver particle's row
  const float* ymask;  // [rows, 64] the layer's own forward output: G *= (ymask > 0) (nullptr: no mask)
  float* g_out;        // optional: the masked G rows, [rows, 64]
  float* partial;      // [gridDim.x][TL_PARTIAL] per-CTA weight-gradient partial sums
};

constexpr int TL_PARTIAL = MAXSRC * H * H + H * 8 + H + H;     // dW blocks | dW8 | dbias | dwd

struct TlTile {
  long long row0;      // first absolute row (node row or edge slot)
  int nrows;
  int b;               // sample of an edge tile
};

template <bool EDGE>
__device__ __forceinline__ int tl_num_tiles(const TlArgs& a) {
  if (EDGE) return a.B * ((KMAX * a.N + TILE - 1) / TILE);
  return (int)(((long long)a.B * a.N + TILE - 1) / TILE);
}

template <bool EDGE>
__device__ __forceinline__ TlTile tl_tile(const TlArgs& a, int t) {
  TlTile q;
  if (EDGE) {
    const int tps = (KMAX * a.N + TILE - 1) / TILE;
    q.b = t / tps;
    const int e0 = (t - q.b * tps) * TILE;
    const int ne = a.rowptr[(long long)q.b * (a.N + 1) + a.N];
    q.nrows = min(TILE, ne - e0);
    q.row0 = (long long)q.b * KMAX * a.N + e0;
  } else {
    q.b = 0;
    q.row0 = (long long)t * TILE;
    q.nrows = (int)min((long long)TILE, (long long)a.B * a.N - q.row0);
  }
  return q;
}

// absolute source row of tile row r for a block with the given gather mode
template <bool EDGE>
__device__ __forceinline__ long long tl_src_row(const TlArgs& a, const TlTile& q, int r, int gather) {
  if (!EDGE || gather == 0) return q.row0 + r;
  const int* ix = gather == 1 ? a.row : a.col;
  return (long long)q.b * a.N + ix[q.row0 + r];
}

// rows of a [*, 64] array -> shared tile [TILE][LDA] (half-warp per row, zero fill past nrows)
template <bool EDGE>
__device__ __forceinline__ void tl_load_rows(const TlArgs& a, const TlTile& q, const float* __restrict__ x, int gather,
                                             float* __restrict__ A) {
  const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
  for (int r = hw; r < TILE; r += NT / 16) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < q.nrows) v = ld4(x + tl_src_row<EDGE>(a, q, r, gather) * H + 4 * l16);
    st4(A + r * LDA + 4 * l16, v);
  }
}

__device__ __forceinline__ void tl_load_rows8(const TlTile& q, const float* __restrict__ x8, float* __restrict__ A8) {
  for (int idx = threadIdx.x; idx < TILE * 2; idx += NT) {
    const int r = idx >> 1, h = idx & 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < q.nrows) v = ld4(x8 + (q.row0 + r) * 8 + 4 * h);
    st4(A8 + r * LDX + 4 * h, v);
  }
}

template <bool EDGE>
__device__ __forceinline__ float tl_dens(const TlArgs& a, const TlTile& q, int r) {
  if (a.dens == nullptr || r >= q.nrows) return 0.f;
  const int b = EDGE ? q.b : (int)((q.row0 + r) / a.N);
  return a.dens[b] / 5000.f;           // gnn_dyn.py:158
}

// ------------------------------------------------------------------------------------------------
// forward of one linear layer: y = act(sum_s x_s W_s^T + x8 W8^T + d wd + bias + res)
// ------------------------------------------------------------------------------------------------
struct TlFwdSmem {
  float w[MAXSRC][H * H];
  float w8[8 * H];
  float bias[H], wd[H];
  float A[TILE * LDA];
  float A8[TILE * LDX];
};

template <bool EDGE>
__global__ void __launch_bounds__(NT, 2) k_tl_fwd(TlArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TlFwdSmem& S = *reinterpret_cast<TlFwdSmem*>(smem_raw);
  for (int s = 0; s < a.nsrc; ++s) load_block(S.w[s], a.src[s].w, H * H);
  if (a.x8) load_block(S.w8, a.w8, 8 * H);
  if (threadIdx.x < H) {
    S.bias[threadIdx.x] = a.bias[threadIdx.x];
    S.wd[threadIdx.x] = a.wd ? a.wd[threadIdx.x] : 0.f;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c0 = warp * 8;
  const int ntiles = tl_num_tiles<EDGE>(a);
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const TlTile q = tl_tile<EDGE>(a, t);
    if (q.nrows <= 0) continue;          // CTA-uniform
    __syncthreads();
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float d = tl_dens<EDGE>(a, q, lane + 32 * i);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(d, S.wd[c0 + j], S.bias[c0 + j]);
    }
    if (a.res) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = lane + 32 * i;
        if (r < q.nrows) {
          const float4 u = ld4(a.res + (q.row0 + r) * H + c0), v = ld4(a.res + (q.row0 + r) * H + c0 + 4);
          acc[i][0] += u.x; acc[i][1] += u.y; acc[i][2] += u.z; acc[i][3] += u.w;
          acc[i][4] += v.x; acc[i][5] += v.y; acc[i][6] += v.z; acc[i][7] += v.w;
        }
      }
    }
    for (int s = 0; s < a.nsrc; ++s) {
      if (s) __syncthreads();
      tl_load_rows<EDGE>(a, q, a.src[s].x, a.src[s].gather, S.A);
      __syncthreads();
      gemm_rows4x8<H>(S.A, LDA, S.w[s], lane, c0, acc);
    }
    if (a.x8) {
      tl_load_rows8(q, a.x8, S.A8);
      __syncthreads();
      gemm_rows4x8<8>(S.A8, LDX, S.w8, lane, c0, acc);
    }
    if (a.relu) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaxf(acc[i][j], 0.f);
    }
    acc_to_global(a.y, q.row0, q.nrows, lane, c0, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// backward of one linear layer
// ------------------------------------------------------------------------------------------------
struct TlBwdSmem {
  float w[MAXSRC][H * H];      // [out][in]
  float w8[H * 8];             // [out][8]
  float G[TILE * LDA];
  float X[TILE * LDA];
  float X8[TILE * LDX];
  float dn[TILE];
};

template <bool EDGE>
__global__ void __launch_bounds__(NT, 1) k_tl_bwd(TlArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TlBwdSmem& S = *reinterpret_cast<TlBwdSmem*>(smem_raw);
  for (int s = 0; s < a.nsrc; ++s) load_block(S.w[s], a.src[s].w, H * H);
  if (a.x8) load_block(S.w8, a.w8, H * 8);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c0 = warp * 8;
  const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
  const int o0 = hw * 4, k0 = l16 * 4;           // this thread's 4 x 4 block of every dW
  float dw[MAXSRC][4][4];
#pragma unroll
  for (int s = 0; s < MAXSRC; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dw[s][i][j] = 0.f;
  float dw8[2] = {0.f, 0.f};                     // thread t: out = t / 4, in = 2 * (t % 4) + {0, 1}
  float dbias = 0.f, dwd = 0.f;                  // threads < 64: column threadIdx.x
  const int ntiles = tl_num_tiles<EDGE>(a);
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const TlTile q = tl_tile<EDGE>(a, t);
    if (q.nrows <= 0) continue;          // CTA-uniform
    __syncthreads();
    // G tile: upstream gradient (gathered by receiver for relation rows), masked by the layer's own output
    for (int r = hw; r < TILE; r += NT / 16) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < q.nrows) {
        v = ld4(a.g + tl_src_row<EDGE>(a, q, r, a.g_gather) * H + 4 * l16);
        if (a.ymask) {
          const float4 y = ld4(a.ymask + (q.row0 + r) * H + 4 * l16);
          v.x = y.x > 0.f ? v.x : 0.f; v.y = y.y > 0.f ? v.y : 0.f; v.z = y.z > 0.f ? v.z : 0.f; v.w = y.w > 0.f ? v.w : 0.f;
        }
        if (a.g_out) st4(a.g_out + (q.row0 + r) * H + 4 * l16, v);
      }
      st4(S.G + r * LDA + 4 * l16, v);
    }
    if (threadIdx.x < TILE) S.dn[threadIdx.x] = tl_dens<EDGE>(a, q, threadIdx.x);
    __syncthreads();
    if (threadIdx.x < H) {
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
      if (a.src[s].dx) {                       // dX = G W_s
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        gemm_rows4x8<H>(S.G, LDA, S.w[s], lane, c0, acc);
        if (a.src[s].dx_accumulate) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = lane + 32 * i;
            if (r < q.nrows) {
              float* p = a.src[s].dx + (q.row0 + r) * H + c0;
              float4 u = ld4(p), v = ld4(p + 4);
              u.x += acc[i][0]; u.y += acc[i][1]; u.z += acc[i][2]; u.w += acc[i][3];
              v.x += acc[i][4]; v.y += acc[i][5]; v.z += acc[i][6]; v.w += acc[i][7];
              st4(p, u); st4(p + 4, v);
            }
          }
        } else {
          acc_to_global(a.src[s].dx, q.row0, q.nrows, lane, c0, acc);
        }
      }
      // dW_s += G^T X_s
      if (s) __syncthreads();
      tl_load_rows<EDGE>(a, q, a.src[s].x, a.src[s].gather, S.X);
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
    if (a.x8) {
      tl_load_rows8(q, a.x8, S.X8);
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
      if (a.dx8 && (int)threadIdx.x < q.nrows) {       // dX8 = G W8: one row per thread
        const int r = threadIdx.x;
        float o8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < H; ++k) {
          const float gv = S.G[r * LDA + k];
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] = fmaf(gv, S.w8[k * 8 + j], o8[j]);
        }
        st4(a.dx8 + (q.row0 + r) * 8, make_float4(o8[0], o8[1], o8[2], o8[3]));
        st4(a.dx8 + (q.row0 + r) * 8 + 4, make_float4(o8[4], o8[5], o8[6], o8[7]));
      }
    }
  }
  // this CTA's partial sums
  float* P = a.partial + (long long)blockIdx.x * TL_PARTIAL;
#pragma unroll
  for (int s = 0; s < MAXSRC; ++s) {
    if (s >= a.nsrc) break;
#pragma unroll
    for (int i = 0; i < 4; ++i) st4(P + s * H * H + (o0 + i) * H + k0, make_float4(dw[s][i][0], dw[s][i][1], dw[s][i][2], dw[s][i][3]));
  }
  {
    const int o = threadIdx.x >> 2, kk = (threadIdx.x & 3) * 2;
    P[MAXSRC * H * H + o * 8 + kk] = dw8[0];
    P[MAXSRC * H * H + o * 8 + kk + 1] = dw8[1];
  }
  if (threadIdx.x < H) {
    P[MAXSRC * H * H + H * 8 + threadIdx.x] = dbias;
    P[MAXSRC * H * H + H * 8 + H + threadIdx.x] = dwd;
  }
}

// dest[o * ld + c0 + k] += sum over CTAs of partial[cta][off + o * ps + k], o < rows, k < K   (fixed order)
struct TlFinishItem { float* dest; int ld, c0, K, rows, off, ps; };
struct TlFinishArgs { const float* partial; int nparts; int stride; int nitems; TlFinishItem item[8]; };

__global__ void k_tl_finish(TlFinishArgs f) {
  for (int it = 0; it < f.nitems; ++it) {
    const TlFinishItem& m = f.item[it];
    const int total = m.rows * m.K;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
      const int o = idx / m.K, k = idx - o * m.K;
      const float* src = f.partial + m.off + o * m.ps + k;
      float s = 0.f;
      int c = 0;
      for (; c + 8 <= f.nparts; c += 8) {       // eight loads in flight, summed in the fixed order c = 0, 1, 2, ...
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = src[(long long)(c + u) * f.stride];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
      }
      for (; c < f.nparts; ++c) s += src[(long long)c * f.stride];
      m.dest[(long long)o * m.ld + m.c0 + k] += s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
// X0[r] = (s_delta xyz, attr, dens / 5000, 0, 0, 0)   (gnn_dyn.py:174-175)
__global__ void k_tl_node_in(const float* __restrict__ s_delta, const float* __restrict__ attr,
                             const float* __restrict__ dens, float* __restrict__ X0, int B, int N) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (long long)B * N) return;
  const float* sd = s_delta + r * 3;
  st4(X0 + r * 8, make_float4(sd[0], sd[1], sd[2], attr[r]));
  st4(X0 + r * 8 + 4, make_float4(dens[r / N] / 5000.f, 0.f, 0.f, 0.f));
}

// Y0[e] = (attr_r, attr_s, s_r - s_s, dens / 5000, 0, 0) for caller-provided relation lists (gnn_dyn.py:164-172, 179-180)
__global__ void k_tl_edge_in(const float* __restrict__ attr, const float* __restrict__ dens, const float* __restrict__ s_cur,
                             const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ row,
                             float* __restrict__ Y0, int B, int N) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rowptr[(long long)b * (N + 1) + N]) return;
  const long long slot = (long long)b * KMAX * N + e;
  const int r = row[slot], c = col[slot];
  const float* pr = s_cur + ((long long)b * N + r) * 3;
  const float* ps = s_cur + ((long long)b * N + c) * 3;
  st4(Y0 + slot * 8, make_float4(attr[(long long)b * N + r], attr[(long long)b * N + c], pr[0] - ps[0], pr[1] - ps[1]));
  st4(Y0 + slot * 8 + 4, make_float4(pr[2] - ps[2], dens[b] / 5000.f, 0.f, 0.f));
}

// agg[i] = sum_{e in row i} M[e]   (gnn_dyn.py:189), half-warp per particle
__global__ void k_tl_segsum(const int* __restrict__ rowptr, const float* __restrict__ M, float* __restrict__ agg, int B, int N) {
  const int l16 = threadIdx.x & 15;
  const long long node = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  if (node >= (long long)B * N) return;
  const int b = (int)(node / N), i = (int)(node - (long long)b * N);
  const int* rp = rowptr + (long long)b * (N + 1) + i;
  const long long slot = (long long)b * KMAX * N;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = rp[0]; e < rp[1]; ++e) {
    const float4 v = ld4(M + (slot + e) * H + 4 * l16);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  st4(agg + node * H + 4 * l16, s);
}

// s_pred = Q V1^T + b + s_cur   (gnn_dyn.py:196-198); V1T [64][4], b [4]
__global__ void k_tl_predict(const float* __restrict__ Q, const float* __restrict__ v1t, const float* __restrict__ b1,
                             const float* __restrict__ s_cur, float* __restrict__ s_out, int B, int N) {
  __shared__ float w[H * 4];
  for (int i = threadIdx.x; i < H * 4; i += blockDim.x) w[i] = v1t[i];
  __syncthreads();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (long long)B * N) return;
  float o0 = b1[0], o1 = b1[1], o2 = b1[2];
  for (int k = 0; k < H; k += 4) {
    const float4 q = ld4(Q + r * H + k);
    o0 = fmaf(q.x, w[(k + 0) * 4 + 0], o0); o1 = fmaf(q.x, w[(k + 0) * 4 + 1], o1); o2 = fmaf(q.x, w[(k + 0) * 4 + 2], o2);
    o0 = fmaf(q.y, w[(k + 1) * 4 + 0], o0); o1 = fmaf(q.y, w[(k + 1) * 4 + 1], o1); o2 = fmaf(q.y, w[(k + 1) * 4 + 2], o2);
    o0 = fmaf(q.z, w[(k + 2) * 4 + 0], o0); o1 = fmaf(q.z, w[(k + 2) * 4 + 1], o1); o2 = fmaf(q.z, w[(k + 2) * 4 + 2], o2);
    o0 = fmaf(q.w, w[(k + 3) * 4 + 0], o0); o1 = fmaf(q.w, w[(k + 3) * 4 + 1], o1); o2 = fmaf(q.w, w[(k + 3) * 4 + 2], o2);
  }
  s_out[r * 3 + 0] = o0 + s_cur[r * 3 + 0];
  s_out[r * 3 + 1] = o1 + s_cur[r * 3 + 1];
  s_out[r * 3 + 2] = o2 + s_cur[r * 3 + 2];
}

// gQ[r] = g[r] V1 (unmasked; the V0 layer's backward masks with Q > 0);  partial[cta] = (dV1 [3][64] | db [3] | 0)
constexpr int PB_THREADS = 256;
__global__ void __launch_bounds__(PB_THREADS)
k_tl_predict_bwd(const float* __restrict__ g, const float* __restrict__ Q, const float* __restrict__ v1,
                 float* __restrict__ gQ, float* __restrict__ partial, int B, int N) {
  __shared__ float w[4 * H];
  __shared__ float red[PB_THREADS / 64][4 * H];
  for (int i = threadIdx.x; i < 4 * H; i += blockDim.x) w[i] = v1[i];
  __syncthreads();
  const long long R = (long long)B * N;
  const int k = threadIdx.x & 63, sub = threadIdx.x >> 6;        // column k, row phase sub
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (long long r = (long long)blockIdx.x * (PB_THREADS / 64) + sub; r < R; r += (long long)gridDim.x * (PB_THREADS / 64)) {
    const float g0 = g[r * 3], g1 = g[r * 3 + 1], g2 = g[r * 3 + 2];
    gQ[r * H + k] = g0 * w[k] + g1 * w[H + k] + g2 * w[2 * H + k];
    const float q = Q[r * H + k];
    a0 = fmaf(g0, q, a0); a1 = fmaf(g1, q, a1); a2 = fmaf(g2, q, a2);
    s0 += g0; s1 += g1; s2 += g2;
  }
  red[sub][k] = a0; red[sub][H + k] = a1; red[sub][2 * H + k] = a2;
  red[sub][3 * H + k] = k == 0 ? s0 : (k == 1 ? s1 : (k == 2 ? s2 : 0.f));
  __syncthreads();
  if (sub == 0) {
    float* P = partial + (long long)blockIdx.x * TL_PARTIAL;
    for (int c = 0; c < 4; ++c) {
      float t = 0.f;
      for (int u = 0; u < PB_THREADS / 64; ++u) t += red[u][c * H + k];
      P[c * H + k] = t;
    }
  }
}

// out[i] (+)= base[i] + sum_{e in row i} A[e] + sum_{k in trow i} Bs[tedge k]      ([*, 64] rows; half-warp per particle)
__global__ void k_tl_gather_nodes(const int* __restrict__ rowptr, const int* __restrict__ trowptr,
                                  const int* __restrict__ tedge, const float* __restrict__ base,
                                  const float* __restrict__ A, const float* __restrict__ Bs, float* __restrict__ out,
                                  int accumulate, int B, int N) {
  const int l16 = threadIdx.x & 15;
  const long long node = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  if (node >= (long long)B * N) return;
  const int b = (int)(node / N), i = (int)(node - (long long)b * N);
  const long long slot = (long long)b * KMAX * N;
  float4 s = ld4(base + node * H + 4 * l16);
  const int* rp = rowptr + (long long)b * (N + 1) + i;
  for (int e = rp[0]; e < rp[1]; ++e) {
    const float4 v = ld4(A + (slot + e) * H + 4 * l16);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const int* tp = trowptr + (long long)b * (N + 1) + i;
  for (int k = tp[0]; k < tp[1]; ++k) {
    const float4 v = ld4(Bs + (slot + tedge[slot + k]) * H + 4 * l16);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float* o = out + node * H + 4 * l16;
  if (accumulate) { const float4 u = ld4(o); s.x += u.x; s.y += u.y; s.z += u.z; s.w += u.w; }
  st4(o, s);
}

// g_s_cur[i] = g_pred[i] + sum_{e in row i} dY0[e][2:5] - sum_{e: sender = i} dY0[e][2:5];  g_s_delta[i] = dX0[i][0:3]
__global__ void k_tl_positions(const int* __restrict__ rowptr, const int* __restrict__ trowptr,
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

// what the forward leaves for the backward
struct TrainTape {
  Csr csr;
  float *X0, *H0, *P, *eff[PSTEP], *agg[PSTEP], *Q;      // node rows
  float *Y0, *R1, *R2, *R3, *M[PSTEP];                   // relation slots
  size_t bytes;
};

TrainTape carve_train_tape(void* p, int B, int N) {
  Carve c(p);
  const size_t R = (size_t)B * N, E = (size_t)B * KMAX * N;
  TrainTape t;
  t.csr.rowptr = c.take<int>((size_t)B * (N + 1));
  t.csr.col = c.take<int>(E);
  t.csr.row = c.take<int>(E);
  t.csr.trowptr = c.take<int>((size_t)B * (N + 1));
  t.csr.trecv = c.take<int>(E);
  t.csr.tedge = c.take<int>(E);
  t.X0 = c.take<float>(R * 8);
  t.H0 = c.take<float>(R * H);
  t.P = c.take<float>(R * H);
  for (int p2 = 0; p2 < PSTEP; ++p2) t.eff[p2] = c.take<float>(R * H);
  for (int p2 = 0; p2 < PSTEP; ++p2) t.agg[p2] = c.take<float>(R * H);
  t.Q = c.take<float>(R * H);
  t.Y0 = c.take<float>((E + TILE) * 8);
  t.R1 = c.take<float>(E * H);
  t.R2 = c.take<float>(E * H);
  t.R3 = c.take<float>(E * H);
  for (int p2 = 0; p2 < PSTEP; ++p2) t.M[p2] = c.take<float>(E * H);
  t.bytes = c.off;
  return t;
}

struct TrainBwdScratch {
  float *gA, *gEff, *gZ, *gP, *gAgg, *gH0, *dX0;     // node rows
  float *gR3, *dZr, *dZs, *gR2, *dY0;                // relation slots (gR1 reuses dZr)
  float* partial;
  size_t bytes;
};

TrainBwdScratch carve_train_bwd(void* p, int B, int N) {
  Carve c(p);
  const size_t R = (size_t)B * N, E = (size_t)B * KMAX * N;
  TrainBwdScratch s;
  s.gA = c.take<float>(R * H);
  s.gEff = c.take<float>(R * H);
  s.gZ = c.take<float>(R * H);
  s.gP = c.take<float>(R * H);
  s.gAgg = c.take<float>(R * H);
  s.gH0 = c.take<float>(R * H);
  s.dX0 = c.take<float>(R * 8);
  s.gR3 = c.take<float>(E * H);
  s.dZr = c.take<float>(E * H);
  s.dZs = c.take<float>(E * H);
  s.gR2 = c.take<float>(E * H);
  s.dY0 = c.take<float>(E * 8);
  s.partial = c.take<float>((size_t)2 * NSM * TL_PARTIAL);
  s.bytes = c.off;
  return s;
}

template <bool EDGE>
int tl_grid(int B, int N, int per_sm) {
  const long long tiles = EDGE ? (long long)B * ((KMAX * N + TILE - 1) / TILE) : ((long long)B * N + TILE - 1) / TILE;
  const long long cap = (long long)per_sm * NSM;
  return (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
}

int tl_configure() {
  static DeviceOnce once;
  const int dev = once.pending();
  if (dev < 0) return 0;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k_tl_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TlFwdSmem)))) return (int)e;
  if ((e = cudaFuncSetAttribute(k_tl_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TlFwdSmem)))) return (int)e;
  if ((e = cudaFuncSetAttribute(k_tl_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TlBwdSmem)))) return (int)e;
  if ((e = cudaFuncSetAttribute(k_tl_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TlBwdSmem)))) return (int)e;
  once.done(dev);
  return 0;
}

template <bool EDGE>
int tl_forward(TlArgs a, cudaStream_t st) {
  k_tl_fwd<EDGE><<<tl_grid<EDGE>(a.B, a.N, 2), NT, sizeof(TlFwdSmem), st>>>(a);
  PILE_CHECK_LAUNCH();
  return 0;
}

// backward of one layer + the fixed-order reduction of its weight-gradient partial sums into the gradient buffer
template <bool EDGE>
int tl_backward(TlArgs a, const TlFinishItem* items, int nitems, cudaStream_t st) {
  const int grid = tl_grid<EDGE>(a.B, a.N, 1);
  k_tl_bwd<EDGE><<<grid, NT, sizeof(TlBwdSmem), st>>>(a);
  PILE_CHECK_LAUNCH();
  TlFinishArgs f{};
  f.partial = a.partial;
  f.nparts = grid;
  f.stride = TL_PARTIAL;
  f.nitems = nitems;
  for (int i = 0; i < nitems; ++i) f.item[i] = items[i];
  k_tl_finish<<<64, 256, 0, st>>>(f);
  PILE_CHECK_LAUNCH();
  return 0;
}

TlArgs tl_base(int B, int N, const Csr& csr) {
  TlArgs a{};
  a.B = B; a.N = N;
  a.rowptr = csr.rowptr; a.col = csr.col; a.row = csr.row;
  return a;
}

// offsets (floats) of the 18 gradient tensors inside the gradient buffer = reference state_dict order
struct GradOff {
  long long pe0_w, pe0_b, pe1_w, pe1_b, re0_w, re0_b, re1_w, re1_b, re2_w, re2_b, pp_w, pp_b, rp_w, rp_b, v0_w, v0_b, v1_w, v1_b,
      total;
};
GradOff grad_offsets() {
  GradOff g;
  long long o = 0;
  auto take = [&](long long n) { const long long r = o; o += n; return r; };
  g.pe0_w = take(H * 5); g.pe0_b = take(H);
  g.pe1_w = take(H * H); g.pe1_b = take(H);
  g.re0_w = take(H * 6); g.re0_b = take(H);
  g.re1_w = take(H * H); g.re1_b = take(H);
  g.re2_w = take(H * H); g.re2_b = take(H);
  g.pp_w = take(H * (2 * H + 1)); g.pp_b = take(H);
  g.rp_w = take(H * (3 * H + 1)); g.rp_b = take(H);
  g.v0_w = take(H * H); g.v0_b = take(H);
  g.v1_w = take(3 * H); g.v1_b = take(3);
  g.total = o;
  return g;
}

}  // namespace

long long train_tape_bytes(int B, int N) { return (long long)carve_train_tape(nullptr, B, N).bytes; }
long long train_bwd_scratch_bytes(int B, int N) { return (long long)carve_train_bwd(nullptr, B, N).bytes; }
long long train_grad_offset(int tensor_index) {
  const GradOff g = grad_offsets();
  const long long t[19] = {g.pe0_w, g.pe0_b, g.pe1_w, g.pe1_b, g.re0_w, g.re0_b, g.re1_w, g.re1_b, g.re2_w, g.re2_b,
                           g.pp_w, g.pp_b, g.rp_w, g.rp_b, g.v0_w, g.v0_b, g.v1_w, g.v1_b, g.total};
  return (tensor_index < 0 || tensor_index > 18) ? -1 : t[tensor_index];
}

int train_relations_view(void* tape, int B, int N, int** rowptr, int** col, int** row) {
  const TrainTape t = carve_train_tape(tape, B, N);
  *rowptr = t.csr.rowptr; *col = t.csr.col; *row = t.csr.row;
  return 0;
}

static int train_forward_body(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                              const float* s_delta, int B, int N, const TrainTape& t, float* s_pred, cudaStream_t st);

int launch_train_forward(const float* wpack, const float* attr, const float* dens, const int* particle_nums,
                         const float* s_cur, const float* s_delta, float adj_thresh, int B, int N, void* tape,
                         float* s_pred, cudaStream_t st) {
  int e = tl_configure();
  if (e) return e;
  const TrainTape t = carve_train_tape(tape, B, N);
  PushCam none{};
  e = launch_nbr_search(s_cur, (long long)N * 3, s_delta, nullptr, 0, none, nullptr, particle_nums, B, N,
                        adj_thresh * adj_thresh, t.csr, st, attr, dens, t.Y0);
  if (e) return e;
  return train_forward_body(wpack, attr, dens, s_cur, s_delta, B, N, t, s_pred, st);
}

// the same step on caller-provided relation lists (receiver-grouped CSR, any number of relations per receiver as long
// as a sample has at most KMAX * N in total): the "Rr / Rs" entry of PropModuleDiffDen.forward (gnn_dyn.py:147)
int launch_train_forward_relations(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                                   const float* s_delta, const int* rowptr, const int* col, const int* row, int B,
                                   int N, void* tape, float* s_pred, cudaStream_t st) {
  int e = tl_configure();
  if (e) return e;
  const TrainTape t = carve_train_tape(tape, B, N);
  const size_t E = (size_t)B * KMAX * N;
  cudaError_t ce;
  if ((ce = cudaMemcpyAsync(t.csr.rowptr, rowptr, sizeof(int) * (size_t)B * (N + 1), cudaMemcpyDeviceToDevice, st))) return (int)ce;
  if ((ce = cudaMemcpyAsync(t.csr.col, col, sizeof(int) * E, cudaMemcpyDeviceToDevice, st))) return (int)ce;
  if ((ce = cudaMemcpyAsync(t.csr.row, row, sizeof(int) * E, cudaMemcpyDeviceToDevice, st))) return (int)ce;
  if ((e = launch_transpose_relations(t.csr, B, N, st))) return e;
  const dim3 fgrid((KMAX * N + 255) / 256, B);
  k_tl_edge_in<<<fgrid, 256, 0, st>>>(attr, dens, s_cur, t.csr.rowptr, t.csr.col, t.csr.row, t.Y0, B, N);
  PILE_CHECK_LAUNCH();
  return train_forward_body(wpack, attr, dens, s_cur, s_delta, B, N, t, s_pred, st);
}

static int train_forward_body(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                              const float* s_delta, int B, int N, const TrainTape& t, float* s_pred, cudaStream_t st) {
  int e = 0;
  const long long R = (long long)B * N;
  k_tl_node_in<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(s_delta, attr, dens, t.X0, B, N);
  PILE_CHECK_LAUNCH();
  auto W = [&](int slot) { return wpack + wslot_offset(slot); };
  {  // particle encoder
    TlArgs a = tl_base(B, N, t.csr);
    a.x8 = t.X0; a.w8 = W(W_PE0T); a.bias = W(B_PE0); a.y = t.H0; a.relu = 1;
    if ((e = tl_forward<false>(a, st))) return e;
    a = tl_base(B, N, t.csr);
    a.nsrc = 1; a.src[0] = {t.H0, W(W_PE1T), 0, nullptr, 0}; a.bias = W(B_PE1); a.y = t.P; a.relu = 1;
    if ((e = tl_forward<false>(a, st))) return e;
  }
  {  // relation encoder
    TlArgs a = tl_base(B, N, t.csr);
    a.x8 = t.Y0; a.w8 = W(W_RE0T); a.bias = W(B_RE0); a.y = t.R1; a.relu = 1;
    if ((e = tl_forward<true>(a, st))) return e;
    a = tl_base(B, N, t.csr);
    a.nsrc = 1; a.src[0] = {t.R1, W(W_RE1T), 0, nullptr, 0}; a.bias = W(B_RE1); a.y = t.R2; a.relu = 1;
    if ((e = tl_forward<true>(a, st))) return e;
    a.src[0] = {t.R2, W(W_RE2T), 0, nullptr, 0}; a.bias = W(B_RE2); a.y = t.R3;
    if ((e = tl_forward<true>(a, st))) return e;
  }
  for (int p = 0; p < PSTEP; ++p) {
    const float* eff_in = p == 0 ? t.P : t.eff[p - 1];
    TlArgs a = tl_base(B, N, t.csr);
    a.nsrc = 3;
    a.src[0] = {t.R3, W(W_ET), 0, nullptr, 0};
    a.src[1] = {eff_in, W(W_RT), 1, nullptr, 0};
    a.src[2] = {eff_in, W(W_ST), 2, nullptr, 0};
    a.dens = dens; a.wd = W(WD_RP); a.bias = W(B_RP); a.y = t.M[p]; a.relu = 1;
    if ((e = tl_forward<true>(a, st))) return e;
    k_tl_segsum<<<(unsigned)((R * 16 + 255) / 256), 256, 0, st>>>(t.csr.rowptr, t.M[p], t.agg[p], B, N);
    PILE_CHECK_LAUNCH();
    a = tl_base(B, N, t.csr);
    a.nsrc = 2;
    a.src[0] = {t.P, W(W_PT), 0, nullptr, 0};
    a.src[1] = {t.agg[p], W(W_AT), 0, nullptr, 0};
    a.dens = dens; a.wd = W(WD_PP); a.bias = W(B_PP); a.res = eff_in; a.y = t.eff[p]; a.relu = 1;
    if ((e = tl_forward<false>(a, st))) return e;
  }
  {
    TlArgs a = tl_base(B, N, t.csr);
    a.nsrc = 1; a.src[0] = {t.eff[PSTEP - 1], W(W_V0T), 0, nullptr, 0}; a.bias = W(B_V0); a.y = t.Q; a.relu = 1;
    if ((e = tl_forward<false>(a, st))) return e;
  }
  k_tl_predict<<<(unsigned)((R + 127) / 128), 128, 0, st>>>(t.Q, W(W_V1T), W(B_V1), s_cur, s_pred, B, N);
  PILE_CHECK_LAUNCH();
  return 0;
}

int launch_train_backward(const float* wpack, const float* dens, void* tape, int B, int N, const float* g_pred,
                          float* g_s_cur, float* g_s_delta, float* grads, void* scratch, cudaStream_t st) {
  int e = tl_configure();
  if (e) return e;
  const TrainTape t = carve_train_tape(tape, B, N);
  const TrainBwdScratch s = carve_train_bwd(scratch, B, N);
  const GradOff go = grad_offsets();
  const long long R = (long long)B * N;
  auto W = [&](int slot) { return wpack + wslot_offset(slot); };
  constexpr int OFF8 = MAXSRC * H * H, OFFB = OFF8 + H * 8, OFFD = OFFB + H;

  // predictor output layer
  {
    const int grid = (int)((R + 3) / 4 < 2 * NSM ? (R + 3) / 4 : 2 * NSM);
    k_tl_predict_bwd<<<grid, PB_THREADS, 0, st>>>(g_pred, t.Q, W(W_V1), s.gA, s.partial, B, N);
    PILE_CHECK_LAUNCH();
    TlFinishArgs f{};
    f.partial = s.partial; f.nparts = grid; f.stride = TL_PARTIAL; f.nitems = 2;
    f.item[0] = {grads + go.v1_w, H, 0, H, 3, 0, H};
    f.item[1] = {grads + go.v1_b, 1, 0, 1, 3, 3 * H, 1};
    k_tl_finish<<<4, 256, 0, st>>>(f);
    PILE_CHECK_LAUNCH();
  }
  {  // V0
    TlArgs a = tl_base(B, N, t.csr);
    a.g = s.gA; a.ymask = t.Q; a.partial = s.partial;
    a.nsrc = 1; a.src[0] = {t.eff[PSTEP - 1], W(W_V0), 0, s.gEff, 0};
    const TlFinishItem it[2] = {{grads + go.v0_w, H, 0, H, H, 0, H}, {grads + go.v0_b, 1, 0, 1, H, OFFB, 1}};
    if ((e = tl_backward<false>(a, it, 2, st))) return e;
  }
  const int PPLD = 2 * H + 1, RPLD = 3 * H + 1;
  for (int p = PSTEP - 1; p >= 0; --p) {
    const float* eff_in = p == 0 ? t.P : t.eff[p - 1];
    {  // particle propagator: G = gEff masked by eff_{p+1}; also keep the masked rows (residual path)
      TlArgs a = tl_base(B, N, t.csr);
      a.g = s.gEff; a.ymask = t.eff[p]; a.g_out = s.gZ; a.partial = s.partial; a.dens = dens;
      a.nsrc = 2;
      a.src[0] = {t.P, W(W_P), 0, s.gP, p == PSTEP - 1 ? 0 : 1};
      a.src[1] = {t.agg[p], W(W_A), 0, s.gAgg, 0};
      const TlFinishItem it[4] = {{grads + go.pp_w, PPLD, 0, H, H, 0, H}, {grads + go.pp_w, PPLD, H, H, H, H * H, H},
                                  {grads + go.pp_w, PPLD, 2 * H, 1, H, OFFD, 1}, {grads + go.pp_b, 1, 0, 1, H, OFFB, 1}};
      if ((e = tl_backward<false>(a, it, 4, st))) return e;
    }
    {  // relation propagator: G(e) = gAgg[recv e] masked by M_p(e)
      TlArgs a = tl_base(B, N, t.csr);
      a.g = s.gAgg; a.g_gather = 1; a.ymask = t.M[p]; a.partial = s.partial; a.dens = dens;
      a.nsrc = 3;
      a.src[0] = {t.R3, W(W_E), 0, s.gR3, p == PSTEP - 1 ? 0 : 1};
      a.src[1] = {eff_in, W(W_R), 1, s.dZr, 0};
      a.src[2] = {eff_in, W(W_S), 2, s.dZs, 0};
      const TlFinishItem it[5] = {{grads + go.rp_w, RPLD, 0, H, H, 0, H}, {grads + go.rp_w, RPLD, H, H, H, H * H, H},
                                  {grads + go.rp_w, RPLD, 2 * H, H, H, 2 * H * H, H}, {grads + go.rp_w, RPLD, 3 * H, 1, H, OFFD, 1},
                                  {grads + go.rp_b, 1, 0, 1, H, OFFB, 1}};
      if ((e = tl_backward<true>(a, it, 5, st))) return e;
    }
    // d/d eff_p = masked rows (residual) + receiver-side + sender-side relation terms; p == 0: eff_0 is P
    k_tl_gather_nodes<<<(unsigned)((R * 16 + 255) / 256), 256, 0, st>>>(t.csr.rowptr, t.csr.trowptr, t.csr.tedge, s.gZ, s.dZr,
                                                                        s.dZs, p == 0 ? s.gP : s.gEff, p == 0 ? 1 : 0, B, N);
    PILE_CHECK_LAUNCH();
  }
  {  // particle encoder
    TlArgs a = tl_base(B, N, t.csr);
    a.g = s.gP; a.ymask = t.P; a.partial = s.partial;
    a.nsrc = 1; a.src[0] = {t.H0, W(W_PE1), 0, s.gH0, 0};
    const TlFinishItem it[2] = {{grads + go.pe1_w, H, 0, H, H, 0, H}, {grads + go.pe1_b, 1, 0, 1, H, OFFB, 1}};
    if ((e = tl_backward<false>(a, it, 2, st))) return e;
    a = tl_base(B, N, t.csr);
    a.g = s.gH0; a.ymask = t.H0; a.partial = s.partial;
    a.x8 = t.X0; a.w8 = W(W_PE0); a.dx8 = s.dX0;
    const TlFinishItem it0[2] = {{grads + go.pe0_w, 5, 0, 5, H, -1, 8}, {grads + go.pe0_b, 1, 0, 1, H, OFFB, 1}};
    TlFinishItem fix[2] = {it0[0], it0[1]};
    fix[0].off = OFF8;
    if ((e = tl_backward<false>(a, fix, 2, st))) return e;
  }
  {  // relation encoder
    TlArgs a = tl_base(B, N, t.csr);
    a.g = s.gR3; a.ymask = t.R3; a.partial = s.partial;
    a.nsrc = 1; a.src[0] = {t.R2, W(W_RE2), 0, s.gR2, 0};
    const TlFinishItem it2[2] = {{grads + go.re2_w, H, 0, H, H, 0, H}, {grads + go.re2_b, 1, 0, 1, H, OFFB, 1}};
    if ((e = tl_backward<true>(a, it2, 2, st))) return e;
    a = tl_base(B, N, t.csr);
    a.g = s.gR2; a.ymask = t.R2; a.partial = s.partial;
    a.nsrc = 1; a.src[0] = {t.R1, W(W_RE1), 0, s.dZr, 0};          // gR1 reuses the dZr buffer
    const TlFinishItem it1[2] = {{grads + go.re1_w, H, 0, H, H, 0, H}, {grads + go.re1_b, 1, 0, 1, H, OFFB, 1}};
    if ((e = tl_backward<true>(a, it1, 2, st))) return e;
    a = tl_base(B, N, t.csr);
    a.g = s.dZr; a.ymask = t.R1; a.partial = s.partial;
    a.x8 = t.Y0; a.w8 = W(W_RE0); a.dx8 = s.dY0;
    const TlFinishItem it0[2] = {{grads + go.re0_w, 6, 0, 6, H, OFF8, 8}, {grads + go.re0_b, 1, 0, 1, H, OFFB, 1}};
    if ((e = tl_backward<true>(a, it0, 2, st))) return e;
  }
  k_tl_positions<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(t.csr.rowptr, t.csr.trowptr, t.csr.tedge, s.dY0, s.dX0, g_pred,
                                                               g_s_cur, g_s_delta, B, N);
  PILE_CHECK_LAUNCH();
  return 0;
}

}  // namespace pile
