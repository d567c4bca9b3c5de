// K2-K5: forward of the propagation network on a CSR relation list (FP32 CUDA-core path).
//
// Reference: model/gnn_dyn.py:147-198.  The one-hot bmm gathers/scatters become index gathers and a
// receiver-segmented sum; the concatenations of the two Propagators are distributed over their weight
// blocks so the receiver/sender transforms run once per NODE and the relation-encoder contribution
// (W_e . rel_enc + w_d . d + b) is hoisted out of the 3 propagation steps:
//
//   C_e[e]   = W_e . RE(attr_r, attr_s, s_r - s_s, d) + w_d d + b          (k_edge_encode, once / step)
//   C_p[i]   = W_p . PE(s_delta_i, attr_i, d) + w_d' d + b'                 (k_node_encode)
//   repeat 3x (k_propagate):
//     agg[i] = sum_{e=(i<-j)} ReLU(C_e[e] + (W_r eff)[i] + (W_s eff)[j])
//     eff[i] = ReLU(C_p[i] + W_a agg[i] + eff[i])
//   s_pred   = s_cur + V1 ReLU(V0 eff + c0) + c1
//
// All dense contractions are [128 rows] x [K] x [64] tiles with the weights resident in shared memory.
#include "common.cuh"
#include "kernels.h"

namespace pile {

// ------------------------------------------------------------------------------------------------
// node encoder: p_enc, C_p, and the receiver/sender transforms for propagation step 0
// ------------------------------------------------------------------------------------------------
struct NodeEncSmem {
  float w_pe1[H * H], w_p[H * H], w_r[H * H], w_s[H * H];
  float w_pe0[8 * H];
  float b_pe0[H], b_pe1[H], wd_pp[H], b_pp[H];
  float X[TILE * LDX];
  float A[TILE * LDA];
  float dd[TILE];
};

__global__ void __launch_bounds__(NT, 2)
k_node_encode(const float* __restrict__ wpack, const float* __restrict__ attr, const float* __restrict__ dens,
              const float* __restrict__ s_delta, uint8_t* __restrict__ m_pe0, uint8_t* __restrict__ m_pe1,
              float* __restrict__ Cp, float* __restrict__ eff, float* __restrict__ Pr, float* __restrict__ Ps,
              int B, int N) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NodeEncSmem& S = *reinterpret_cast<NodeEncSmem*>(smem_raw);
  load_block(S.w_pe1, wpack + wslot_offset(W_PE1T), H * H);
  load_block(S.w_p, wpack + wslot_offset(W_PT), H * H);
  load_block(S.w_r, wpack + wslot_offset(W_RT), H * H);
  load_block(S.w_s, wpack + wslot_offset(W_ST), H * H);
  load_block(S.w_pe0, wpack + wslot_offset(W_PE0T), 8 * H);
  load_block(S.b_pe0, wpack + wslot_offset(B_PE0), H);
  load_block(S.b_pe1, wpack + wslot_offset(B_PE1), H);
  load_block(S.wd_pp, wpack + wslot_offset(WD_PP), H);
  load_block(S.b_pp, wpack + wslot_offset(B_PP), H);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c0 = warp * 8;
  const long long R = (long long)B * N;
  const int ntiles = (int)((R + TILE - 1) / TILE);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * TILE;
    const int nrows = (int)min((long long)TILE, R - row0);
    __syncthreads();   // previous tile's readers of X / A / dd are done
    if (threadIdx.x < TILE) {
      const int t = threadIdx.x;
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      float d = 0.f;
      if (t < nrows) {
        const long long r = row0 + t;
        d = dens[r / N] / 5000.f;
        const float* sd = s_delta + r * 3;
        x0 = make_float4(sd[0], sd[1], sd[2], attr[r]);
        x1.x = d;
      }
      st4(S.X + t * LDX, x0);
      st4(S.X + t * LDX + 4, x1);
      S.dd[t] = d;
    }
    __syncthreads();

    float acc[4][8];
    unsigned bits[4];
    acc_set_bias(acc, S.b_pe0, c0);
    gemm_rows4x8<8>(S.X, LDX, S.w_pe0, lane, c0, acc);
    acc_relu(acc, bits);
    mask_to_global(m_pe0, row0, nrows, lane, warp, bits);
    acc_to_smem(S.A, LDA, lane, c0, acc);
    __syncthreads();

    acc_set_bias(acc, S.b_pe1, c0);
    gemm_rows4x8<H>(S.A, LDA, S.w_pe1, lane, c0, acc);
    acc_relu(acc, bits);
    mask_to_global(m_pe1, row0, nrows, lane, warp, bits);
    acc_to_global(eff, row0, nrows, lane, c0, acc);          // effect_0 = particle_encode
    __syncthreads();
    acc_to_smem(S.A, LDA, lane, c0, acc);
    __syncthreads();

#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float d = S.dd[lane + 32 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(S.wd_pp[c0 + j], d, S.b_pp[c0 + j]);
    }
    gemm_rows4x8<H>(S.A, LDA, S.w_p, lane, c0, acc);
    acc_to_global(Cp, row0, nrows, lane, c0, acc);

#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    gemm_rows4x8<H>(S.A, LDA, S.w_r, lane, c0, acc);
    acc_to_global(Pr, row0, nrows, lane, c0, acc);

#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    gemm_rows4x8<H>(S.A, LDA, S.w_s, lane, c0, acc);
    acc_to_global(Ps, row0, nrows, lane, c0, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// relation encoder + hoisted relation-propagator term
// ------------------------------------------------------------------------------------------------
struct EdgeEncSmem {
  float w_re1[H * H], w_re2[H * H], w_e[H * H];
  float w_re0[8 * H];
  float b_re0[H], b_re1[H], b_re2[H], wd_rp[H], b_rp[H];
  float X[TILE * LDX];
  float A[TILE * LDA];
};

__global__ void __launch_bounds__(NT, 2)
k_edge_encode(const float* __restrict__ wpack, const float* __restrict__ attr, const float* __restrict__ dens,
              const float* __restrict__ s_cur, long long s_stride, const int* __restrict__ rowptr,
              const int* __restrict__ col, const int* __restrict__ row, uint8_t* __restrict__ m_re0,
              uint8_t* __restrict__ m_re1, uint8_t* __restrict__ m_re2, float* __restrict__ Ce, int B, int N) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EdgeEncSmem& S = *reinterpret_cast<EdgeEncSmem*>(smem_raw);
  load_block(S.w_re1, wpack + wslot_offset(W_RE1T), H * H);
  load_block(S.w_re2, wpack + wslot_offset(W_RE2T), H * H);
  load_block(S.w_e, wpack + wslot_offset(W_ET), H * H);
  load_block(S.w_re0, wpack + wslot_offset(W_RE0T), 8 * H);
  load_block(S.b_re0, wpack + wslot_offset(B_RE0), H);
  load_block(S.b_re1, wpack + wslot_offset(B_RE1), H);
  load_block(S.b_re2, wpack + wslot_offset(B_RE2), H);
  load_block(S.wd_rp, wpack + wslot_offset(WD_RP), H);
  load_block(S.b_rp, wpack + wslot_offset(B_RP), H);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c0 = warp * 8;
  const int tps = (KMAX * N + TILE - 1) / TILE;     // tiles per sample
  const long long ntiles = (long long)B * tps;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = (int)(tile / tps);
    const int e0 = (int)(tile % tps) * TILE;
    const int ne = rowptr[(long long)b * (N + 1) + N];
    if (e0 >= ne) continue;                           // CTA-uniform
    const int nrows = min(TILE, ne - e0);
    const long long slot0 = (long long)b * KMAX * N + e0;
    const float d = dens[b] / 5000.f;
    __syncthreads();
    if (threadIdx.x < TILE) {
      const int t = threadIdx.x;
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      if (t < nrows) {
        const int r = row[slot0 + t], c = col[slot0 + t];
        const float* pr = s_cur + (long long)b * s_stride + r * 3;
        const float* ps = s_cur + (long long)b * s_stride + c * 3;
        x0 = make_float4(attr[(long long)b * N + r], attr[(long long)b * N + c], pr[0] - ps[0], pr[1] - ps[1]);
        x1 = make_float4(pr[2] - ps[2], d, 0.f, 0.f);
      }
      st4(S.X + t * LDX, x0);
      st4(S.X + t * LDX + 4, x1);
    }
    __syncthreads();

    float acc[4][8];
    unsigned bits[4];
    acc_set_bias(acc, S.b_re0, c0);
    gemm_rows4x8<8>(S.X, LDX, S.w_re0, lane, c0, acc);
    acc_relu(acc, bits);
    mask_to_global(m_re0, slot0, nrows, lane, warp, bits);
    acc_to_smem(S.A, LDA, lane, c0, acc);
    __syncthreads();

    acc_set_bias(acc, S.b_re1, c0);
    gemm_rows4x8<H>(S.A, LDA, S.w_re1, lane, c0, acc);
    acc_relu(acc, bits);
    mask_to_global(m_re1, slot0, nrows, lane, warp, bits);
    __syncthreads();
    acc_to_smem(S.A, LDA, lane, c0, acc);
    __syncthreads();

    acc_set_bias(acc, S.b_re2, c0);
    gemm_rows4x8<H>(S.A, LDA, S.w_re2, lane, c0, acc);
    acc_relu(acc, bits);
    mask_to_global(m_re2, slot0, nrows, lane, warp, bits);
    __syncthreads();
    acc_to_smem(S.A, LDA, lane, c0, acc);
    __syncthreads();

#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(S.wd_rp[c0 + j], d, S.b_rp[c0 + j]);
    gemm_rows4x8<H>(S.A, LDA, S.w_e, lane, c0, acc);
    acc_to_global(Ce, slot0, nrows, lane, c0, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// one propagation step (LAST: also the predictor and the residual state update)
// ------------------------------------------------------------------------------------------------
struct PropSmem {
  float w_a[H * H], w_1[H * H], w_2[H * H];   // w_1/w_2: (W_r, W_s) or (V0, V1[H][4])
  float b_v0[H], b_v1[4];
  float A[TILE * LDA];
};

__device__ __forceinline__ float4 relu_add3(const float4 a, const float4 b, const float4 c, unsigned& m4) {
  float4 v = make_float4(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z, a.w + b.w + c.w);
  m4 = (v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) | (v.w > 0.f ? 8u : 0u);
  v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  return v;
}

template <bool LAST>
__global__ void __launch_bounds__(NT, 2)
k_propagate(const float* __restrict__ wpack, const int* __restrict__ rowptr, const int* __restrict__ col,
            const float* __restrict__ Ce, const float* __restrict__ Cp, float* __restrict__ eff,
            const float* __restrict__ Pr, const float* __restrict__ Ps, float* __restrict__ PrOut,
            float* __restrict__ PsOut, uint8_t* __restrict__ m_edge, uint8_t* __restrict__ m_eff,
            uint8_t* __restrict__ m_q, const float* __restrict__ s_cur, long long s_stride,
            float* __restrict__ s_out, long long o_stride, int B, int N) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PropSmem& S = *reinterpret_cast<PropSmem*>(smem_raw);
  load_block(S.w_a, wpack + wslot_offset(W_AT), H * H);
  if (LAST) {
    load_block(S.w_1, wpack + wslot_offset(W_V0T), H * H);
    load_block(S.w_2, wpack + wslot_offset(W_V1T), 4 * H);
    load_block(S.b_v0, wpack + wslot_offset(B_V0), H);
    if (threadIdx.x < 4) S.b_v1[threadIdx.x] = wpack[wslot_offset(B_V1) + threadIdx.x];
  } else {
    load_block(S.w_1, wpack + wslot_offset(W_RT), H * H);
    load_block(S.w_2, wpack + wslot_offset(W_ST), H * H);
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c0 = warp * 8;
  const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;      // 16 half-warps, 4 channels per lane
  const unsigned hmask = 0xffffu << (threadIdx.x & 16);         // the two half-warps of a warp diverge
  const long long R = (long long)B * N;
  const int ntiles = (int)((R + TILE - 1) / TILE);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * TILE;
    const int nrows = (int)min((long long)TILE, R - row0);
    __syncthreads();

    // phase 1: receiver-segmented sum of the relation effects  (gnn_dyn.py:183-189)
    for (int r = hw; r < TILE; r += NT / 16) {
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows) {
        const long long node = row0 + r;
        const int b = (int)(node / N), i = (int)(node % N);
        const int* rp = rowptr + (long long)b * (N + 1) + i;
        const int e_lo = rp[0], cnt = rp[1] - e_lo;
        const long long slot = (long long)b * KMAX * N + e_lo;
        const float4 pr = ld4(Pr + node * H + 4 * l16);
        for (int k0 = 0; k0 < cnt; k0 += 5) {
          float4 ce[5], ps[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            if (k0 + k < cnt) {
              const int s = col[slot + k0 + k];
              ce[k] = ld4(Ce + (slot + k0 + k) * H + 4 * l16);
              ps[k] = ld4(Ps + ((long long)b * N + s) * H + 4 * l16);
            }
          }
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            if (k0 + k < cnt) {
              unsigned m4;
              const float4 v = relu_add3(ce[k], pr, ps[k], m4);
              sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
              if (m_edge) {
                const unsigned hi = __shfl_down_sync(hmask, m4, 1);
                if ((l16 & 1) == 0) m_edge[(slot + k0 + k) * 8 + (l16 >> 1)] = (uint8_t)(m4 | (hi << 4));
              }
            }
          }
        }
      }
      st4(S.A + r * LDA + 4 * l16, sum);
    }
    __syncthreads();

    // phase 2: node update  eff <- ReLU(C_p + W_a agg + eff)   (gnn_dyn.py:191-193)
    float acc[4][8];
    unsigned bits[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = lane + 32 * i;
      if (r < nrows) {
        const float* cp = Cp + (row0 + r) * H + c0;
        const float* ef = eff + (row0 + r) * H + c0;
        const float4 a0 = ld4(cp), a1 = ld4(cp + 4), e0 = ld4(ef), e1 = ld4(ef + 4);
        acc[i][0] = a0.x + e0.x; acc[i][1] = a0.y + e0.y; acc[i][2] = a0.z + e0.z; acc[i][3] = a0.w + e0.w;
        acc[i][4] = a1.x + e1.x; acc[i][5] = a1.y + e1.y; acc[i][6] = a1.z + e1.z; acc[i][7] = a1.w + e1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      }
    }
    gemm_rows4x8<H>(S.A, LDA, S.w_a, lane, c0, acc);
    acc_relu(acc, bits);
    mask_to_global(m_eff, row0, nrows, lane, warp, bits);
    if (!LAST) acc_to_global(eff, row0, nrows, lane, c0, acc);
    __syncthreads();
    acc_to_smem(S.A, LDA, lane, c0, acc);
    __syncthreads();

    if (!LAST) {
      // receiver / sender transforms for the next propagation step
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      gemm_rows4x8<H>(S.A, LDA, S.w_1, lane, c0, acc);
      acc_to_global(PrOut, row0, nrows, lane, c0, acc);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      gemm_rows4x8<H>(S.A, LDA, S.w_2, lane, c0, acc);
      acc_to_global(PsOut, row0, nrows, lane, c0, acc);
    } else {
      // predictor + residual  (gnn_dyn.py:196-198)
      acc_set_bias(acc, S.b_v0, c0);
      gemm_rows4x8<H>(S.A, LDA, S.w_1, lane, c0, acc);
      acc_relu(acc, bits);
      mask_to_global(m_q, row0, nrows, lane, warp, bits);
      __syncthreads();
      acc_to_smem(S.A, LDA, lane, c0, acc);
      __syncthreads();
      if (threadIdx.x < nrows) {
        const int r = threadIdx.x;
        float o0 = S.b_v1[0], o1 = S.b_v1[1], o2 = S.b_v1[2];
#pragma unroll 4
        for (int k = 0; k < H; k += 4) {
          const float4 a = ld4(S.A + r * LDA + k);
          const float4 w0 = ld4(S.w_2 + (k + 0) * 4), w1 = ld4(S.w_2 + (k + 1) * 4);
          const float4 w2 = ld4(S.w_2 + (k + 2) * 4), w3 = ld4(S.w_2 + (k + 3) * 4);
          o0 = fmaf(a.x, w0.x, o0); o1 = fmaf(a.x, w0.y, o1); o2 = fmaf(a.x, w0.z, o2);
          o0 = fmaf(a.y, w1.x, o0); o1 = fmaf(a.y, w1.y, o1); o2 = fmaf(a.y, w1.z, o2);
          o0 = fmaf(a.z, w2.x, o0); o1 = fmaf(a.z, w2.y, o1); o2 = fmaf(a.z, w2.z, o2);
          o0 = fmaf(a.w, w3.x, o0); o1 = fmaf(a.w, w3.y, o1); o2 = fmaf(a.w, w3.z, o2);
        }
        const long long node = row0 + r;
        const int b = (int)(node / N), i = (int)(node % N);
        const float* sc = s_cur + (long long)b * s_stride + i * 3;
        float* so = s_out + (long long)b * o_stride + i * 3;
        so[0] = o0 + sc[0];
        so[1] = o1 + sc[1];
        so[2] = o2 + sc[2];
      }
    }
  }
}

std::atomic<int> g_use_tensor_cores{2};   // default: tcgen05 tiles, relation encoder with A in tensor memory

template <typename Kern>
static int set_smem(Kern k, size_t bytes) {
  return (int)cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int launch_forward(const float* wpack, const float* attr, const float* dens, const float* s_cur,
                   long long s_cur_stride, const float* s_delta, const Csr& csr, const StepScratch& ws,
                   const Masks* mk, float* s_out, long long s_out_stride, int B, int N, cudaStream_t st,
                   cudaEvent_t* ev, bool efeat_ready) {
  static DeviceOnce once;
  const int once_dev = once.pending();
  if (once_dev >= 0) {
    int e;
    if ((e = set_smem(k_node_encode, sizeof(NodeEncSmem)))) return e;
    if ((e = set_smem(k_edge_encode, sizeof(EdgeEncSmem)))) return e;
    if ((e = set_smem(k_propagate<false>, sizeof(PropSmem)))) return e;
    if ((e = set_smem(k_propagate<true>, sizeof(PropSmem)))) return e;
    once.done(once_dev);
  }
  const long long R = (long long)B * N;
  const int node_tiles = (int)((R + TILE - 1) / TILE);
  const long long edge_tiles = (long long)B * ((KMAX * N + TILE - 1) / TILE);
  const int persistent = 2 * NSM;
  const int g_node = node_tiles < persistent ? node_tiles : persistent;
  const int g_edge = edge_tiles < persistent ? (int)edge_tiles : persistent;

  if (g_use_tensor_cores) {
    // tcgen05 path: every dense contraction on the tensor cores, the gather/segmented sum as its own
    // streaming kernel
    int e;
    if (ev) cudaEventRecord(ev[0], st);
    if ((e = launch_node_encode_tc(wpack, attr, dens, s_delta, mk, ws, B, N, st))) return e;
    if (ev) cudaEventRecord(ev[1], st);
    if ((e = launch_edge_encode_tc(wpack, attr, dens, s_cur, s_cur_stride, csr, mk, ws.efeat, ws.Ce, B, N, st, efeat_ready))) return e;
    for (int p = 0; p < PSTEP; ++p) {
      if (ev) cudaEventRecord(ev[2 + 2 * p], st);
      if ((e = launch_propagate_tc(wpack, csr, ws, mk, p, s_cur, s_cur_stride, s_out, s_out_stride, B, N, st,
                                   ev ? ev[3 + 2 * p] : nullptr)))
        return e;
    }
    if (ev) cudaEventRecord(ev[2 + 2 * PSTEP], st);
    return 0;
  }
  if (ev) cudaEventRecord(ev[0], st);
  k_node_encode<<<g_node, NT, sizeof(NodeEncSmem), st>>>(wpack, attr, dens, s_delta, mk ? mk->pe0 : nullptr,
                                                         mk ? mk->pe1 : nullptr, ws.Cp, ws.eff, ws.Pr[0],
                                                         ws.Ps[0], B, N);
  PILE_CHECK_LAUNCH();
  if (ev) cudaEventRecord(ev[1], st);
  k_edge_encode<<<g_edge, NT, sizeof(EdgeEncSmem), st>>>(wpack, attr, dens, s_cur, s_cur_stride, csr.rowptr,
                                                         csr.col, csr.row, mk ? mk->re0 : nullptr,
                                                         mk ? mk->re1 : nullptr, mk ? mk->re2 : nullptr, ws.Ce, B, N);
  PILE_CHECK_LAUNCH();
  for (int p = 0; p < PSTEP; ++p) {
    const int in = p & 1, out = in ^ 1;
    if (ev) { cudaEventRecord(ev[2 + 2 * p], st); cudaEventRecord(ev[3 + 2 * p], st); }   // no separate segmented sum
    if (p < PSTEP - 1) {
      k_propagate<false><<<g_node, NT, sizeof(PropSmem), st>>>(
          wpack, csr.rowptr, csr.col, ws.Ce, ws.Cp, ws.eff, ws.Pr[in], ws.Ps[in], ws.Pr[out], ws.Ps[out],
          mk ? mk->edge[p] : nullptr, mk ? mk->eff[p] : nullptr, nullptr, s_cur, s_cur_stride, s_out,
          s_out_stride, B, N);
    } else {
      k_propagate<true><<<g_node, NT, sizeof(PropSmem), st>>>(
          wpack, csr.rowptr, csr.col, ws.Ce, ws.Cp, ws.eff, ws.Pr[in], ws.Ps[in], nullptr, nullptr,
          mk ? mk->edge[p] : nullptr, mk ? mk->eff[p] : nullptr, mk ? mk->q : nullptr, s_cur, s_cur_stride,
          s_out, s_out_stride, B, N);
    }
    PILE_CHECK_LAUNCH();
  }
  if (ev) cudaEventRecord(ev[2 + 2 * PSTEP], st);
  return 0;
}

}  // namespace pile
