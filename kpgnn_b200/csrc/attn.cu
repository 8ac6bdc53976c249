// attn.cu -- AttentionCombine (layers/combine.py:8-27) as ONE kernel per direction (sm_100a).
//
//   score[n,t,:] = biLSTM(x[n,:,:])            input size d, hidden size K, sequence length K (the hop axis)
//   w[n,:]       = softmax_t( sum_j score[n,t,j] )
//   out[n,:]     = sum_t w[n,t] * x[n,t,:]
//
// The reference runs a cuDNN LSTM over [N,K,d], then sum / softmax / multiply / sum as separate kernels, each
// re-reading [N,K,d] or [N,K,2K] from HBM.  Here one warp owns one node: its [K,d] slice is read ONCE into shared
// memory, the input projection (the LSTM's only GEMM-shaped part, 8K x d per time step) runs as register-tiled FMAs
// against weights staged once per CTA, the 2 x K recurrence steps run on the warp (lane = gate row, lanes < 2K = cell
// states), and the softmax-weighted sum reuses the staged slice.  Nothing of size [N,K,*] is written.
//
// Backward (same reference lines through autograd): the warp recomputes the forward (cheaper than saving 8K^2 gate
// values per node), back-propagates through the softmax and both recurrences, and produces dX [N,K,d] directly; the
// pre-activation gradients dG [N,K,8K] go to a workspace from which dW_ih = dG^T X is formed by a fixed-order
// split reduction (attn_dwih_kernel); dW_hh and the biases are accumulated per warp and reduced in a fixed order.
// No float atomics: bit-reproducible.
//
// Row numbering: R = 8K "global" gate rows, row Rg = dir * 4K + r with r the row of PyTorch's weight_ih_l0 /
// weight_hh_l0 (gate order i, f, g, o); lane l owns rows l, l + 32, ...
#include <math.h>

#include "common.cuh"

namespace kp {

constexpr int AT_MAX_WARPS = 8;      // warps per CTA: as many (<= 8) as the per-warp staging lets fit in shared memory

struct AttnLayout {
  int K, d, R, dq, ds, Ks;          // dq = ceil(d/4); ds = padded row stride of x / W_ih rows (floats); Ks = W_hh row stride
  int o_wih, o_whh, o_bias;         // CTA-wide (floats)
  int o_warp, warp_floats;          // per-warp region
  int w_xs, w_gin, w_H, w_C, w_hbuf, w_S, w_do, w_dG, w_dwhh;
  int total_floats;
};

__host__ __device__ inline AttnLayout attn_layout(int K, int d, bool backward, int nwarps) {
  AttnLayout L;
  L.K = K;
  L.d = d;
  L.R = 8 * K;
  L.dq = (d + 3) / 4;
  L.ds = 4 * (L.dq | 1);            // odd number of 16-byte chunks per row: LDS.128 of 8 consecutive rows is conflict-free
  L.Ks = K | 1;
  int o = 0;
  L.o_wih = o;  o += L.R * L.ds;
  L.o_whh = o;  o += L.R * L.Ks;
  L.o_bias = o; o += L.R;
  o = (o + 3) & ~3;
  L.o_warp = o;
  int w = 0;
  L.w_xs = w;   w += K * L.ds;
  L.w_gin = w;  w += K * L.R;
  L.w_H = w;    w += 2 * K * K;
  L.w_C = w;    w += backward ? 2 * K * K : 0;
  L.w_hbuf = w; w += 2 * K;
  L.w_S = w;    w += 2 * K;         // [0,K) softmax weights, [K,2K) dscore (backward)
  w = (w + 3) & ~3;
  L.w_do = w;   w += backward ? L.ds : 0;
  L.w_dG = w;   w += backward ? K * L.R : 0;
  L.w_dwhh = w; w += backward ? L.R * L.Ks : 0;
  L.warp_floats = (w + 3) & ~3;
  L.total_floats = L.o_warp + nwarps * L.warp_floats;
  return L;
}

__device__ __forceinline__ float at_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float at_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float at_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// weights of both directions -> shared memory (once per CTA)
__device__ __forceinline__ void attn_stage_weights(const kp_attn_desc& a, const AttnLayout& L, float* sm) {
  const int K = L.K, d = L.d, R = L.R, R4 = 4 * K;
  for (int i = threadIdx.x; i < R * L.ds; i += blockDim.x) {
    const int Rg = i / L.ds, c = i - Rg * L.ds;
    const int dir = Rg / R4, r = Rg - dir * R4;
    sm[L.o_wih + i] = c < d ? __ldg(a.w_ih[dir] + (size_t)r * d + c) : 0.f;
  }
  for (int i = threadIdx.x; i < R * L.Ks; i += blockDim.x) {
    const int Rg = i / L.Ks, j = i - Rg * L.Ks;
    const int dir = Rg / R4, r = Rg - dir * R4;
    sm[L.o_whh + i] = j < K ? __ldg(a.w_hh[dir] + (size_t)r * K + j) : 0.f;
  }
  for (int Rg = threadIdx.x; Rg < R; Rg += blockDim.x) {
    const int dir = Rg / R4, r = Rg - dir * R4;
    sm[L.o_bias + Rg] = __ldg(a.b_ih[dir] + r) + __ldg(a.b_hh[dir] + r);
  }
}

// x[v] ([K,d], node / hop strides) -> the warp's staging rows (zero padded to ds)
__device__ __forceinline__ void attn_load_x(const kp_attn_desc& a, const AttnLayout& L, int v, float* xs, int lane,
                                            bool vec) {
  const int K = L.K, d = L.d;
  const float* xv = a.x + (size_t)v * a.x_node_stride;
  if (vec) {
    for (int t = 0; t < K; ++t) {
      const float* row = xv + (size_t)t * a.x_hop_stride;
      for (int q = lane; q < L.ds / 4; q += 32) {
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < L.dq) val = __ldg(reinterpret_cast<const float4*>(row) + q);
        *reinterpret_cast<float4*>(xs + t * L.ds + 4 * q) = val;
      }
    }
  } else {
    for (int t = 0; t < K; ++t) {
      const float* row = xv + (size_t)t * a.x_hop_stride;
      for (int c = lane; c < L.ds; c += 32) xs[t * L.ds + c] = c < d ? __ldg(row + c) : 0.f;
    }
  }
}

// Forward of one node on one warp: fills gin (ACTIVATED gates, time-major [K][R]), H [2][K][K] (hidden outputs),
// C [2][K][K] when SAVE_C, S[0..K) = softmax weights.  xs must already hold the node's slice.
template <int KK, bool SAVE_C>
__device__ __forceinline__ void attn_node_forward(const AttnLayout& L, const float* __restrict__ sm, float* __restrict__ wsm,
                                                  int lane) {
  constexpr int NR = (8 * KK + 31) / 32;
  const int K = L.K, R = L.R, R4 = 4 * K, ds = L.ds, Ks = L.Ks;
  const float* Wih = sm + L.o_wih;
  const float* Whh = sm + L.o_whh;
  const float* bias = sm + L.o_bias;
  const float* xs = wsm + L.w_xs;
  float* gin = wsm + L.w_gin;
  float* H = wsm + L.w_H;
  float* C = wsm + L.w_C;
  float* hbuf = wsm + L.w_hbuf;
  float* S = wsm + L.w_S;
  // ---- input projection: gin[t][Rg] = bias[Rg] + W_ih[Rg,:] . x[t,:]
  {
    float acc[NR][KK];
#pragma unroll
    for (int s = 0; s < NR; ++s)
#pragma unroll
      for (int t = 0; t < KK; ++t) acc[s][t] = 0.f;
    int rows[NR];
#pragma unroll
    for (int s = 0; s < NR; ++s) rows[s] = min(lane + 32 * s, R - 1);
    for (int c4 = 0; c4 < 4 * L.dq; c4 += 4) {
      float4 w[NR];
#pragma unroll
      for (int s = 0; s < NR; ++s) w[s] = *reinterpret_cast<const float4*>(Wih + rows[s] * ds + c4);
#pragma unroll
      for (int t = 0; t < KK; ++t) {
        if (t < K) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + t * ds + c4);
#pragma unroll
          for (int s = 0; s < NR; ++s)
            acc[s][t] = fmaf(w[s].x, xv.x, fmaf(w[s].y, xv.y, fmaf(w[s].z, xv.z, fmaf(w[s].w, xv.w, acc[s][t]))));
        }
      }
    }
#pragma unroll
    for (int s = 0; s < NR; ++s) {
      const int Rg = lane + 32 * s;
      if (Rg < R) {
        const float b = bias[Rg];
#pragma unroll
        for (int t = 0; t < KK; ++t)
          if (t < K) gin[t * R + Rg] = acc[s][t] + b;
      }
    }
  }
  if (lane < 2 * K) hbuf[lane] = 0.f;
  __syncwarp();
  // ---- recurrence, both directions at once: step q handles time q (forward) and K-1-q (reverse)
  float cstate = 0.f;
  const int cdir = lane / K, cj = lane - cdir * K;        // meaningful on lanes < 2K
  for (int q = 0; q < K; ++q) {
#pragma unroll
    for (int s = 0; s < NR; ++s) {
      const int Rg = lane + 32 * s;
      if (Rg < R) {
        const int dir = Rg / R4, r = Rg - dir * R4;
        const int t = dir ? K - 1 - q : q;
        float pre = gin[t * R + Rg];
        const float* wr = Whh + Rg * Ks;
        const float* hb = hbuf + dir * K;
        for (int j = 0; j < K; ++j) pre = fmaf(wr[j], hb[j], pre);
        gin[t * R + Rg] = (r / K == 2) ? tanhf(pre) : at_sigmoid(pre);
      }
    }
    __syncwarp();
    if (lane < 2 * K) {
      const int t = cdir ? K - 1 - q : q;
      const float* g = gin + t * R + cdir * R4;
      const float gi = g[cj], gf = g[K + cj], gg = g[2 * K + cj], go = g[3 * K + cj];
      cstate = fmaf(gf, cstate, gi * gg);
      const float h = go * tanhf(cstate);
      hbuf[lane] = h;
      H[(cdir * K + t) * K + cj] = h;
      if (SAVE_C) C[(cdir * K + t) * K + cj] = cstate;
    }
    __syncwarp();
  }
  // ---- scores and softmax over the hop axis
  float s = -INFINITY;
  if (lane < K) {
    s = 0.f;
    for (int j = 0; j < K; ++j) s += H[lane * K + j];
    for (int j = 0; j < K; ++j) s += H[(K + lane) * K + j];
  }
  const float m = at_warp_max(s);
  const float e = lane < K ? expf(s - m) : 0.f;
  const float den = at_warp_sum(e);
  if (lane < K) S[lane] = e / den;
  __syncwarp();
}

template <int KK>
__global__ void __launch_bounds__(32 * AT_MAX_WARPS)
attn_fwd_kernel(const kp_attn_desc a, float* __restrict__ out, float* __restrict__ wts, int vec_in, int vec_out) {
  extern __shared__ __align__(16) float sm[];
  const int AT_WARPS = blockDim.x >> 5;
  const AttnLayout L = attn_layout(a.K, a.d, false, AT_WARPS);
  attn_stage_weights(a, L, sm);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wsm = sm + L.o_warp + warp * L.warp_floats;
  float* xs = wsm + L.w_xs;
  const float* S = wsm + L.w_S;
  const int K = L.K, d = L.d;
  for (int v = blockIdx.x * AT_WARPS + warp; v < a.N; v += gridDim.x * AT_WARPS) {
    __syncwarp();
    attn_load_x(a, L, v, xs, lane, vec_in != 0);
    __syncwarp();
    attn_node_forward<KK, false>(L, sm, wsm, lane);
    if (wts && lane < K) wts[(size_t)v * K + lane] = S[lane];
    float* ov = out + (size_t)v * d;
    for (int q = lane; q < L.dq; q += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int t = 0; t < K; ++t) {
        const float w = S[t];
        const float4 xv = *reinterpret_cast<const float4*>(xs + t * L.ds + 4 * q);
        acc.x = fmaf(w, xv.x, acc.x); acc.y = fmaf(w, xv.y, acc.y);
        acc.z = fmaf(w, xv.z, acc.z); acc.w = fmaf(w, xv.w, acc.w);
      }
      if (vec_out) {
        *reinterpret_cast<float4*>(ov + 4 * q) = acc;
      } else {
        const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
        for (int i = 0; i < 4; ++i)
          if (4 * q + i < d) ov[4 * q + i] = vals[i];
      }
    }
  }
}

// per-CTA partial layout of the small parameter gradients: [R*K] dW_hh rows (row Rg, column j) then [R] bias
template <int KK>
__global__ void __launch_bounds__(32 * AT_MAX_WARPS)
attn_bwd_kernel(const kp_attn_desc a, const float* __restrict__ dOut, float* __restrict__ dX, float* __restrict__ dG,
                float* __restrict__ small_part, int vec_in, int vec_out) {
  constexpr int NR = (8 * KK + 31) / 32;
  extern __shared__ __align__(16) float sm[];
  const int AT_WARPS = blockDim.x >> 5;
  const AttnLayout L = attn_layout(a.K, a.d, true, AT_WARPS);
  attn_stage_weights(a, L, sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wsm = sm + L.o_warp + warp * L.warp_floats;
  const int K = L.K, d = L.d, R = L.R, R4 = 4 * K, ds = L.ds, Ks = L.Ks;
  float* xs = wsm + L.w_xs;
  float* gin = wsm + L.w_gin;
  float* H = wsm + L.w_H;
  float* C = wsm + L.w_C;
  float* S = wsm + L.w_S;
  float* DS = S + K;
  float* dos = wsm + L.w_do;
  float* dGs = wsm + L.w_dG;
  float* dwhh = wsm + L.w_dwhh;
  const float* Wih = sm + L.o_wih;
  const float* Whh = sm + L.o_whh;
  for (int i = lane; i < R * Ks; i += 32) dwhh[i] = 0.f;
  float accb[NR];
#pragma unroll
  for (int s = 0; s < NR; ++s) accb[s] = 0.f;
  __syncthreads();
  const int cdir = lane / K, cj = lane - cdir * K;
  for (int v = blockIdx.x * AT_WARPS + warp; v < a.N; v += gridDim.x * AT_WARPS) {
    __syncwarp();
    attn_load_x(a, L, v, xs, lane, vec_in != 0);
    {
      const float* dv = dOut + (size_t)v * d;
      for (int c = lane; c < ds; c += 32) dos[c] = c < d ? __ldg(dv + c) : 0.f;
    }
    __syncwarp();
    attn_node_forward<KK, true>(L, sm, wsm, lane);
    // ---- softmax backward: dw[t] = dOut . x[t];  dscore[t] = w[t] (dw[t] - sum_u w[u] dw[u])
    {
      float mydw = 0.f, dot = 0.f;
      for (int t = 0; t < K; ++t) {
        float p = 0.f;
        for (int c = lane; c < 4 * L.dq; c += 32) p = fmaf(dos[c], xs[t * ds + c], p);
        p = at_warp_sum(p);
        dot = fmaf(S[t], p, dot);
        if (lane == t) mydw = p;
      }
      if (lane < K) DS[lane] = S[lane] * (mydw - dot);
    }
    __syncwarp();
    // ---- back-propagation through time, both directions at once (processing step q: forward time q, reverse K-1-q)
    float dh_rec = 0.f, dc_next = 0.f;
    for (int q = K - 1; q >= 0; --q) {
      if (lane < 2 * K) {
        const int t = cdir ? K - 1 - q : q;
        const int tprev = cdir ? K - q : q - 1;
        const float* g = gin + t * R + cdir * R4;
        const float gi = g[cj], gf = g[K + cj], gg = g[2 * K + cj], go = g[3 * K + cj];
        const float ct = C[(cdir * K + t) * K + cj];
        const float cprev = q > 0 ? C[(cdir * K + tprev) * K + cj] : 0.f;
        const float tc = tanhf(ct);
        const float dh = DS[t] + dh_rec;
        const float dc = fmaf(dh * go, 1.f - tc * tc, dc_next);
        dc_next = dc * gf;
        float* o = dGs + t * R + cdir * R4;
        o[cj] = dc * gg * gi * (1.f - gi);
        o[K + cj] = dc * cprev * gf * (1.f - gf);
        o[2 * K + cj] = dc * gi * (1.f - gg * gg);
        o[3 * K + cj] = dh * tc * go * (1.f - go);
      }
      __syncwarp();
      if (lane < 2 * K) {                       // dh of the previous processing step: W_hh^T dG
        const int t = cdir ? K - 1 - q : q;
        const float* gr = dGs + t * R + cdir * R4;
        const float* wc = Whh + (cdir * R4) * Ks + cj;
        float s = 0.f;
        for (int r = 0; r < R4; ++r) s = fmaf(wc[r * Ks], gr[r], s);
        dh_rec = s;
      }
      if (q > 0) {                              // dW_hh[Rg][j] += dG[t][Rg] * h_prev[j]; rows owned by lanes
#pragma unroll
        for (int s = 0; s < NR; ++s) {
          const int Rg = lane + 32 * s;
          if (Rg < R) {
            const int dir = Rg / R4;
            const int t = dir ? K - 1 - q : q;
            const int tprev = dir ? K - q : q - 1;
            const float gval = dGs[t * R + Rg];
            const float* hp = H + (dir * K + tprev) * K;
            float* wrow = dwhh + Rg * Ks;
            for (int j = 0; j < K; ++j) wrow[j] = fmaf(gval, hp[j], wrow[j]);
          }
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int s = 0; s < NR; ++s) {
      const int Rg = lane + 32 * s;
      if (Rg < R)
        for (int t = 0; t < K; ++t) accb[s] += dGs[t * R + Rg];
    }
    // ---- dG to the workspace (for dW_ih), coalesced
    {
      float* gv = dG + (size_t)v * K * R;
      for (int i = lane; i < K * R; i += 32) gv[i] = dGs[i];
    }
    // ---- dX[t][c] = w[t] dOut[c] + sum_Rg dG[t][Rg] W_ih[Rg][c]
    float* dxv = dX + (size_t)v * K * d;
    for (int q4 = lane; q4 < L.dq; q4 += 32) {
      float4 acc[KK];
      const float4 dov = *reinterpret_cast<const float4*>(dos + 4 * q4);
#pragma unroll
      for (int t = 0; t < KK; ++t) {
        const float w = t < K ? S[t] : 0.f;
        acc[t] = make_float4(w * dov.x, w * dov.y, w * dov.z, w * dov.w);
      }
      for (int Rg = 0; Rg < R; ++Rg) {
        const float4 wv = *reinterpret_cast<const float4*>(Wih + Rg * ds + 4 * q4);
#pragma unroll
        for (int t = 0; t < KK; ++t) {
          if (t < K) {
            const float g = dGs[t * R + Rg];
            acc[t].x = fmaf(g, wv.x, acc[t].x); acc[t].y = fmaf(g, wv.y, acc[t].y);
            acc[t].z = fmaf(g, wv.z, acc[t].z); acc[t].w = fmaf(g, wv.w, acc[t].w);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < KK; ++t) {
        if (t < K) {
          float* o = dxv + (size_t)t * d + 4 * q4;
          if (vec_out) {
            *reinterpret_cast<float4*>(o) = acc[t];
          } else {
            const float vals[4] = {acc[t].x, acc[t].y, acc[t].z, acc[t].w};
            for (int i = 0; i < 4; ++i)
              if (4 * q4 + i < d) o[i] = vals[i];
          }
        }
      }
    }
  }
  // ---- fixed-order reduction of the per-warp dW_hh / bias accumulators over the CTA's warps
  __syncwarp();
  float* bsum = wsm + L.w_dG;                 // the warp's dG staging is free now: park the bias sums there
#pragma unroll
  for (int s = 0; s < NR; ++s) {
    const int Rg = lane + 32 * s;
    if (Rg < R) bsum[Rg] = accb[s];
  }
  __syncthreads();
  float* part = small_part + (size_t)blockIdx.x * (R * K + R);
  for (int i = threadIdx.x; i < R * K + R; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < AT_WARPS; ++w) {
      const float* ws = sm + L.o_warp + w * L.warp_floats;
      s += i < R * K ? ws[L.w_dwhh + (i / K) * Ks + (i % K)] : ws[L.w_dG + (i - R * K)];
    }
    part[i] = s;
  }
}

// dW_ih[Rg][c] = sum over m = (n,t) of dG[m][Rg] * x[m][c].  One CTA per chunk of m; 4 x 4 register tiles.
constexpr int DW_TILE = 32;
constexpr int DW_TPT = 4;
__global__ void __launch_bounds__(256)
attn_dwih_kernel(const kp_attn_desc a, const float* __restrict__ dG, float* __restrict__ part, int chunk, int vec_in) {
  extern __shared__ __align__(16) float sm[];
  const int K = a.K, d = a.d, R = 8 * K;
  const int dq = (d + 3) / 4, dsx = 4 * dq + 4;
  float* gt = sm;                          // [DW_TILE][R]
  float* xt = sm + DW_TILE * R;            // [DW_TILE][dsx]
  const long long M = (long long)a.N * K;
  const long long m0 = (long long)blockIdx.x * chunk;
  const long long m1 = m0 + chunk < M ? m0 + chunk : M;
  const int rt = R / 4, ntiles = rt * dq;
  // each thread owns up to DW_TPT output tiles (4 rows x 4 columns): R/4 x ceil(d/4) <= 32 x 32 tiles
  float acc[DW_TPT][4][4];
#pragma unroll
  for (int u = 0; u < DW_TPT; ++u)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[u][i][j] = 0.f;
  int trow[DW_TPT], tcol[DW_TPT];
  bool tval[DW_TPT];
#pragma unroll
  for (int u = 0; u < DW_TPT; ++u) {
    const int tile = threadIdx.x + 256 * u;
    tval[u] = tile < ntiles;
    const int tt = tval[u] ? tile : 0;
    trow[u] = (tt / dq) * 4;
    tcol[u] = (tt % dq) * 4;
  }
  for (long long mb = m0; mb < m1; mb += DW_TILE) {
    const int rows = (int)(m1 - mb < DW_TILE ? m1 - mb : DW_TILE);
    __syncthreads();
    for (int i = threadIdx.x; i < DW_TILE * R; i += 256) {
      const int r = i / R;
      gt[i] = r < rows ? __ldg(dG + (size_t)mb * R + i) : 0.f;
    }
    for (int i = threadIdx.x; i < DW_TILE * dsx; i += 256) {
      const int r = i / dsx, c = i - r * dsx;
      float val = 0.f;
      if (r < rows && c < d) {
        const long long m = mb + r;
        const long long n = m / K;
        const int t = (int)(m - n * K);
        val = __ldg(a.x + (size_t)n * a.x_node_stride + (size_t)t * a.x_hop_stride + c);
      }
      xt[i] = val;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < DW_TPT; ++u) {
      if (!tval[u]) continue;
      for (int r = 0; r < DW_TILE; ++r) {
        const float4 g = *reinterpret_cast<const float4*>(gt + r * R + trow[u]);
        const float4 x = *reinterpret_cast<const float4*>(xt + r * dsx + tcol[u]);
        const float gv[4] = {g.x, g.y, g.z, g.w};
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[u][i][j] = fmaf(gv[i], xv[j], acc[u][i][j]);
      }
    }
  }
  (void)vec_in;
  float* p = part + (size_t)blockIdx.x * R * d;
#pragma unroll
  for (int u = 0; u < DW_TPT; ++u) {
    if (!tval[u]) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (tcol[u] + j < d) p[(size_t)(trow[u] + i) * d + tcol[u] + j] = acc[u][i][j];
  }
}

// out[i] = sum over parts (fixed order) of part[p][i]; routes global row Rg of [R][cols] to direction dir = Rg / 4K
__global__ void attn_reduce_kernel(const float* __restrict__ part, int nparts, long long stride, int n, float* out0,
                                   float* out1, int split) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int p = 0;
  for (; p + 4 <= nparts; p += 4) {
    s0 += part[(size_t)p * stride + i];
    s1 += part[(size_t)(p + 1) * stride + i];
    s2 += part[(size_t)(p + 2) * stride + i];
    s3 += part[(size_t)(p + 3) * stride + i];
  }
  for (; p < nparts; ++p) s0 += part[(size_t)p * stride + i];
  const float s = (s0 + s1) + (s2 + s3);
  if (i < split) out0[i] = s;
  else out1[i - split] = s;
}

struct AttnCfg {
  int KK, grid_f, grid_b, chunks, chunk, warps_f, warps_b;
  size_t smem_f, smem_b, smem_dw;
  size_t o_dG, o_wpart, o_spart, ws_total;
  int vec_in;
};

static int attn_config(const kp_attn_desc& a, AttnCfg* c) {
  KP_CHECK_ARG(a.N >= 0 && a.K >= 1 && a.K <= 16 && a.d >= 1 && a.d <= 128,
               "kp_attn_combine: need 1 <= K <= 16 and 1 <= d <= 128 (got K=%d d=%d)", a.K, a.d);
  KP_CHECK_ARG(a.x && a.w_ih[0] && a.w_ih[1] && a.w_hh[0] && a.w_hh[1] && a.b_ih[0] && a.b_ih[1] && a.b_hh[0] && a.b_hh[1],
               "kp_attn_combine: null argument");
  KP_CHECK_ARG(a.x_node_stride >= 0 && a.x_hop_stride >= a.d, "kp_attn_combine: bad strides");
  c->KK = a.K <= 4 ? 4 : (a.K <= 8 ? 8 : 16);
  c->warps_f = c->warps_b = AT_MAX_WARPS;
  while (c->warps_f > 1 && sizeof(float) * (size_t)attn_layout(a.K, a.d, false, c->warps_f).total_floats > 100 * 1024)
    c->warps_f >>= 1;
  while (c->warps_b > 1 && sizeof(float) * (size_t)attn_layout(a.K, a.d, true, c->warps_b).total_floats > 200 * 1024)
    c->warps_b >>= 1;
  c->smem_f = sizeof(float) * (size_t)attn_layout(a.K, a.d, false, c->warps_f).total_floats;
  c->smem_b = sizeof(float) * (size_t)attn_layout(a.K, a.d, true, c->warps_b).total_floats;
  KP_CHECK_ARG(c->smem_f <= 220 * 1024 && c->smem_b <= 220 * 1024,
               "kp_attn_combine: K=%d d=%d needs %zu bytes of shared memory", a.K, a.d, c->smem_b);
  const long long want_f = ((long long)a.N + c->warps_f - 1) / c->warps_f;
  const long long want_b = ((long long)a.N + c->warps_b - 1) / c->warps_b;
  const int per_sm_f = (int)((200 * 1024) / (c->smem_f ? c->smem_f : 1));
  const int per_sm_b = (int)((200 * 1024) / (c->smem_b ? c->smem_b : 1));
  const long long cap_f = (long long)kNumSMs * (per_sm_f < 1 ? 1 : (per_sm_f > 4 ? 4 : per_sm_f));
  const long long cap_b = (long long)kNumSMs * (per_sm_b < 1 ? 1 : (per_sm_b > 4 ? 4 : per_sm_b));
  c->grid_f = (int)(want_f < 1 ? 1 : (want_f > cap_f ? cap_f : want_f));
  c->grid_b = (int)(want_b < 1 ? 1 : (want_b > cap_b ? cap_b : want_b));
  const long long M = (long long)a.N * a.K;
  long long chunks = (M + 4 * DW_TILE - 1) / (4 * DW_TILE);      // >= 128 rows per CTA
  if (chunks > 2 * kNumSMs) chunks = 2 * kNumSMs;
  if (chunks < 1) chunks = 1;
  long long chunk = (M + chunks - 1) / chunks;
  chunk = (chunk + DW_TILE - 1) / DW_TILE * DW_TILE;
  if (chunk < DW_TILE) chunk = DW_TILE;
  chunks = M > 0 ? (M + chunk - 1) / chunk : 1;
  c->chunks = (int)chunks;
  c->chunk = (int)chunk;
  const int R = 8 * a.K, dq = (a.d + 3) / 4;
  c->smem_dw = sizeof(float) * ((size_t)DW_TILE * R + (size_t)DW_TILE * (4 * dq + 4));
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 256);
    return o;
  };
  c->o_dG = take(sizeof(float) * (size_t)M * R);
  c->o_wpart = take(sizeof(float) * (size_t)c->chunks * R * a.d);
  c->o_spart = take(sizeof(float) * (size_t)c->grid_b * (R * a.K + R));
  c->ws_total = off;
  c->vec_in = (a.d % 4 == 0 && a.x_node_stride % 4 == 0 && a.x_hop_stride % 4 == 0 && (((uintptr_t)a.x) & 15) == 0) ? 1 : 0;
  return 0;
}

}  // namespace kp

extern "C" {

int kp_attn_combine_forward(const kp_attn_desc* desc, float* out, float* weights, void* stream) {
  KP_CHECK_ARG(desc && out, "kp_attn_combine_forward: null argument");
  const kp_attn_desc& a = *desc;
  kp::AttnCfg c;
  if (kp::attn_config(a, &c)) return 1;
  if (a.N == 0) return 0;
  const int vec_out = (a.d % 4 == 0 && (((uintptr_t)out) & 15) == 0) ? 1 : 0;
#define KP_ATF(KKV)                                                                                              \
  do {                                                                                                           \
    if (c.smem_f > 32 * 1024)                                                                                    \
      KP_CUDA(cudaFuncSetAttribute(kp::attn_fwd_kernel<KKV>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                   (int)c.smem_f));                                                              \
    KP_LAUNCH(kp::attn_fwd_kernel<KKV>, c.grid_f, 32 * c.warps_f, c.smem_f, stream, a, out, weights, c.vec_in,    \
              vec_out);                                                                                          \
  } while (0)
  if (c.KK == 4) KP_ATF(4);
  else if (c.KK == 8) KP_ATF(8);
  else KP_ATF(16);
#undef KP_ATF
  return 0;
}

int kp_attn_combine_backward_workspace_bytes(const kp_attn_desc* desc, size_t* bytes) {
  KP_CHECK_ARG(desc && bytes, "kp_attn_combine_backward_workspace_bytes: null argument");
  kp::AttnCfg c;
  if (kp::attn_config(*desc, &c)) return 1;
  *bytes = c.ws_total;
  return 0;
}

int kp_attn_combine_backward(const kp_attn_desc* desc, const float* dOut, float* dX, float* dw_ih_f, float* dw_ih_r,
                             float* dw_hh_f, float* dw_hh_r, float* db_f, float* db_r, void* workspace,
                             size_t workspace_bytes, void* stream) {
  KP_CHECK_ARG(desc && dOut && dX && dw_ih_f && dw_ih_r && dw_hh_f && dw_hh_r && db_f && db_r,
               "kp_attn_combine_backward: null argument");
  const kp_attn_desc& a = *desc;
  kp::AttnCfg c;
  if (kp::attn_config(a, &c)) return 1;
  KP_CHECK_ARG(workspace_bytes >= c.ws_total && (workspace || c.ws_total == 0) && (((uintptr_t)workspace) & 15) == 0,
               "kp_attn_combine_backward: workspace too small or misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int R = 8 * a.K, R4 = 4 * a.K;
  if (a.N == 0) {
    KP_CUDA(cudaMemsetAsync(dw_ih_f, 0, sizeof(float) * R4 * a.d, st));
    KP_CUDA(cudaMemsetAsync(dw_ih_r, 0, sizeof(float) * R4 * a.d, st));
    KP_CUDA(cudaMemsetAsync(dw_hh_f, 0, sizeof(float) * R4 * a.K, st));
    KP_CUDA(cudaMemsetAsync(dw_hh_r, 0, sizeof(float) * R4 * a.K, st));
    KP_CUDA(cudaMemsetAsync(db_f, 0, sizeof(float) * R4, st));
    KP_CUDA(cudaMemsetAsync(db_r, 0, sizeof(float) * R4, st));
    return 0;
  }
  char* ws = (char*)workspace;
  float* dG = (float*)(ws + c.o_dG);
  float* wpart = (float*)(ws + c.o_wpart);
  float* spart = (float*)(ws + c.o_spart);
  const int vec_out = (a.d % 4 == 0 && (((uintptr_t)dX) & 15) == 0) ? 1 : 0;
#define KP_ATB(KKV)                                                                                              \
  do {                                                                                                           \
    if (c.smem_b > 32 * 1024)                                                                                    \
      KP_CUDA(cudaFuncSetAttribute(kp::attn_bwd_kernel<KKV>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                   (int)c.smem_b));                                                              \
    KP_LAUNCH(kp::attn_bwd_kernel<KKV>, c.grid_b, 32 * c.warps_b, c.smem_b, st, a, dOut, dX, dG, spart, c.vec_in, \
              vec_out);                                                                                          \
  } while (0)
  if (c.KK == 4) KP_ATB(4);
  else if (c.KK == 8) KP_ATB(8);
  else KP_ATB(16);
#undef KP_ATB
  if (c.smem_dw > 32 * 1024)
    KP_CUDA(cudaFuncSetAttribute(kp::attn_dwih_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem_dw));
  KP_LAUNCH(kp::attn_dwih_kernel, c.chunks, 256, c.smem_dw, st, a, dG, wpart, c.chunk, c.vec_in);
  {
    const int n = R * a.d;
    KP_LAUNCH(kp::attn_reduce_kernel, kp::ceil_div(n, 256), 256, 0, st, wpart, c.chunks, (long long)n, n, dw_ih_f,
              dw_ih_r, R4 * a.d);
  }
  {
    const long long stride = (long long)R * a.K + R;
    const int n = R * a.K;
    KP_LAUNCH(kp::attn_reduce_kernel, kp::ceil_div(n, 256), 256, 0, st, spart, c.grid_b, stride, n, dw_hh_f, dw_hh_r,
              R4 * a.K);
    KP_LAUNCH(kp::attn_reduce_kernel, kp::ceil_div(R, 256), 256, 0, st, spart + n, c.grid_b, stride, R, db_f, db_r, R4);
  }
  return 0;
}

}  // extern "C"
