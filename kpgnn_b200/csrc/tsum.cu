// tsum.cu -- multi-table gather-sum and its deterministic gradient (sm_100a).
//
//   out[r,:]   = sum_{s<S} table[slot_off[s] + idx[r,s], :]
//   dTable[t,:] = sum_{(r,s): slot_off[s]+idx[r,s] == t} dOut[r,:]
//
// This is the peripheral-attribute embedding stage of the reference's backbones (models/GNNs.py:172-179,
// 393-400 with layers/feature_encoder.py:37-67) after folding each embedding table through its slice of the
// encoder's Linear (M_i = E_i W_i^T, done by the caller as tiny GEMMs): the [N,K,c,2H] concatenated embeddings,
// the [N,K,c,H] projected tensor and the sort-based embedding backward of the reference never exist; P [N,K,H]
// is written once and its gradient is read once.  HBM-bound row streaming; no float atomics: the gradient uses
// group-private sub-tables in shared memory and fixed-order reductions.
#include "agg_common.cuh"

namespace kp {

__device__ __forceinline__ float4 tld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int G>
__global__ void __launch_bounds__(256) tsum_fwd_kernel(const kp_tsum_desc t, const float* __restrict__ table,
                                                       float* __restrict__ out) {
  const int lane = threadIdx.x & (G - 1);
  const int c = min(lane * 4, t.d - 4);
  const bool active = lane * 4 < t.d;
  constexpr int gpb = 256 / G;
  const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  for (long long r = (long long)blockIdx.x * gpb + threadIdx.x / G; r < t.R; r += (long long)gridDim.x * gpb) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s0 = 0; s0 < t.S; s0 += G) {
      int row = 0;
      if (s0 + lane < t.S) row = t.slot_off[s0 + lane] + (int)__ldg(t.idx + r * t.S + s0 + lane);
      const int n = min(G, t.S - s0);
      for (int q = 0; q < n; ++q) {
        const int tr = __shfl_sync(gm, row, q, G);
        const float4 v = tld4(table + (size_t)tr * t.d + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (active) __stcs(reinterpret_cast<float4*>(out + r * t.d + c), acc);
  }
}

// grid B*Q: CTA (b,q) folds row chunk b into the table rows [t0,t1) that slots [s0,s1) address.  256 threads
// zero / flush the sub-tables; the first ngroups*G of them walk the rows.
template <int G>
__global__ void __launch_bounds__(256)
tsum_bwd_kernel(const kp_tsum_desc t, const float* __restrict__ dOut, int ngroups, int rows_per_group,
                float* __restrict__ part) {
  extern __shared__ __align__(16) float smem[];
  const int q = blockIdx.x % t.num_ranges, bx = blockIdx.x / t.num_ranges;
  const int s0 = t.range_slot[q], s1 = t.range_slot[q + 1];
  const int t0 = t.range_row[q], t1 = t.range_row[q + 1];
  const int tsz = (t1 - t0) * t.d;
  for (int i = threadIdx.x * 4; i < tsz * ngroups; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(smem + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int lane = threadIdx.x & (G - 1);
  const int gib = threadIdx.x / G;
  const int c = min(lane * 4, t.d - 4);
  const bool active = lane * 4 < t.d;
  const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  const long long gid = (long long)bx * ngroups + gib;
  const long long r0 = gid * rows_per_group, r1 = min((long long)t.R, r0 + rows_per_group);
  float* tab = smem + (size_t)tsz * (gib < ngroups ? gib : 0) + c;
  const int ns = s1 - s0;                      // <= 32 slots, one per lane (host guarantees ns <= G)
  if (gib < ngroups) {
    constexpr int RB = 8;                    // rows in flight per group
    for (long long rb = r0; rb < r1; rb += RB) {
      float4 g[RB];
      int row[RB];
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        row[u] = 0;
        if (rb + u < r1) {
          g[u] = __ldcs(reinterpret_cast<const float4*>(dOut + (rb + u) * t.d + c));
          if (lane < ns) row[u] = t.slot_off[s0 + lane] + (int)__ldg(t.idx + (rb + u) * t.S + s0 + lane) - t0;
        }
      }
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        if (rb + u < r1) {
          for (int k = 0; k < ns; ++k) {
            const int tr = __shfl_sync(gm, row[u], k, G);
            if (active) {
              float4* dst = reinterpret_cast<float4*>(tab + (size_t)tr * t.d);
              float4 v = *dst;
              v.x += g[u].x; v.y += g[u].y; v.z += g[u].z; v.w += g[u].w;
              *dst = v;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  float* dstp = part + (size_t)bx * t.table_rows * t.d + (size_t)t0 * t.d;
  for (int i = threadIdx.x; i < tsz; i += blockDim.x) {
    float s = 0.f;
    for (int g = 0; g < ngroups; ++g) s += smem[(size_t)tsz * g + i];
    dstp[i] = s;
  }
}

__global__ void tsum_reduce_kernel(const float* __restrict__ part, int nblocks, int n, float* __restrict__ out) {
  // one warp per output element, fixed order (see reduce_partials_kernel in agg.cu)
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  float s = 0.f;
  for (int b = lane; b < nblocks; b += 32) s += __ldcs(part + (size_t)b * n + i);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[i] = s;
}

// tile-sorted register-accumulation gradient (tsum_sorted.cu)
size_t ts_sorted_workspace_bytes(const kp_tsum_desc& t);
int ts_sorted_backward(const kp_tsum_desc& t, const float* dOut, float* dTable, void* workspace, size_t workspace_bytes,
                       cudaStream_t st);

static size_t g_tsum_smem_cap = 200 * 1024;   // profiling hook: kp_table_sum_set_smem_cap

struct TsumCfg {
  int G, grid_fwd, B, Q, ngroups, threads, rows_per_group;
  size_t smem;
};

static int tsum_config(const kp_tsum_desc& t, TsumCfg* c) {
  KP_CHECK_ARG(t.R >= 0 && t.S >= 1 && t.S <= 32 && t.d >= 4 && t.d <= 128 && t.d % 4 == 0 && t.table_rows >= 1,
               "kp_table_sum: need 1 <= S <= 32, d %% 4 == 0, 4 <= d <= 128 (got S=%d d=%d)", t.S, t.d);
  KP_CHECK_ARG(t.idx, "kp_table_sum: null idx");
  KP_CHECK_ARG(t.num_ranges >= 1 && t.num_ranges <= 16 && t.range_slot[0] == 0 && t.range_slot[t.num_ranges] == t.S &&
                   t.range_row[0] == 0 && t.range_row[t.num_ranges] == t.table_rows,
               "kp_table_sum: slot/row ranges must partition the slots and the table");
  int lanes = t.d / 4, G = 4;
  while (G < lanes) G <<= 1;
  c->G = G;
  const int gpb = 256 / G;
  long long want = ((long long)t.R + gpb - 1) / gpb;
  c->grid_fwd = (int)(want < 1 ? 1 : (want > kNumSMs * 8 ? kNumSMs * 8 : want));
  c->Q = t.num_ranges;
  size_t maxsub = 0;
  for (int q = 0; q < t.num_ranges; ++q) {
    KP_CHECK_ARG(t.range_slot[q + 1] - t.range_slot[q] <= G, "kp_table_sum: a slot range exceeds the group width");
    size_t sub = sizeof(float) * (size_t)(t.range_row[q + 1] - t.range_row[q]) * t.d;
    if (sub > maxsub) maxsub = sub;
  }
  KP_CHECK_ARG(maxsub <= 200 * 1024, "kp_table_sum: a table range needs %zu bytes of shared memory (> 200 KB)", maxsub);
  int ng = 256 / G;
  while (ng > 1 && maxsub * ng > g_tsum_smem_cap) ng >>= 1;
  c->ngroups = ng;
  c->threads = 256;
  c->smem = maxsub * ng;
  long long wantg = ((long long)t.R + 63) / 64;           // >= 64 rows per group
  long long B = (wantg + ng - 1) / ng;
  // one CTA per SM over the B*Q grid (the sub-tables fill an SM's shared memory: a second wave would only
  // repeat the zero / flush of ~190 KB and double the partial tables)
  const long long maxB = kNumSMs / c->Q > 0 ? kNumSMs / c->Q : 1;
  if (B > maxB) B = maxB;
  if (B < 1) B = 1;
  c->B = (int)B;
  const long long tg = B * ng;
  c->rows_per_group = (int)((t.R + tg - 1) / tg);
  return 0;
}

}  // namespace kp

extern "C" {

int kp_table_sum_set_smem_cap(size_t bytes) {
  kp::g_tsum_smem_cap = bytes < 16 * 1024 ? 16 * 1024 : (bytes > 200 * 1024 ? 200 * 1024 : bytes);
  return 0;
}

int kp_table_sum_forward(const kp_tsum_desc* desc, const float* table, float* out, void* stream) {
  KP_CHECK_ARG(desc && table && out, "kp_table_sum_forward: null argument");
  kp::TsumCfg c;
  if (kp::tsum_config(*desc, &c)) return 1;
  if (desc->R == 0) return 0;
  KP_CHECK_ARG((((uintptr_t)table | (uintptr_t)out) & 15) == 0, "kp_table_sum_forward: table/out must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  switch (c.G) {
    case 32: KP_LAUNCH(kp::tsum_fwd_kernel<32>, c.grid_fwd, 256, 0, st, *desc, table, out); break;
    case 16: KP_LAUNCH(kp::tsum_fwd_kernel<16>, c.grid_fwd, 256, 0, st, *desc, table, out); break;
    case 8:  KP_LAUNCH(kp::tsum_fwd_kernel<8>, c.grid_fwd, 256, 0, st, *desc, table, out); break;
    default: KP_LAUNCH(kp::tsum_fwd_kernel<4>, c.grid_fwd, 256, 0, st, *desc, table, out); break;
  }
  return 0;
}

int kp_table_sum_backward_workspace_bytes(const kp_tsum_desc* desc, size_t* bytes) {
  KP_CHECK_ARG(desc && bytes, "kp_table_sum_backward_workspace_bytes: null argument");
  kp::TsumCfg c;
  if (kp::tsum_config(*desc, &c)) return 1;
  *bytes = sizeof(float) * (size_t)c.B * desc->table_rows * desc->d;
  const size_t sorted = kp::ts_sorted_workspace_bytes(*desc);
  if (sorted > *bytes) *bytes = sorted;
  return 0;
}

int kp_table_sum_backward(const kp_tsum_desc* desc, const float* dOut, float* dTable, void* workspace,
                          size_t workspace_bytes, void* stream) {
  KP_CHECK_ARG(desc && dOut && dTable, "kp_table_sum_backward: null argument");
  kp::TsumCfg c;
  if (kp::tsum_config(*desc, &c)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)desc->table_rows * desc->d;
  if (desc->R == 0) {
    KP_CUDA(cudaMemsetAsync(dTable, 0, sizeof(float) * n, st));
    return 0;
  }
  {
    const int rc = kp::ts_sorted_backward(*desc, dOut, dTable, workspace, workspace_bytes, st);
    if (rc >= 0) return rc;
  }
  KP_CHECK_ARG(workspace && workspace_bytes >= sizeof(float) * (size_t)c.B * n, "kp_table_sum_backward: workspace too small");
  KP_CHECK_ARG((((uintptr_t)dOut | (uintptr_t)workspace) & 15) == 0, "kp_table_sum_backward: dOut/workspace alignment");
  float* part = (float*)workspace;
  const int grid = c.B * c.Q;
#define KP_TSB(GG)                                                                                             \
  do {                                                                                                         \
    if (c.smem > 32 * 1024)                                                                                    \
      KP_CUDA(cudaFuncSetAttribute(kp::tsum_bwd_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem)); \
    KP_LAUNCH(kp::tsum_bwd_kernel<GG>, grid, c.threads, c.smem, st, *desc, dOut, c.ngroups, c.rows_per_group, part); \
  } while (0)
  switch (c.G) {
    case 32: KP_TSB(32); break;
    case 16: KP_TSB(16); break;
    case 8:  KP_TSB(8); break;
    default: KP_TSB(4); break;
  }
#undef KP_TSB
  KP_LAUNCH(kp::tsum_reduce_kernel, kp::ceil_div((long long)n * 32, 256), 256, 0, st, part, c.B, (int)n, dTable);
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// GeometricCombine weights (combine.py:51-58): theta[h,c] = softmax_h( a_c (1-a_c)^h ),  a = sigmoid(alphas).
// The reference builds them with ~8 elementwise launches per layer (and ~12 more in backward) on a [K,d] tensor;
// here forward and backward are one single-CTA kernel each.
// ------------------------------------------------------------------------------------------------------------
namespace kp {

__global__ void geo_theta_fwd_kernel(const float* __restrict__ alphas, int K, int d, float* __restrict__ theta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const float a = 1.f / (1.f + expf(-alphas[c]));
  float m = -INFINITY, pw = 1.f;
  for (int h = 0; h < K; ++h) {
    m = fmaxf(m, a * pw);
    pw *= (1.f - a);
  }
  float sum = 0.f;
  pw = 1.f;
  for (int h = 0; h < K; ++h) {
    sum += expf(a * pw - m);
    pw *= (1.f - a);
  }
  pw = 1.f;
  const float inv = 1.f / sum;
  for (int h = 0; h < K; ++h) {
    theta[(size_t)h * d + c] = expf(a * pw - m) * inv;
    pw *= (1.f - a);
  }
}

__global__ void geo_theta_bwd_kernel(const float* __restrict__ alphas, const float* __restrict__ theta,
                                     const float* __restrict__ dtheta, int K, int d, float* __restrict__ dalphas) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const float a = 1.f / (1.f + expf(-alphas[c]));
  float dot = 0.f;
  for (int h = 0; h < K; ++h) dot += theta[(size_t)h * d + c] * dtheta[(size_t)h * d + c];
  float da = 0.f, pw = 1.f, pwm1 = 0.f;          // pw = (1-a)^h, pwm1 = (1-a)^(h-1)
  for (int h = 0; h < K; ++h) {
    const float dt = theta[(size_t)h * d + c] * (dtheta[(size_t)h * d + c] - dot);
    da += dt * (pw - a * (float)h * pwm1);
    pwm1 = pw;
    pw *= (1.f - a);
  }
  dalphas[c] = da * a * (1.f - a);
}

}  // namespace kp

extern "C" {

int kp_geometric_theta_forward(const float* alphas, int32_t K, int32_t d, float* theta, void* stream) {
  KP_CHECK_ARG(alphas && theta && K >= 1 && d >= 1, "kp_geometric_theta_forward: bad arguments");
  KP_LAUNCH(kp::geo_theta_fwd_kernel, kp::ceil_div(d, 128), 128, 0, (cudaStream_t)stream, alphas, K, d, theta);
  return 0;
}

int kp_geometric_theta_backward(const float* alphas, const float* theta, const float* dtheta, int32_t K, int32_t d,
                                float* dalphas, void* stream) {
  KP_CHECK_ARG(alphas && theta && dtheta && dalphas && K >= 1 && d >= 1, "kp_geometric_theta_backward: bad arguments");
  KP_LAUNCH(kp::geo_theta_bwd_kernel, kp::ceil_div(d, 128), 128, 0, (cudaStream_t)stream, alphas, theta, dtheta, K, d,
            dalphas);
  return 0;
}

}  // extern "C"

// all layers of a stack in one launch: blockIdx.y = layer
namespace kp {
__global__ void geo_theta_fwd_batched_kernel(const kp_theta_batch b) {
  const int l = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = b.k[l], d = b.d;
  if (c >= d) return;
  const float a = 1.f / (1.f + expf(-b.alphas[l][c]));
  float m = -INFINITY, pw = 1.f;
  for (int h = 0; h < K; ++h) {
    m = fmaxf(m, a * pw);
    pw *= (1.f - a);
  }
  float sum = 0.f;
  pw = 1.f;
  for (int h = 0; h < K; ++h) {
    sum += expf(a * pw - m);
    pw *= (1.f - a);
  }
  pw = 1.f;
  const float inv = 1.f / sum;
  for (int h = 0; h < K; ++h) {
    b.theta[l][(size_t)h * d + c] = expf(a * pw - m) * inv;
    pw *= (1.f - a);
  }
}
}  // namespace kp

extern "C" int kp_geometric_theta_forward_batched(const kp_theta_batch* batch, void* stream) {
  KP_CHECK_ARG(batch && batch->L >= 0 && batch->L <= 32 && batch->d >= 1, "kp_geometric_theta_forward_batched: bad arguments");
  for (int l = 0; l < batch->L; ++l)
    KP_CHECK_ARG(batch->alphas[l] && batch->theta[l] && batch->k[l] >= 1, "kp_geometric_theta_forward_batched: layer %d", l);
  if (batch->L == 0) return 0;
  KP_LAUNCH(kp::geo_theta_fwd_batched_kernel, dim3(kp::ceil_div(batch->d, 128), batch->L), 128, 0, (cudaStream_t)stream,
            *batch);
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// kp_peripheral_grad: dP[v,h,:] = sum_{l: k_l > h} theta_l[h,:] * dAgg_l[v,:]   (see include/kpgnn.h)
// thread per (node, 4 channels): the L gradients are read once into registers, the K hop rows written once.
// ------------------------------------------------------------------------------------------------------------
namespace kp {

template <int LMAX>
__global__ void __launch_bounds__(256) pgrad_kernel(const kp_pgrad_desc p, float* __restrict__ dP) {
  const int c4n = p.d >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.N * c4n) return;
  const int v = (int)(i / c4n), c = (int)(i - (long long)v * c4n) * 4;
  float4 g[LMAX];
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    g[l] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < p.L) g[l] = __ldg(reinterpret_cast<const float4*>(p.dagg[l] + (size_t)v * p.d + c));
  }
  for (int h = 0; h < p.K; ++h) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
      if (l < p.L && p.k[l] > h) {
        float4 t = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p.theta[l]) t = __ldg(reinterpret_cast<const float4*>(p.theta[l] + (size_t)h * p.d + c));
        s.x = fmaf(t.x, g[l].x, s.x); s.y = fmaf(t.y, g[l].y, s.y);
        s.z = fmaf(t.z, g[l].z, s.z); s.w = fmaf(t.w, g[l].w, s.w);
      }
    }
    *reinterpret_cast<float4*>(dP + ((size_t)v * p.K + h) * p.d + c) = s;
  }
}

}  // namespace kp

extern "C" int kp_peripheral_grad(const kp_pgrad_desc* desc, float* dP, void* stream) {
  KP_CHECK_ARG(desc && dP, "kp_peripheral_grad: null argument");
  const kp_pgrad_desc& p = *desc;
  KP_CHECK_ARG(p.N >= 0 && p.K >= 1 && p.d >= 4 && p.d % 4 == 0 && p.L >= 1 && p.L <= 32,
               "kp_peripheral_grad: need d %% 4 == 0, 1 <= L <= 32 (got d=%d L=%d)", p.d, p.L);
  for (int l = 0; l < p.L; ++l)
    KP_CHECK_ARG(p.dagg[l] && p.k[l] >= 1 && p.k[l] <= p.K && (((uintptr_t)p.dagg[l] | (uintptr_t)p.theta[l]) & 15) == 0,
                 "kp_peripheral_grad: layer %d: null / misaligned gradient or bad hop count", l);
  KP_CHECK_ARG(((uintptr_t)dP & 15) == 0, "kp_peripheral_grad: dP must be 16-byte aligned");
  if (p.N == 0) return 0;
  const long long n = (long long)p.N * (p.d >> 2);
  cudaStream_t st = (cudaStream_t)stream;
  if (p.L <= 8) KP_LAUNCH(kp::pgrad_kernel<8>, kp::ceil_div(n, 256), 256, 0, st, p, dP);
  else if (p.L <= 16) KP_LAUNCH(kp::pgrad_kernel<16>, kp::ceil_div(n, 256), 256, 0, st, p, dP);
  else KP_LAUNCH(kp::pgrad_kernel<32>, kp::ceil_div(n, 256), 256, 0, st, p, dP);
  return 0;
}
