// bn.cu -- training-mode BatchNorm1d (+ optional ReLU) over [N, C] fp32 node features, forward and backward, one
// kernel each (sm_100a).
//
// The KP-GIN+ layer is  aggregate -> Linear -> BN -> ReLU -> Linear -> BN -> ReLU  followed by the backbone's
// BatchNorm (layers/KPGINplus.py:25-30, models/GNNs.py:430).  For the node counts of a molecule batch (N of a few
// thousand, C ~ 100) PyTorch runs each BN as 3 kernels forward (statistics, transform, running-stat update) and 2
// backward, plus 2 for the ReLU: ~24 % of the step after the aggregation itself was fused (profiles/r1j).
// Here a CTA owns 4 channels; thread t keeps rows t, t+256, ... of its channel quad in registers (N <= 4096), so x
// is read from HBM/L2 exactly once: mean, then centred variance (two-pass, in registers), normalise, ReLU, write,
// and the running statistics.  Fixed reduction tree -> bit-reproducible.
#include "common.cuh"

namespace kp {

constexpr int BN_ROWS_MAX = 16;   // rows per thread held in registers -> N <= 256 * 16

__device__ __forceinline__ float4 block_sum4(float4 v, float4* red) {   // blockDim.x == 256; red: 8 float4 in smem
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
    v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float4 s = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) {
    s.x += red[w].x; s.y += red[w].y; s.z += red[w].z; s.w += red[w].w;
  }
  return s;
}

template <bool RELU>
__global__ void __launch_bounds__(256)
bn_fwd_kernel(const float* __restrict__ x, int N, int C, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
              float* __restrict__ running_var, long long* __restrict__ num_batches, float* __restrict__ y,
              float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  __shared__ float4 red[8];
  const int c = blockIdx.x * 4;
  float4 v[BN_ROWS_MAX];
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < BN_ROWS_MAX; ++i) {
    const int r = threadIdx.x + i * 256;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < N) v[i] = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * C + c));
    s.x += v[i].x; s.y += v[i].y; s.z += v[i].z; s.w += v[i].w;
  }
  s = block_sum4(s, red);
  const float invn = 1.f / (float)N;
  const float4 mean = make_float4(s.x * invn, s.y * invn, s.z * invn, s.w * invn);
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < BN_ROWS_MAX; ++i) {
    const int r = threadIdx.x + i * 256;
    if (r < N) {
      const float dx = v[i].x - mean.x, dy = v[i].y - mean.y, dz = v[i].z - mean.z, dw = v[i].w - mean.w;
      q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
    }
  }
  q = block_sum4(q, red);
  const float4 var = make_float4(q.x * invn, q.y * invn, q.z * invn, q.w * invn);          // biased, as BN uses
  const float4 istd = make_float4(rsqrtf(var.x + eps), rsqrtf(var.y + eps), rsqrtf(var.z + eps), rsqrtf(var.w + eps));
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
  const float4 sc = make_float4(g.x * istd.x, g.y * istd.y, g.z * istd.z, g.w * istd.w);
#pragma unroll
  for (int i = 0; i < BN_ROWS_MAX; ++i) {
    const int r = threadIdx.x + i * 256;
    if (r < N) {
      float4 o = make_float4(fmaf(v[i].x - mean.x, sc.x, b.x), fmaf(v[i].y - mean.y, sc.y, b.y),
                             fmaf(v[i].z - mean.z, sc.z, b.z), fmaf(v[i].w - mean.w, sc.w, b.w));
      if (RELU) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
      }
      *reinterpret_cast<float4*>(y + (size_t)r * C + c) = o;
    }
  }
  if (threadIdx.x == 0) {
    *reinterpret_cast<float4*>(save_mean + c) = mean;
    *reinterpret_cast<float4*>(save_invstd + c) = istd;
    if (running_mean) {
      const float unb = N > 1 ? (float)N / (float)(N - 1) : 1.f;
      float4 rm = *reinterpret_cast<float4*>(running_mean + c), rv = *reinterpret_cast<float4*>(running_var + c);
      rm.x = fmaf(momentum, mean.x - rm.x, rm.x); rm.y = fmaf(momentum, mean.y - rm.y, rm.y);
      rm.z = fmaf(momentum, mean.z - rm.z, rm.z); rm.w = fmaf(momentum, mean.w - rm.w, rm.w);
      rv.x = fmaf(momentum, var.x * unb - rv.x, rv.x); rv.y = fmaf(momentum, var.y * unb - rv.y, rv.y);
      rv.z = fmaf(momentum, var.z * unb - rv.z, rv.z); rv.w = fmaf(momentum, var.w * unb - rv.w, rv.w);
      *reinterpret_cast<float4*>(running_mean + c) = rm;
      *reinterpret_cast<float4*>(running_var + c) = rv;
    }
    if (num_batches && blockIdx.x == 0) *num_batches += 1;
  }
}

template <bool RELU>
__global__ void __launch_bounds__(256)
bn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int N, int C,
              const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
              const float* __restrict__ save_invstd, float* __restrict__ dx, float* __restrict__ dgamma,
              float* __restrict__ dbeta) {
  __shared__ float4 red[8];
  const int c = blockIdx.x * 4;
  const float4 mean = __ldg(reinterpret_cast<const float4*>(save_mean + c));
  const float4 istd = __ldg(reinterpret_cast<const float4*>(save_invstd + c));
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
  float4 xh[BN_ROWS_MAX], gr[BN_ROWS_MAX];
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll
  for (int i = 0; i < BN_ROWS_MAX; ++i) {
    const int r = threadIdx.x + i * 256;
    xh[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    gr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < N) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * C + c));
      float4 dv = __ldg(reinterpret_cast<const float4*>(dy + (size_t)r * C + c));
      xh[i] = make_float4((xv.x - mean.x) * istd.x, (xv.y - mean.y) * istd.y, (xv.z - mean.z) * istd.z,
                          (xv.w - mean.w) * istd.w);
      if (RELU) {          // y = gamma * xhat + beta; gradient passes where y > 0
        if (fmaf(g.x, xh[i].x, b.x) <= 0.f) dv.x = 0.f;
        if (fmaf(g.y, xh[i].y, b.y) <= 0.f) dv.y = 0.f;
        if (fmaf(g.z, xh[i].z, b.z) <= 0.f) dv.z = 0.f;
        if (fmaf(g.w, xh[i].w, b.w) <= 0.f) dv.w = 0.f;
      }
      gr[i] = dv;
      s1.x += dv.x; s1.y += dv.y; s1.z += dv.z; s1.w += dv.w;
      s2.x = fmaf(dv.x, xh[i].x, s2.x); s2.y = fmaf(dv.y, xh[i].y, s2.y);
      s2.z = fmaf(dv.z, xh[i].z, s2.z); s2.w = fmaf(dv.w, xh[i].w, s2.w);
    }
  }
  s1 = block_sum4(s1, red);
  s2 = block_sum4(s2, red);
  const float invn = 1.f / (float)N;
  const float4 k = make_float4(g.x * istd.x, g.y * istd.y, g.z * istd.z, g.w * istd.w);
  const float4 m1 = make_float4(s1.x * invn, s1.y * invn, s1.z * invn, s1.w * invn);
  const float4 m2 = make_float4(s2.x * invn, s2.y * invn, s2.z * invn, s2.w * invn);
#pragma unroll
  for (int i = 0; i < BN_ROWS_MAX; ++i) {
    const int r = threadIdx.x + i * 256;
    if (r < N) {
      float4 o = make_float4(k.x * (gr[i].x - m1.x - xh[i].x * m2.x), k.y * (gr[i].y - m1.y - xh[i].y * m2.y),
                             k.z * (gr[i].z - m1.z - xh[i].z * m2.z), k.w * (gr[i].w - m1.w - xh[i].w * m2.w));
      *reinterpret_cast<float4*>(dx + (size_t)r * C + c) = o;
    }
  }
  if (threadIdx.x == 0) {
    *reinterpret_cast<float4*>(dgamma + c) = s2;
    *reinterpret_cast<float4*>(dbeta + c) = s1;
  }
}

}  // namespace kp

extern "C" {

int kp_bn_max_rows(void) { return 256 * kp::BN_ROWS_MAX; }

int kp_bn_forward(const float* x, int32_t N, int32_t C, const float* gamma, const float* beta, float eps,
                  float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked, int32_t relu,
                  float* y, float* save_mean, float* save_invstd, void* stream) {
  KP_CHECK_ARG(x && gamma && beta && y && save_mean && save_invstd, "kp_bn_forward: null argument");
  KP_CHECK_ARG(N >= 1 && N <= 256 * kp::BN_ROWS_MAX && C >= 4 && C % 4 == 0,
               "kp_bn_forward: needs 1 <= N <= %d and C %% 4 == 0 (got N=%d C=%d)", 256 * kp::BN_ROWS_MAX, N, C);
  KP_CHECK_ARG(((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)save_mean |
                  (uintptr_t)save_invstd | (uintptr_t)running_mean | (uintptr_t)running_var) & 15) == 0),
               "kp_bn_forward: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (relu) KP_LAUNCH(kp::bn_fwd_kernel<true>, C / 4, 256, 0, st, x, N, C, gamma, beta, eps, momentum, running_mean,
                      running_var, (long long*)num_batches_tracked, y, save_mean, save_invstd);
  else      KP_LAUNCH(kp::bn_fwd_kernel<false>, C / 4, 256, 0, st, x, N, C, gamma, beta, eps, momentum, running_mean,
                      running_var, (long long*)num_batches_tracked, y, save_mean, save_invstd);
  return 0;
}

int kp_bn_backward(const float* x, const float* dy, int32_t N, int32_t C, const float* gamma, const float* beta,
                   const float* save_mean, const float* save_invstd, int32_t relu, float* dx, float* dgamma,
                   float* dbeta, void* stream) {
  KP_CHECK_ARG(x && dy && gamma && beta && save_mean && save_invstd && dx && dgamma && dbeta,
               "kp_bn_backward: null argument");
  KP_CHECK_ARG(N >= 1 && N <= 256 * kp::BN_ROWS_MAX && C >= 4 && C % 4 == 0, "kp_bn_backward: bad sizes");
  KP_CHECK_ARG(((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)dgamma | (uintptr_t)dbeta) & 15) == 0),
               "kp_bn_backward: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (relu) KP_LAUNCH(kp::bn_bwd_kernel<true>, C / 4, 256, 0, st, x, dy, N, C, gamma, beta, save_mean, save_invstd, dx,
                      dgamma, dbeta);
  else      KP_LAUNCH(kp::bn_bwd_kernel<false>, C / 4, 256, 0, st, x, dy, N, C, gamma, beta, save_mean, save_invstd, dx,
                      dgamma, dbeta);
  return 0;
}

}  // extern "C"
