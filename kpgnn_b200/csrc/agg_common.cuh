// agg_common.cuh -- vector load/store helpers and activation math shared by the aggregation kernels.
#pragma once
#include "common.cuh"

namespace kp {

template <int VEC>
struct Vf {
  float v[VEC];
};

template <int VEC>
__device__ __forceinline__ Vf<VEC> vload(const float* __restrict__ p);
template <>
__device__ __forceinline__ Vf<4> vload<4>(const float* __restrict__ p) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  return Vf<4>{{t.x, t.y, t.z, t.w}};
}
template <>
__device__ __forceinline__ Vf<2> vload<2>(const float* __restrict__ p) {
  float2 t = __ldg(reinterpret_cast<const float2*>(p));
  return Vf<2>{{t.x, t.y}};
}
template <>
__device__ __forceinline__ Vf<1> vload<1>(const float* __restrict__ p) {
  return Vf<1>{{__ldg(p)}};
}
// streaming variants for data touched exactly once (P, dOut, outputs): keep L1/L2 for the gathered rows
template <int VEC>
__device__ __forceinline__ Vf<VEC> vload_stream(const float* __restrict__ p);
template <>
__device__ __forceinline__ Vf<4> vload_stream<4>(const float* __restrict__ p) {
  float4 t = __ldcs(reinterpret_cast<const float4*>(p));
  return Vf<4>{{t.x, t.y, t.z, t.w}};
}
template <>
__device__ __forceinline__ Vf<2> vload_stream<2>(const float* __restrict__ p) {
  float2 t = __ldcs(reinterpret_cast<const float2*>(p));
  return Vf<2>{{t.x, t.y}};
}
template <>
__device__ __forceinline__ Vf<1> vload_stream<1>(const float* __restrict__ p) {
  return Vf<1>{{__ldcs(p)}};
}
template <int VEC>
__device__ __forceinline__ void vstore(float* __restrict__ p, const Vf<VEC>& x);
template <>
__device__ __forceinline__ void vstore<4>(float* __restrict__ p, const Vf<4>& x) {
  *reinterpret_cast<float4*>(p) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
}
template <>
__device__ __forceinline__ void vstore<2>(float* __restrict__ p, const Vf<2>& x) {
  *reinterpret_cast<float2*>(p) = make_float2(x.v[0], x.v[1]);
}
template <>
__device__ __forceinline__ void vstore<1>(float* __restrict__ p, const Vf<1>& x) {
  *p = x.v[0];
}

// exact-erf GELU (KPGINplus.py:87-88, F.gelu default) evaluated through the Abramowitz-Stegun 7.1.26 erfc form:
//   0.5*erfc(|x|/sqrt2) = t*(a1/2 + t*(a2/2 + ...)) * exp(-x^2/2),  t = 1/(1 + p*|x|/sqrt2)
// |error| < 5e-7 absolute over all x in fp32 (torch's own fp32 gelu is 1.2e-6 from the exact value), no
// cancellation in the negative tail.  Raw MUFU.RCP / MUFU.EX2 (2 ulp) instead of the ~35-instruction erff.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void gelu_parts(float x, float& half_erfc, float& gauss) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f));
  float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  p = fmaf(t, p, 0.5f * 1.421413741f);
  p = fmaf(t, p, 0.5f * -0.284496736f);
  p = fmaf(t, p, 0.5f * 0.254829592f);
  gauss = ex2_approx((x * x) * (-0.5f * 1.44269504088896340736f));   // exp(-x^2/2)
  half_erfc = (p * t) * gauss;
}
template <int ACT>
__device__ __forceinline__ float act_fwd(float x) {
  if (ACT == KP_ACT_GELU) {
    float hc, g;
    gelu_parts(x, hc, g);
    return fmaf(-fabsf(x), hc, fmaxf(x, 0.f));      // x>0: x - x*hc ; x<=0: x*hc
  }
  if (ACT == KP_ACT_RELU) return x > 0.f ? x : 0.f;
  return x;
}
template <int ACT>
__device__ __forceinline__ float act_bwd(float x) {
  if (ACT == KP_ACT_GELU) {
    float hc, g;
    gelu_parts(x, hc, g);
    const float cdf = x > 0.f ? 1.0f - hc : hc;
    return fmaf(x * 0.39894228040143267794f, g, cdf);
  }
  if (ACT == KP_ACT_RELU) return x > 0.f ? 1.f : 0.f;
  return 1.f;
}

}  // namespace kp
