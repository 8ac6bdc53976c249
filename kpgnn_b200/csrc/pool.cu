// pool.cu -- graph readout over a SORTED segment vector (PyG global_add_pool / global_mean_pool on `data.batch`,
// models/GraphRegression.py:26, GraphClassification.py): out[g,:] = sum of the rows i with seg[i] == g, added in
// ascending row order by one thread per column -- bit-reproducible, no float atomics (torch's index_add_ is a float
// atomicAdd).  One CTA per graph finds its row range with two binary searches over `seg`.
#include "common.cuh"

namespace kp {

__device__ __forceinline__ int seg_lower_bound(const int64_t* __restrict__ seg, int N, int64_t key) {
  int lo = 0, hi = N;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(seg + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(128)
segment_sum_kernel(const float* __restrict__ x, long long x_stride, const int64_t* __restrict__ seg, int N, int C,
                   int mean, float* __restrict__ out) {
  __shared__ int s_lo, s_hi;
  const int g = blockIdx.x;
  if (threadIdx.x == 0) s_lo = seg_lower_bound(seg, N, g);
  if (threadIdx.x == 32) s_hi = seg_lower_bound(seg, N, (int64_t)g + 1);
  __syncthreads();
  const int lo = s_lo, hi = s_hi;
  const float scale = (mean && hi > lo) ? 1.f / (float)(hi - lo) : 1.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* p = x + (size_t)lo * x_stride + c;
    float s = 0.f;
    int i = lo;
    for (; i + 4 <= hi; i += 4) {            // loads batched, adds in row order
      const float a0 = __ldg(p), a1 = __ldg(p + x_stride), a2 = __ldg(p + 2 * x_stride), a3 = __ldg(p + 3 * x_stride);
      s += a0; s += a1; s += a2; s += a3;
      p += 4 * x_stride;
    }
    for (; i < hi; ++i, p += x_stride) s += __ldg(p);
    out[(size_t)g * C + c] = s * scale;
  }
}

}  // namespace kp

extern "C" int kp_segment_sum(const float* x, int64_t x_stride, const int64_t* seg, int32_t N, int32_t C, int32_t G,
                              int32_t mean, float* out, void* stream) {
  KP_CHECK_ARG(out && (G == 0 || (x || N == 0)) && (seg || N == 0) && N >= 0 && C >= 1 && G >= 0 && x_stride >= C,
               "kp_segment_sum: bad argument");
  if (G == 0) return 0;
  KP_LAUNCH(kp::segment_sum_kernel, G, 128, 0, stream, x, (long long)x_stride, seg, N, C, mean, out);
  return 0;
}
