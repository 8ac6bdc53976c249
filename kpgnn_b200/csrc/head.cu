// head.cu -- the graph-regression head of the ZINC training step as one kernel each way:
//     h = relu(rep)            (the ReLU of the backbone's output projection, models/GNNs.py:276-277, dropout p = 0)
//     pooled[g] = sum (mean) of h over the nodes of graph g                 (models/GraphRegression.py:26)
//     score[g]  = <w, pooled[g]> + b                                         (GraphRegression.py regressor)
//     loss      = mean_g |score[g] - y[g]|   (L1, train_ZINC.py:42)   or   mean_g (score[g] - y[g])^2
// As PyTorch ops this was ~18 kernels of 1-7 us between the output projection and its backward (clamp, pooling, gemv,
// add, abs, mean; fill, sign, mul, two gemv, reduce, cat, gather, mask).  One CTA per graph; every sum runs in a fixed
// order (rows ascending per column thread, a shared-memory tree for the dot product, the last CTA to arrive adds the
// per-graph terms in graph order): bit-reproducible, no float atomics.
#include "common.cuh"

namespace kp {

constexpr int HEAD_THREADS = 128;

__device__ __forceinline__ int head_lower_bound(const int64_t* __restrict__ seg, int N, int64_t key) {
  int lo = 0, hi = N;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(seg + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// fixed-order block sum through shared memory (HEAD_THREADS a power of two); every thread gets the result
__device__ __forceinline__ float head_block_sum(float v, float* red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int w = HEAD_THREADS / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  const float s = red[0];
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(HEAD_THREADS)
head_fwd_kernel(const kp_head_desc h, float* __restrict__ pooled, float* __restrict__ score, float* __restrict__ loss,
                float* __restrict__ terms, unsigned* __restrict__ counter) {
  __shared__ float red[HEAD_THREADS];
  __shared__ int s_lo, s_hi;
  __shared__ bool last;
  const int g = blockIdx.x;
  const int N = h.n_dev ? min(__ldg(h.n_dev), h.N) : h.N;
  if (threadIdx.x == 0) s_lo = head_lower_bound(h.batch, N, g);
  if (threadIdx.x == 32) s_hi = head_lower_bound(h.batch, N, (int64_t)g + 1);
  __syncthreads();
  const int lo = s_lo, hi = s_hi;
  const float inv = (h.mean && hi > lo) ? 1.f / (float)(hi - lo) : 1.f;
  float dot = 0.f;
  for (int c = threadIdx.x; c < h.H; c += HEAD_THREADS) {
    const float* p = h.rep + (size_t)lo * h.rep_stride + c;
    float s = 0.f;
    int i = lo;
    for (; i + 4 <= hi; i += 4) {            // loads batched, adds in row order
      const float a0 = __ldg(p), a1 = __ldg(p + h.rep_stride), a2 = __ldg(p + 2 * h.rep_stride),
                  a3 = __ldg(p + 3 * h.rep_stride);
      s += fmaxf(a0, 0.f); s += fmaxf(a1, 0.f); s += fmaxf(a2, 0.f); s += fmaxf(a3, 0.f);
      p += 4 * h.rep_stride;
    }
    for (; i < hi; ++i, p += h.rep_stride) s += fmaxf(__ldg(p), 0.f);
    s *= inv;
    pooled[(size_t)g * h.H + c] = s;
    dot = fmaf(__ldg(h.w + c), s, dot);
  }
  const float sc = head_block_sum(dot, red) + __ldg(h.b);
  if (threadIdx.x == 0) {
    score[g] = sc;
    const float e = sc - __ldg(h.y + g);
    __stcg(terms + g, h.loss_kind == 0 ? fabsf(e) : e * e);
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {                                 // per-graph terms in graph order
    __threadfence();
    float s = 0.f;
    for (int q = threadIdx.x; q < h.G; q += HEAD_THREADS) s += __ldcg(terms + q);   // strided partials, then the tree:
    const float tot = head_block_sum(s, red);                                       // fixed for a given G
    if (threadIdx.x == 0) {
      loss[0] = tot / (float)h.G;
      *counter = 0u;                          // self-resetting
    }
  }
}

// d(rep)[i,c] = dscore[g(i)] * w[c] * (rep[i,c] > 0) (/ count for mean pooling); dw, db by the last CTA in graph order
__global__ void __launch_bounds__(HEAD_THREADS)
head_bwd_kernel(const kp_head_desc h, const float* __restrict__ pooled, const float* __restrict__ score,
                const float* __restrict__ dloss, float* __restrict__ drep, float* __restrict__ dw, float* __restrict__ db,
                float* __restrict__ dscore, unsigned* __restrict__ counter) {
  __shared__ int s_lo, s_hi;
  __shared__ bool last;
  const int g = blockIdx.x;
  const int N = h.n_dev ? min(__ldg(h.n_dev), h.N) : h.N;
  if (g == h.G) {                             // extra CTA: rows behind the last graph (padding of a capacity batch) get zeros
    for (long long t = (long long)N * h.H + threadIdx.x; t < (long long)h.N * h.H; t += HEAD_THREADS)
      drep[(t / h.H) * h.rep_stride_out + (t % h.H)] = 0.f;
  } else {
    if (threadIdx.x == 0) s_lo = head_lower_bound(h.batch, N, g);
    if (threadIdx.x == 32) s_hi = head_lower_bound(h.batch, N, (int64_t)g + 1);
    __syncthreads();
    const int lo = s_lo, hi = s_hi;
    const float e = __ldg(score + g) - __ldg(h.y + g);
    const float de = h.loss_kind == 0 ? (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) : 2.f * e;
    const float ds = __ldg(dloss) * de / (float)h.G;
    if (threadIdx.x == 0) __stcg(dscore + g, ds);
    const float inv = (h.mean && hi > lo) ? 1.f / (float)(hi - lo) : 1.f;
    for (int c = threadIdx.x; c < h.H; c += HEAD_THREADS) {
      const float k = ds * __ldg(h.w + c) * inv;
      for (int i = lo; i < hi; ++i) {
        const float r = __ldg(h.rep + (size_t)i * h.rep_stride + c);
        drep[(size_t)i * h.rep_stride_out + c] = r > 0.f ? k : 0.f;
      }
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    for (int c = threadIdx.x; c < h.H; c += HEAD_THREADS) {
      float s = 0.f;
      for (int q = 0; q < h.G; ++q) s = fmaf(__ldcg(dscore + q), __ldg(pooled + (size_t)q * h.H + c), s);
      dw[c] = s;
    }
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int q = 0; q < h.G; ++q) s += __ldcg(dscore + q);
      db[0] = s;
      *counter = 0u;
    }
  }
}

}  // namespace kp

extern "C" {

size_t kp_head_workspace_bytes(int32_t G) { return 256 + 2 * sizeof(float) * (size_t)(G < 1 ? 1 : G); }

static int head_check(const kp_head_desc& h) {
  KP_CHECK_ARG(h.N >= 0 && h.H >= 1 && h.G >= 1 && h.rep && h.batch && h.w && h.b && h.y && h.rep_stride >= h.H &&
                   (h.loss_kind == 0 || h.loss_kind == 1),
               "kp_head: bad descriptor");
  return 0;
}

int kp_head_forward(const kp_head_desc* desc, float* pooled, float* score, float* loss, void* workspace,
                    size_t workspace_bytes, void* stream) {
  KP_CHECK_ARG(desc && pooled && score && loss && workspace, "kp_head_forward: null argument");
  if (head_check(*desc)) return 1;
  KP_CHECK_ARG(workspace_bytes >= kp_head_workspace_bytes(desc->G), "kp_head_forward: workspace too small");
  unsigned* counter = (unsigned*)workspace;
  float* terms = (float*)((char*)workspace + 256);
  KP_LAUNCH(kp::head_fwd_kernel, desc->G, kp::HEAD_THREADS, 0, stream, *desc, pooled, score, loss, terms, counter);
  return 0;
}

int kp_head_backward(const kp_head_desc* desc, const float* pooled, const float* score, const float* dloss, float* drep,
                     float* dw, float* db, void* workspace, size_t workspace_bytes, void* stream) {
  KP_CHECK_ARG(desc && pooled && score && dloss && drep && dw && db && workspace, "kp_head_backward: null argument");
  if (head_check(*desc)) return 1;
  KP_CHECK_ARG(desc->rep_stride_out >= desc->H, "kp_head_backward: rep_stride_out < H");
  KP_CHECK_ARG(workspace_bytes >= kp_head_workspace_bytes(desc->G), "kp_head_backward: workspace too small");
  unsigned* counter = (unsigned*)((char*)workspace + 128);
  float* dscore = (float*)((char*)workspace + 256) + desc->G;
  KP_LAUNCH(kp::head_bwd_kernel, desc->G + 1, kp::HEAD_THREADS, 0, stream, *desc, pooled, score, dloss, drep, dw, db,
            dscore, counter);
  return 0;
}

}  // extern "C"
