// adam.cu -- Adam (torch.optim.Adam semantics: no amsgrad, no weight decay) over a LIST of parameter tensors in one
// launch.  The training step of the headline workload has ~100 parameter tensors with 0.5 M elements in total;
// torch's fused multi-tensor Adam handles them in 4 launches of ~15 us each (one 512-thread block per tensor chunk,
// profiles/r1z_step_kineto.txt).  Here a device-side table lists every 1024-element chunk of every tensor, so one
// launch of ~500 blocks covers the whole model (~3 us).  The step counter lives on the device (CUDA-graph replays
// cannot change kernel arguments): every block reads it, the block that finishes last advances it.
//     m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// (train_ZINC.py:244 uses torch.optim.Adam(lr) with its defaults.)
#include "common.cuh"

namespace kp {

constexpr int ADAM_CHUNK = 1024;

__global__ void __launch_bounds__(256)
adam_kernel(const kp_adam_tensor* __restrict__ tensors, const int2* __restrict__ chunks, double lr, double b1d,
            double b2d, float eps, int* __restrict__ state /* [0] = steps taken, [1] = finished blocks */) {
  const int2 ck = chunks[blockIdx.x];
  const kp_adam_tensor t = tensors[ck.x];
  const int step = state[0] + 1;
  // Bias corrections in double, as torch.optim.Adam forms them on the host (1 - beta^t cancels catastrophically in
  // fp32 at small t: 1 - 0.999f is off by 1.3e-5 relative).  One thread per block, broadcast through shared memory.
  __shared__ float s_step_size, s_bc2s;
  if (threadIdx.x == 0) {
    const double bc1 = -expm1((double)step * log(b1d));
    const double bc2 = -expm1((double)step * log(b2d));
    s_step_size = (float)(lr / bc1);
    s_bc2s = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_step_size, bc2s = s_bc2s;
  const float b1 = (float)b1d, b2 = (float)b2d, omb1 = (float)(1.0 - b1d), omb2 = (float)(1.0 - b2d);
  const int end = min(t.n, ck.y + ADAM_CHUNK);
  for (int i = ck.y + threadIdx.x; i < end; i += blockDim.x) {
    const float g = t.g[i];
    const float m = fmaf(b1, t.m[i], omb1 * g);
    const float v = fmaf(b2, t.v[i], omb2 * g * g);
    t.m[i] = m;
    t.v[i] = v;
    t.p[i] -= step_size * (m / (sqrtf(v) / bc2s + eps));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(state + 1, 1) == (int)gridDim.x - 1) {      // every block has read state[0] by now
      state[0] = step;
      state[1] = 0;
    }
  }
}

}  // namespace kp

extern "C" int kp_adam_step(const kp_adam_tensor* tensors_dev, const int32_t* chunks_dev, int32_t nchunks, float lr,
                            double beta1, double beta2, float eps, int32_t* state_dev, void* stream) {
  KP_CHECK_ARG(tensors_dev && chunks_dev && state_dev && nchunks >= 0, "kp_adam_step: null argument");
  KP_CHECK_ARG((((uintptr_t)chunks_dev) & 7) == 0 && (((uintptr_t)tensors_dev) & 7) == 0, "kp_adam_step: misaligned table");
  if (nchunks == 0) return 0;
  KP_LAUNCH(kp::adam_kernel, nchunks, 256, 0, (cudaStream_t)stream, tensors_dev, (const int2*)chunks_dev, (double)lr, beta1,
            beta2, eps, state_dev);
  return 0;
}
