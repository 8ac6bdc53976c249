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
adam_kernel(const kp_adam_tensor* __restrict__ tensors, const int2* __restrict__ chunks, float lr, float b1, float b2,
            float omb1, float omb2, float eps, int* __restrict__ state /* [0] = steps taken, [1] = finished blocks */) {
  const int2 ck = chunks[blockIdx.x];
  const kp_adam_tensor t = tensors[ck.x];
  const int step = state[0] + 1;
  const float bc1 = 1.f - powf(b1, (float)step);
  const float bc2s = sqrtf(1.f - powf(b2, (float)step));
  const float step_size = lr / bc1;
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
  KP_LAUNCH(kp::adam_kernel, nchunks, 256, 0, (cudaStream_t)stream, tensors_dev, (const int2*)chunks_dev, lr, (float)beta1,
            (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), eps, state_dev);
  return 0;
}
