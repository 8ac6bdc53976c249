// agg_fast_b1.cu -- instantiations of the fast backward-by-destination kernel (see agg_fast.cuh).
#include "agg_fast_host.h"

namespace kp {

template <int G, int ACT, bool FUSE, int TAB, bool EXTRA>
static int launch(const FastArgs& fa, int grid, size_t smem, const float* dOut, float* Gs, float* dP, float* dth,
                  float* dep, cudaStream_t st) {
  if (smem > 48 * 1024)
    KP_CUDA(cudaFuncSetAttribute(agg_bwd_dst_fast_kernel<G, ACT, FUSE, TAB, EXTRA>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KP_LAUNCH((agg_bwd_dst_fast_kernel<G, ACT, FUSE, TAB, EXTRA>), grid, 256, smem, st, fa, dOut, Gs, dP, dth, dep);
  return 0;
}

int fast_b1(const FastArgs& fa, int G, int act, bool fuse, int tab, bool extra, int grid, size_t smem,
            const float* dOut, float* Gs, float* dP, float* dth, float* dep, cudaStream_t st) {
  return KP_FAST_G(launch, G, act, fuse, extra, tab, fa, grid, smem, dOut, Gs, dP, dth, dep, st);
}

}  // namespace kp
