// agg_fast_fwd.cu -- instantiations of the fast forward aggregation kernel (see agg_fast.cuh).
#include "agg_fast_host.h"

namespace kp {

template <int G, int ACT, bool FUSE, int TAB, bool EXTRA>
static int launch(const FastArgs& fa, int grid, size_t smem, float* out, cudaStream_t st) {
  if (smem > 48 * 1024)
    KP_CUDA(cudaFuncSetAttribute(agg_fwd_fast_kernel<G, ACT, FUSE, TAB, EXTRA>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KP_LAUNCH((agg_fwd_fast_kernel<G, ACT, FUSE, TAB, EXTRA>), grid, 256, smem, st, fa, out);
  return 0;
}

int fast_fwd(const FastArgs& fa, int G, int act, bool fuse, int tab, bool extra, int grid, size_t smem, float* out,
             cudaStream_t st) {
  return KP_FAST_G(launch, G, act, fuse, extra, tab, fa, grid, smem, out, st);
}

}  // namespace kp
