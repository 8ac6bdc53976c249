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

template <int G, int ACT, bool FUSE, int TAB, bool EXTRA>
static int launch_ring(const FastArgs& fa, int grid, size_t smem, float* out, cudaStream_t st) {
  // G is fixed at 32 for the ring kernel; the template parameter only keeps the dispatch macro uniform
  const unsigned slot = fa.d.d <= 104 ? KP_RING_SLOT_BYTES : 512u;
  const int stage_floats = (int)(smem / sizeof(float));
  const size_t total = smem + (size_t)8 * 9 * slot;
  if (total > 48 * 1024)
    KP_CUDA(cudaFuncSetAttribute(agg_fwd_ring_kernel<ACT, FUSE, TAB, EXTRA>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));
  KP_LAUNCH((agg_fwd_ring_kernel<ACT, FUSE, TAB, EXTRA>), grid, 256, total, st, fa, out, stage_floats, slot);
  return 0;
}

static int g_use_ring = 1;
void fast_fwd_set_ring(int flag) { g_use_ring = flag; }

int fast_fwd(const FastArgs& fa, int G, int act, bool fuse, int tab, bool extra, int grid, size_t smem, float* out,
             cudaStream_t st) {
  if (g_use_ring && G == 32 && fa.d.k <= 31 && smem + 8 * 9 * 512 <= 100 * 1024)
    return KP_FAST_COMBO(launch_ring, 32, act, fuse, extra, tab, fa, grid, smem, out, st);
  return KP_FAST_G(launch, G, act, fuse, extra, tab, fa, grid, smem, out, st);
}

}  // namespace kp
