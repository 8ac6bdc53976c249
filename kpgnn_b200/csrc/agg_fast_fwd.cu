// agg_fast_fwd.cu -- instantiations of the fast forward aggregation kernel (see agg_fast.cuh).
#include "agg_fast_host.h"
#include "agg_lean.cuh"

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

template <int G, int ACT, bool FUSE, int TAB, bool EXTRA>
static int launch_lean(const FastArgs& fa, int grid, size_t smem, float* out, cudaStream_t st) {
  const kp_agg_desc& a = fa.d;
  // L2 prefetch: lane l of a group asks for the l-th 128-byte line of a node's k hop rows when they are contiguous
  const unsigned lines = (unsigned)((a.k * a.d * 4 + 127) / 128);
  const unsigned pfx = (fa.xh == (unsigned)a.d) ? (lines < (unsigned)G ? lines : (unsigned)G) : 0u;
  const unsigned pfp = (a.P && fa.ph == (unsigned)a.d) ? (lines < (unsigned)G ? lines : (unsigned)G) : 0u;
  const int dist = a.k >= 4 ? 1 : (a.k >= 2 ? 2 : 4);
  const size_t total = smem + (size_t)(256 / G) * 12 * G;       // + per-group entry window and row pointers
  if (total > 48 * 1024)
    KP_CUDA(cudaFuncSetAttribute(agg_fwd_lean_kernel<G, ACT, FUSE, TAB, EXTRA>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));
  KP_LAUNCH((agg_fwd_lean_kernel<G, ACT, FUSE, TAB, EXTRA>), grid, 256, total, st, fa, out, pfx, pfp, dist);
  return 0;
}

// lean kernels exist for G = 32 / 16 and tables in shared memory (or none)
#define KP_LEAN_TAB(G, A, F, X, tab, ...) \
  ((tab) == TAB_SMEM ? launch_lean<G, A, F, TAB_SMEM, X>(__VA_ARGS__) : launch_lean<G, A, F, TAB_NONE, X>(__VA_ARGS__))
#define KP_LEAN_COMBO(G, act, fuse, extra, tab, ...)                                                   \
  ((act) == KP_ACT_GELU ? ((fuse) ? KP_LEAN_TAB(G, KP_ACT_GELU, true, false, tab, __VA_ARGS__)         \
                                  : KP_LEAN_TAB(G, KP_ACT_GELU, false, false, tab, __VA_ARGS__))       \
   : (act) == KP_ACT_RELU ? ((fuse) ? KP_LEAN_TAB(G, KP_ACT_RELU, true, true, tab, __VA_ARGS__)        \
                                    : KP_LEAN_TAB(G, KP_ACT_RELU, false, true, tab, __VA_ARGS__))      \
   : (extra) ? KP_LEAN_TAB(G, KP_ACT_NONE, false, true, tab, __VA_ARGS__)                              \
             : KP_LEAN_TAB(G, KP_ACT_NONE, false, false, tab, __VA_ARGS__))

static int g_use_ring = 1;
static int g_use_lean = 1;
void fast_fwd_set_ring(int flag) { g_use_ring = flag; }
void fast_fwd_set_lean(int flag) { g_use_lean = flag; }

int fast_fwd(const FastArgs& fa, int G, int act, bool fuse, int tab, bool extra, int grid, size_t smem, float* out,
             cudaStream_t st) {
  if (g_use_lean && (G == 32 || G == 16) && fa.d.k + 1 <= G && !fa.d.dinv && tab != TAB_GLOBAL) {
    if (G == 32) return KP_LEAN_COMBO(32, act, fuse, extra, tab, fa, grid, smem, out, st);
    return KP_LEAN_COMBO(16, act, fuse, extra, tab, fa, grid, smem, out, st);
  }
  if (g_use_ring && G == 32 && fa.d.k <= 31 && smem + 8 * 9 * 512 <= 100 * 1024)
    return KP_FAST_COMBO(launch_ring, 32, act, fuse, extra, tab, fa, grid, smem, out, st);
  return KP_FAST_G(launch, G, act, fuse, extra, tab, fa, grid, smem, out, st);
}

}  // namespace kp
