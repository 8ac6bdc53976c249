// agg_fast_b1.cu -- instantiations of the fast backward-by-destination kernel (see agg_fast.cuh).
#include "agg_fast_host.h"
#include "agg_lean.cuh"

namespace kp {

template <int G, int ACT, bool FUSE, int TAB, bool EXTRA>
static int launch(const FastArgs& fa, int grid, size_t smem, const float* dOut, float* Gs, float* dP, float* dth,
                  float* dep, cudaStream_t st) {
  if (smem > 32 * 1024)
    KP_CUDA(cudaFuncSetAttribute(agg_bwd_dst_fast_kernel<G, ACT, FUSE, TAB, EXTRA>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KP_LAUNCH((agg_bwd_dst_fast_kernel<G, ACT, FUSE, TAB, EXTRA>), grid, 256, smem, st, fa, dOut, Gs, dP, dth, dep);
  return 0;
}

template <int G, int ACT, bool FUSE, int TAB>
static int launch_lean(const FastArgs& fa, int grid, int threads, size_t smem, const float* dOut, float* Gs, float* dP,
                       float* dth, cudaStream_t st) {
  if (smem > 32 * 1024)
    KP_CUDA(cudaFuncSetAttribute(agg_bwd_dst_lean_kernel<G, ACT, FUSE, TAB>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KP_LAUNCH_PDL((agg_bwd_dst_lean_kernel<G, ACT, FUSE, TAB>), grid, threads, smem, st, fa, dOut, Gs, dP, dth);
  return 0;
}
template <int G>
static int lean_g(const FastArgs& fa, int act, bool fuse, int tab, int grid, int threads, size_t smem,
                  const float* dOut, float* Gs, float* dP, float* dth, cudaStream_t st) {
#define KP_LB1(A, F)                                                                                       \
  ((tab) == TAB_SMEM ? launch_lean<G, A, F, TAB_SMEM>(fa, grid, threads, smem, dOut, Gs, dP, dth, st)      \
                     : launch_lean<G, A, F, TAB_NONE>(fa, grid, threads, smem, dOut, Gs, dP, dth, st))
  if (act == KP_ACT_GELU) return fuse ? KP_LB1(KP_ACT_GELU, true) : KP_LB1(KP_ACT_GELU, false);
  if (act == KP_ACT_RELU) return fuse ? KP_LB1(KP_ACT_RELU, true) : KP_LB1(KP_ACT_RELU, false);
  return KP_LB1(KP_ACT_NONE, true);      // act none is only routed here when fused (z needed for dtheta)
#undef KP_LB1
}
// lean B1: G = 32 / 16, no extras, tables in shared memory or none (eligibility decided in agg.cu make_config)
int lean_b1(const FastArgs& fa, int G, int act, bool fuse, int tab, int grid, int threads, size_t smem,
            const float* dOut, float* Gs, float* dP, float* dth, cudaStream_t st) {
  return G == 32 ? lean_g<32>(fa, act, fuse, tab, grid, threads, smem, dOut, Gs, dP, dth, st)
                 : lean_g<16>(fa, act, fuse, tab, grid, threads, smem, dOut, Gs, dP, dth, st);
}

int fast_b1(const FastArgs& fa, int G, int act, bool fuse, int tab, bool extra, int grid, size_t smem,
            const float* dOut, float* Gs, float* dP, float* dth, float* dep, cudaStream_t st) {
  return KP_FAST_G(launch, G, act, fuse, extra, tab, fa, grid, smem, dOut, Gs, dP, dth, dep, st);
}

}  // namespace kp
