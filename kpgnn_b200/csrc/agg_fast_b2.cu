// agg_fast_b2.cu -- instantiations of the fast backward-by-source kernel (see agg_fast.cuh).
#include "agg_fast_host.h"

namespace kp {

template <int G>
static int launch(const FastArgs& fa, bool fuse, bool extra, int grid, const float* Gs, const float* dOut, float* dX,
                  cudaStream_t st) {
  if (fuse && extra)       KP_LAUNCH((agg_bwd_src_fast_kernel<G, true, true>), grid, 256, 0, st, fa, Gs, dOut, dX);
  else if (fuse)           KP_LAUNCH((agg_bwd_src_fast_kernel<G, true, false>), grid, 256, 0, st, fa, Gs, dOut, dX);
  else if (extra)          KP_LAUNCH((agg_bwd_src_fast_kernel<G, false, true>), grid, 256, 0, st, fa, Gs, dOut, dX);
  else                     KP_LAUNCH((agg_bwd_src_fast_kernel<G, false, false>), grid, 256, 0, st, fa, Gs, dOut, dX);
  return 0;
}

int fast_b2(const FastArgs& fa, int G, bool fuse, bool extra, int grid, const float* Gs, const float* dOut, float* dX,
            cudaStream_t st) {
  int rc = 0;
  if (!extra && lean_b2(fa, G, Gs, dX, st, &rc)) return rc;
  if (fa.d.dx_node_stride || fa.d.dx_hop_stride || fa.d.dx_accumulate) {
    set_error("kp_agg_backward: strided / accumulated dX is only available on the lean gather kernel");
    return 3;
  }
  return G == 32 ? launch<32>(fa, fuse, extra, grid, Gs, dOut, dX, st)
         : G == 16 ? launch<16>(fa, fuse, extra, grid, Gs, dOut, dX, st)
         : G == 8 ? launch<8>(fa, fuse, extra, grid, Gs, dOut, dX, st)
                  : launch<4>(fa, fuse, extra, grid, Gs, dOut, dX, st);
}

}  // namespace kp
