// agg_fast_host.h -- host entry points of the fast aggregation kernels; each is instantiated in its own
// translation unit (agg_fast_fwd.cu / agg_fast_b1.cu / agg_fast_b2.cu) so the library builds in parallel.
#pragma once
#include "agg_fast.cuh"

namespace kp {

// Template combinations that exist (anything else takes the generic kernels):
//   (GELU, fuse or not, no extras)  KPGINPlus
//   (RELU, fuse or not, extras)     KPGCN (norm)
//   (NONE, unfused, no extras)      KPGraphSAGE add
//   (NONE, unfused, extras)         KPGIN / GINE / KGIN (self term), KPGraphSAGE mean
inline bool fast_combo(int act, bool fuse, bool need_extra, bool* extra) {
  if (act == KP_ACT_GELU) {
    *extra = false;
    return !need_extra;
  }
  if (act == KP_ACT_RELU) {
    *extra = true;
    return true;
  }
  *extra = need_extra;
  return !fuse;
}

// Test hook (kp_agg_set_launch_geometry): g_geom_max_ctas > 0 caps the grid of every persistent aggregation kernel so
// that each lane group / CTA loops over many nodes even on a small batch; g_geom_lean_threads > 0 forces the CTA size
// of the lean kernels (256..1024).  Both 0 in production.
extern int g_geom_max_ctas, g_geom_lean_threads;
inline int geom_cap(long long grid) {
  if (g_geom_max_ctas > 0 && grid > g_geom_max_ctas) grid = g_geom_max_ctas;
  return (int)(grid < 1 ? 1 : grid);
}

void fast_fwd_set_lean(int flag);   // 1 (default): packed-math + L2-prefetch kernels (agg_lean.cuh) where k + 1 <= G
void fast_fwd_set_ring(int flag);   // 0: plain register-prefetch forward kernel, 1 (default): cp.async ring
int fast_fwd(const FastArgs& fa, int G, int act, bool fuse, int tab, bool extra, int grid, size_t smem, float* out,
             cudaStream_t st);
bool lean_b2(const FastArgs& fa, int G, const float* Gs, float* dX, cudaStream_t st, int* rc);
bool fast_lean_enabled();
void fast_fwd_set_tma(int mode);   // 0 = never, 1 = large batches (default), 2 = always
bool tma_fwd(const FastArgs& fa, int G, int act, bool fuse, int tab, bool extra, size_t staged_bytes, float* out,
             cudaStream_t st, int* rc);
int lean_b1(const FastArgs& fa, int G, int act, bool fuse, int tab, int grid, int threads, size_t smem,
            const float* dOut, float* Gs, float* dP, float* dth, cudaStream_t st);
int fast_b1(const FastArgs& fa, int G, int act, bool fuse, int tab, bool extra, int grid, size_t smem,
            const float* dOut, float* Gs, float* dP, float* dth, float* dep, cudaStream_t st);
int fast_b2(const FastArgs& fa, int G, bool fuse, bool extra, int grid, const float* Gs, const float* dOut, float* dX,
            cudaStream_t st);

// block-resident kernels for long rows (agg_tile.cu)
void tile_set_mode(int mode);      // 0 = never, 1 = when the caller supplies closed blocks (default)
bool tile_eligible(const kp_agg_desc& a, int tab);
int tile_fwd(const FastArgs& fa, int G, int act, int tab, bool extra, float* out, cudaStream_t st);
int tile_b2(const FastArgs& fa, int G, bool extra, const float* Gs, const float* dOut, float* dX, cudaStream_t st);

// the whole backward as one block-resident kernel (agg_block_bwd.cu)
void block_bwd_set_mode(int mode);   // 0 = never, 1 = when the caller supplies closed blocks (default)
bool block_bwd_eligible(const kp_agg_desc& a, int G, int tab, bool want_dtheta);
int block_bwd_grid();
int block_bwd(const FastArgs& fa, const float* dOut, float* dX, float* dP, float* dth_part, float* tab_part, int* grid_out,
              cudaStream_t st);

// dispatch helper shared by the three translation units
#define KP_FAST_TAB(FN, G, A, F, X, tab, ...)                                  \
  ((tab) == TAB_SMEM ? FN<G, A, F, TAB_SMEM, X>(__VA_ARGS__)                   \
   : (tab) == TAB_GLOBAL ? FN<G, A, F, TAB_GLOBAL, X>(__VA_ARGS__)             \
                         : FN<G, A, F, TAB_NONE, X>(__VA_ARGS__))
#define KP_FAST_COMBO(FN, G, act, fuse, extra, tab, ...)                                                          \
  ((act) == KP_ACT_GELU ? ((fuse) ? KP_FAST_TAB(FN, G, KP_ACT_GELU, true, false, tab, __VA_ARGS__)                \
                                  : KP_FAST_TAB(FN, G, KP_ACT_GELU, false, false, tab, __VA_ARGS__))              \
   : (act) == KP_ACT_RELU ? ((fuse) ? KP_FAST_TAB(FN, G, KP_ACT_RELU, true, true, tab, __VA_ARGS__)               \
                                    : KP_FAST_TAB(FN, G, KP_ACT_RELU, false, true, tab, __VA_ARGS__))             \
   : (extra) ? KP_FAST_TAB(FN, G, KP_ACT_NONE, false, true, tab, __VA_ARGS__)                                     \
             : KP_FAST_TAB(FN, G, KP_ACT_NONE, false, false, tab, __VA_ARGS__))
#define KP_FAST_G(FN, g, act, fuse, extra, tab, ...)                                      \
  ((g) == 32 ? KP_FAST_COMBO(FN, 32, act, fuse, extra, tab, __VA_ARGS__)                  \
   : (g) == 16 ? KP_FAST_COMBO(FN, 16, act, fuse, extra, tab, __VA_ARGS__)                \
   : (g) == 8 ? KP_FAST_COMBO(FN, 8, act, fuse, extra, tab, __VA_ARGS__)                  \
              : KP_FAST_COMBO(FN, 4, act, fuse, extra, tab, __VA_ARGS__))

}  // namespace kp
