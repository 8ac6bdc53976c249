// agg_tma_fwd.cu -- instantiations and launch logic of the TMA-staged forward kernel (see agg_tma.cuh).
#include "agg_fast_host.h"
#include "agg_tma.cuh"

namespace kp {

static int g_tma_mode = 0;      // 0 = never (default: measured 344 us vs 333 us for the lean kernel), 1 = large batches, 2 = always (tests)
void fast_fwd_set_tma(int mode) { g_tma_mode = mode; }

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// a [N][k][d] fp32 tensor (element strides node_stride, hop_stride, 1), box = {d, 1, box_rows}, no swizzle
static bool make_map(const kp_agg_desc& a, const float* base, unsigned node_stride, unsigned hop_stride, int box_rows,
                     CUtensorMap* tm) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)a.d, (cuuint64_t)a.k, (cuuint64_t)a.N};
  const cuuint64_t strides[2] = {(cuuint64_t)hop_stride * 4ull, (cuuint64_t)node_stride * 4ull};
  const cuuint32_t box[3] = {(cuuint32_t)a.d, 1u, (cuuint32_t)box_rows};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int ACT, bool FUSE, int TAB>
static int launch_tma(const FastArgs& fa, const CUtensorMap& tm, const CUtensorMap& tm32, const CUtensorMap& tmp,
                      size_t staged_bytes,
                      int stage_rows, int ntiles, float* out, cudaStream_t st) {
  const kp_agg_desc& a = fa.d;
  const size_t fixed = staged_bytes + (size_t)TMA_TILE * 12 * 32 + 32 * TMA_STAGES + 16 + 512 + 128;
  const size_t total = fixed + (size_t)TMA_STAGES * (stage_rows + (a.P ? 32 : 0)) * a.d * 4;
  KP_CUDA(cudaFuncSetAttribute(agg_fwd_tma_kernel<ACT, FUSE, TAB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)total));
  const int grid = geom_cap(ntiles < kNumSMs ? ntiles : kNumSMs);
  KP_LAUNCH((agg_fwd_tma_kernel<ACT, FUSE, TAB>), grid, TMA_THREADS, total, st, fa, tm, tm32, tmp, out, stage_rows, ntiles);
  return 0;
}

// Returns true (and launches) when the TMA-staged kernel applies: d in (64,128], no extras, tables in shared
// memory or none, 32-bit row offsets, and -- unless forced -- enough tiles to give every SM a few.
bool tma_fwd(const FastArgs& fa, int G, int act, bool fuse, int tab, bool extra, size_t staged_bytes, float* out,
             cudaStream_t st, int* rc) {
  const kp_agg_desc& a = fa.d;
  if (g_tma_mode == 0 || G != 32 || a.k + 1 > 32 || extra || a.dinv || a.indeg || a.eps || tab == TAB_GLOBAL) return false;
  if ((unsigned long long)a.N * (unsigned long long)a.d * 4ull >= 0xffffffffull) return false;
  const int ntiles = (a.N + TMA_TILE - 1) / TMA_TILE;
  if (g_tma_mode == 1 && ntiles < 2 * kNumSMs) return false;
  const size_t fixed = staged_bytes + (size_t)TMA_TILE * 12 * 32 + 32 * TMA_STAGES + 16 + 512 + 128;
  const size_t budget = 227 * 1024;
  const size_t prow = a.P ? 32 : 0;                                            // the tile's own P rows ride in the stage
  if (fixed + (size_t)TMA_STAGES * (48 + prow) * a.d * 4 > budget) return false;   // fewer than 48 rows per stage: no point
  int stage_rows = (int)((budget - fixed) / ((size_t)TMA_STAGES * a.d * 4)) - (int)prow;
  if (stage_rows > 256) stage_rows = 256;
  stage_rows -= stage_rows % TMA_BOX_ROWS;
  if (fa.xh % 4 || fa.xs % 4 || ((uintptr_t)a.X & 15)) return false;
  if (a.P && (fa.ph % 4 || fa.ps % 4 || ((uintptr_t)a.P & 15))) return false;
  CUtensorMap tm, tm32, tmp;
  if (!make_map(a, a.X, fa.xs, fa.xh, TMA_BOX_ROWS, &tm)) return false;
  if (!make_map(a, a.X, fa.xs, fa.xh, 2 * TMA_BOX_ROWS, &tm32)) return false;
  tmp = tm;
  if (a.P && !make_map(a, a.P, fa.ps, fa.ph, 32, &tmp)) return false;
#define KP_TMA_TAB(A, F) \
  (tab == TAB_SMEM ? launch_tma<A, F, TAB_SMEM>(fa, tm, tm32, tmp, staged_bytes, stage_rows, ntiles, out, st) \
                   : launch_tma<A, F, TAB_NONE>(fa, tm, tm32, tmp, staged_bytes, stage_rows, ntiles, out, st))
  if (act == KP_ACT_GELU) *rc = fuse ? KP_TMA_TAB(KP_ACT_GELU, true) : KP_TMA_TAB(KP_ACT_GELU, false);
  else if (act == KP_ACT_RELU) *rc = fuse ? KP_TMA_TAB(KP_ACT_RELU, true) : KP_TMA_TAB(KP_ACT_RELU, false);
  else *rc = fuse ? KP_TMA_TAB(KP_ACT_NONE, true) : KP_TMA_TAB(KP_ACT_NONE, false);
#undef KP_TMA_TAB
  return true;
}

}  // namespace kp
