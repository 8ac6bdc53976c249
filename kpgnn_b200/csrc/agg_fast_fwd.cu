// agg_fast_fwd.cu -- instantiations of the fast forward aggregation kernel (see agg_fast.cuh).
#include "agg_fast_host.h"
#include "agg_lean.cuh"
#include <stdlib.h>

namespace kp {

template <int G, int ACT, bool FUSE, int TAB, bool EXTRA>
static int launch(const FastArgs& fa, int grid, size_t smem, float* out, cudaStream_t st) {
  if (smem > 32 * 1024)
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
  if (total > 32 * 1024)
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
  int dist = a.k >= 4 ? 1 : (a.k >= 2 ? 2 : 4);
  unsigned pfx_ = pfx, pfp_ = pfp, bulk = 0;
  // tuning knobs (profiling only): KP_LEAN_PF_MODE 0 = no L2 prefetch (default: measured fastest), 1 = per-line hints, 2 = bulk prefetch
  static const int env_mode = getenv("KP_LEAN_PF_MODE") ? atoi(getenv("KP_LEAN_PF_MODE")) : 0;
  static const int env_dist = getenv("KP_LEAN_PF_DIST") ? atoi(getenv("KP_LEAN_PF_DIST")) : 0;
  if (env_dist > 0) dist = env_dist;
  if (env_mode == 0) { pfx_ = 0; pfp_ = 0; }
  if (env_mode == 2 && pfx) bulk = (unsigned)(a.k * a.d) * 4u;
  // one 1024-thread CTA per SM when there are enough nodes to give every SM a few blocks (tables staged once per
  // SM, 32 neighbouring nodes in flight on one L1); batches that fit one wave get one equally sized CTA per SM
  // (lean_balanced_threads); 256-thread CTAs in between so that all SMs get work
  static const int env_threads = getenv("KP_LEAN_THREADS") ? atoi(getenv("KP_LEAN_THREADS")) : 0;
  const int balanced = lean_balanced_threads(a.N, G, 1);
  const int threads = g_geom_lean_threads ? g_geom_lean_threads
                      : env_threads ? env_threads
                      : balanced ? balanced
                                 : ((long long)a.N >= (long long)kNumSMs * (1024 / G) * 2 ? 1024 : 256);
  const int gpb = threads / G, ctas_per_sm = 1024 / threads;
  const size_t total = smem + (size_t)gpb * lean_group_scratch_bytes(G);   // + per-group entry window and row pointers
  const long long want = ((long long)a.N + gpb - 1) / gpb;
  grid = geom_cap(want < kNumSMs * ctas_per_sm ? (want < 1 ? 1 : want) : kNumSMs * ctas_per_sm);   // persistent
  if (total > 32 * 1024)
    KP_CUDA(cudaFuncSetAttribute(agg_fwd_lean_kernel<G, ACT, FUSE, TAB, EXTRA>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));
  KP_LAUNCH_PDL((agg_fwd_lean_kernel<G, ACT, FUSE, TAB, EXTRA>), grid, threads, total, st, fa, out, pfx_, pfp_, dist, bulk);
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
bool fast_lean_enabled() { return g_use_lean != 0; }

// B2 (dX = gather of Gs through the transposed CSR) without self term / norm is the same computation as the
// forward with no tables, no activation, no P and an unfused [N,k,d] output: run it on the lean forward kernel.
bool lean_b2(const FastArgs& fa, int G, const float* Gs, float* dX, cudaStream_t st, int* rc) {
  const kp_agg_desc& a = fa.d;
  if (!g_use_lean || !(G == 32 || G == 16) || a.k + 1 > G) return false;
  FastArgs t = fa;
  t.d.rowptr = a.rowptrT;
  t.d.col = a.colT;
  t.d.attr16 = nullptr;
  t.d.dinv = nullptr;
  t.d.indeg = nullptr;
  t.d.X = Gs;
  t.d.P = nullptr;
  t.d.T0 = nullptr;
  t.d.Tk = nullptr;
  t.d.rows0 = t.d.rowsk = 0;
  t.d.theta = nullptr;
  t.d.eps = nullptr;
  t.d.act = KP_ACT_NONE;
  t.d.fuse = 0;
  t.xs = (unsigned)(a.k * a.d);
  t.xh = (unsigned)a.d;
  t.ps = t.ph = 0;
  t.os = (unsigned)a.dx_node_stride;
  t.oh = (unsigned)a.dx_hop_stride;
  t.oacc = a.dx_accumulate;
  *rc = (G == 32) ? launch_lean<32, KP_ACT_NONE, false, TAB_NONE, false>(t, 0, 0, dX, st)
                  : launch_lean<16, KP_ACT_NONE, false, TAB_NONE, false>(t, 0, 0, dX, st);
  return true;
}

int fast_fwd(const FastArgs& fa, int G, int act, bool fuse, int tab, bool extra, int grid, size_t smem, float* out,
             cudaStream_t st) {
  if (g_use_lean) {
    int rc = 0;
    if (tma_fwd(fa, G, act, fuse, tab, extra, smem, out, st, &rc)) return rc;
  }
  if (g_use_lean && (G == 32 || G == 16) && fa.d.k + 1 <= G && !fa.d.dinv && tab != TAB_GLOBAL) {
    if (G == 32) return KP_LEAN_COMBO(32, act, fuse, extra, tab, fa, grid, smem, out, st);
    return KP_LEAN_COMBO(16, act, fuse, extra, tab, fa, grid, smem, out, st);
  }
  if (g_use_ring && G == 32 && fa.d.k <= 31 && smem + 8 * 9 * 512 <= 100 * 1024)
    return KP_FAST_COMBO(launch_ring, 32, act, fuse, extra, tab, fa, grid, smem, out, st);
  return KP_FAST_G(launch, G, act, fuse, extra, tab, fa, grid, smem, out, st);
}

}  // namespace kp
