// agg_tile.cu -- block-resident aggregation for LONG rows (sm_100a).
//
// Why: on the regular-graph simulation workload (BASELINE.json configs[4]: n = 1 280, K = 6, d = 16; run_simulation.py)
// a destination node has ~177 in-entries, each a 64-byte gather.  The row-streaming kernels (agg_fast.cuh, agg_lean.cuh)
// take those gathers from L1 / L2 one dependent load at a time -- 230 GB/s of algorithmic bytes, 3.5 % of the HBM
// roofline (profiles/r1_other_configs.json) -- although the data they touch is tiny: one hop slice X[graph, :, h, :] of a
// 1 280-node graph is 80 KB.  Here a CTA owns one (closed node block, hop) unit at a time (blocks = the graphs of the
// batch, kp_plan_blocks): it copies the block's hop slice into shared memory with cp.async (each byte of X leaves
// L2 / HBM once per unit), then its lane groups walk the block's rows and take every gather from shared memory.  Units
// are dealt heavy-hop-first, round-robin over persistent CTAs; a unit's output rows are written by exactly one CTA, in
// entry order: bit-reproducible.  Unfused outputs only ([N,k,d]); the same kernel serves B2 (dX = gather of the
// hand-over gradient through the transposed CSR) in its second mode.
#include "agg_fast_host.h"
#include "agg_lean.cuh"

namespace kp {

constexpr int TILE_THREADS = 512;

struct TileArgs {
  const int32_t* rowptr;       // (dst,hop) CSR in forward mode, (src,hop) in B2 mode
  const int32_t* col;
  const uint16_t* attr16;
  const int32_t* block_ptr;
  const int32_t* block_stats;  // device [0] = number of blocks (NULL: num_blocks is exact)
  int num_blocks, dpad, stage_floats;
  int chunks;                  // row chunks per (block, hop): a unit is (chunk, block, hop)
  const float* self_src;       // B2 mode: dOut [N,k,d] for the (1+eps) self term; forward mode: unused
};

template <int G, int ACT, int TAB, bool EXTRA, bool B2>
__global__ void __launch_bounds__(TILE_THREADS)
agg_tile_kernel(const FastArgs fa, const TileArgs ta, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const kp_agg_desc& a = fa.d;
  stage_tables<TAB, false>(a, sm);
  const int d = a.d, k = a.k, Kp = a.Kplan, dq = d >> 2;
  const int lane = threadIdx.x & (G - 1);
  const bool active = lane * 4 < d;
  const unsigned c = (unsigned)min(lane * 4, d - 4);
  constexpr int gpb = TILE_THREADS / G;
  const int gib = threadIdx.x / G;
  const unsigned gm = group_mask<G>();
  const unsigned sm_base = sh_addr(sm);
  const unsigned tab0_sh = sm_base + c * 4u;
  const unsigned tabk_sh = tab0_sh + ((TAB == TAB_SMEM) ? (unsigned)(a.rows0 * d) * 4u : 0u);
  float* Xs = sm + ta.stage_floats;
  const unsigned xs_sh = sm_base + (unsigned)ta.stage_floats * 4u + c * 4u;
  const unsigned dpad4 = (unsigned)ta.dpad * 4u;
  float self_c = 0.f;
  if (EXTRA && a.eps) self_c = 1.f + __ldg(a.eps);
  const int nblocks = ta.block_stats ? __ldg(ta.block_stats) : ta.num_blocks;
  const int per_hop = nblocks * ta.chunks;
  const int units = per_hop * k;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    // far hops (the long rows of an spd plan: 96 of a node's 177 entries sit in hop 6 of a 3-regular graph) are dealt
    // first, and every (block, hop) is cut into row chunks: without the cut the 64 hop-6 units of a 64-graph batch
    // carried 54 % of the work on 64 CTAs and set the kernel time
    const int hq = u / per_hop, rem = u - hq * per_hop;
    const int h = k - 1 - hq;
    const int b = rem / ta.chunks, ch = rem - b * ta.chunks;
    const int v0 = __ldg(ta.block_ptr + b), v1 = __ldg(ta.block_ptr + b + 1);
    const int nb = v1 - v0;
    const int per = (nb + ta.chunks - 1) / ta.chunks;
    const int r0 = min(nb, ch * per), r1 = min(nb, r0 + per);
    if (r0 >= r1) continue;                               // (block-uniform: no barrier is skipped by part of a CTA)
    __syncthreads();                                      // everyone is done with the previous unit's slice
    {
      const float* src = a.X + (size_t)v0 * fa.xs + (size_t)h * fa.xh;
      for (int i = threadIdx.x; i < nb * dq; i += TILE_THREADS) {
        const int vl = i / dq, q = i - vl * dq;
        cp_async16_cg(sm_base + ((unsigned)ta.stage_floats + (unsigned)(vl * ta.dpad + 4 * q)) * 4u,
                      src + (size_t)vl * fa.xs + 4 * q);
      }
      cp_async_commit();
      cp_async_wait<0>();
    }
    __syncthreads();
    for (int vl = r0 + gib; vl < r1; vl += gpb) {
      const int v = v0 + vl;
      const int r = v * Kp + h;
      const int rb = __ldg(ta.rowptr + r), re = __ldg(ta.rowptr + r + 1);
      P4 acc = p4zero();
      int ncol = (rb + lane < re) ? __ldg(ta.col + rb + lane) : v0;      // entry indices run one window ahead
      for (int j0 = rb; j0 < re; j0 += G) {
        const int cnt = min(G, re - j0);
        unsigned my_off = 0, my_tab = 0;
        float my_w = 0.f;
        const int cj = ncol;
        if (j0 + G + lane < re) ncol = __ldg(ta.col + j0 + G + lane);
        if (lane < cnt) {
          my_off = (unsigned)(cj - v0) * dpad4;
          if (TAB != TAB_NONE) my_tab = (unsigned)__ldg(ta.attr16 + j0 + lane) * (unsigned)d * 4u;
          if (EXTRA && !B2 && a.dinv) my_w = __ldg(a.dinv + (size_t)cj * Kp + h);
        }
        int i = 0;
        for (; i + 4 <= cnt; i += 4) {
          unsigned o[4];
          P4 x[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) o[t] = __shfl_sync(gm, my_off, i + t, G);
#pragma unroll
          for (int t = 0; t < 4; ++t) x[t] = lds4p(xs_sh + o[t]);
          if (TAB == TAB_SMEM) {
#pragma unroll
            for (int t = 0; t < 4; ++t) x[t] = add4p(x[t], lds4p((h == 0 ? tab0_sh : tabk_sh) + __shfl_sync(gm, my_tab, i + t, G)));
          }
          if (EXTRA && !B2 && a.dinv) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const u64 w = splat2(__shfl_sync(gm, my_w, i + t, G));
              acc.lo = fma2(w, x[t].lo, acc.lo); acc.hi = fma2(w, x[t].hi, acc.hi);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) acc = add4p(acc, x[t]);
          }
        }
        for (; i < cnt; ++i) {
          P4 x = lds4p(xs_sh + __shfl_sync(gm, my_off, i, G));
          if (TAB == TAB_SMEM) x = add4p(x, lds4p((h == 0 ? tab0_sh : tabk_sh) + __shfl_sync(gm, my_tab, i, G)));
          if (EXTRA && !B2 && a.dinv) {
            const u64 w = splat2(__shfl_sync(gm, my_w, i, G));
            acc.lo = fma2(w, x.lo, acc.lo); acc.hi = fma2(w, x.hi, acc.hi);
          } else {
            acc = add4p(acc, x);
          }
        }
      }
      // ---- epilogue of row (v,h)
      if (EXTRA) {
        float s = 1.f;
        if (a.dinv) s *= __ldg(a.dinv + (size_t)v * Kp + h);
        if (!B2 && a.indeg) s *= 1.f / (float)max(__ldg(a.indeg + v), 1);
        const u64 sp = splat2(s);
        acc.lo = mul2(acc.lo, sp); acc.hi = mul2(acc.hi, sp);
      }
      if (!B2) {
        acc.lo = act_fwd2<ACT>(acc.lo);
        acc.hi = act_fwd2<ACT>(acc.hi);
      }
      const size_t orow = ((size_t)v * k + h) * d + c;
      if (EXTRA && a.eps) {
        P4 xv;
        if (B2) xv = ldg4p_stream(ta.self_src + orow);
        else xv = lds4p(xs_sh + (unsigned)vl * dpad4);
        const u64 sc = splat2(self_c);
        acc.lo = fma2(sc, xv.lo, acc.lo); acc.hi = fma2(sc, xv.hi, acc.hi);
      }
      if (!B2 && a.P) acc = add4p(acc, ldg4p_stream(a.P + (size_t)v * fa.ps + (size_t)h * fa.ph + c));
      if (active) stg4p_stream(out + orow, acc);
    }
  }
}

static int g_tile_mode = 1;     // 0 = never, 1 = when the caller supplies blocks (default), 2 = also for short rows (tests)
void tile_set_mode(int mode) { g_tile_mode = mode; }

// row chunks per (block, hop): about two passes of the CTA's lane groups per unit
static int tile_chunks(const kp_agg_desc& a, int G) {
  const int gpb = TILE_THREADS / G;
  int c = a.max_block_nodes / (2 * gpb);
  return c < 1 ? 1 : (c > 64 ? 64 : c);
}

static size_t tile_smem(const kp_agg_desc& a, int tab, int* dpad, int* stage_floats) {
  *dpad = a.d <= 16 ? 16 : a.d + 4;   // 4-lane groups (d <= 16): a 64-byte row stride halves the bank conflicts of two rows per quarter-warp
  *stage_floats = tab == TAB_SMEM ? (a.rows0 + a.rowsk) * a.d : 0;
  return sizeof(float) * ((size_t)*stage_floats + (size_t)a.max_block_nodes * *dpad);
}

bool tile_eligible(const kp_agg_desc& a, int tab) {
  if (g_tile_mode == 0 || !a.block_ptr || a.num_blocks <= 0 || a.max_block_nodes <= 0 || a.fuse) return false;
  if (tab == TAB_GLOBAL) return false;
  int dpad, sf;
  return tile_smem(a, tab, &dpad, &sf) <= 200 * 1024;
}

template <int G, int ACT, int TAB, bool EXTRA, bool B2>
static int tile_launch(const FastArgs& fa, const TileArgs& ta, size_t smem, float* out, cudaStream_t st) {
  if (smem > 32 * 1024)
    KP_CUDA(cudaFuncSetAttribute(agg_tile_kernel<G, ACT, TAB, EXTRA, B2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 3 ? 3 : per_sm);
  const long long units = (long long)ta.num_blocks * ta.chunks * fa.d.k;
  const int grid = geom_cap(units < (long long)kNumSMs * per_sm ? units : (long long)kNumSMs * per_sm);
  KP_LAUNCH((agg_tile_kernel<G, ACT, TAB, EXTRA, B2>), grid, TILE_THREADS, smem, st, fa, ta, out);
  return 0;
}

#define KP_TILE_TAB(G, A, X, B2V, tab, ...)                                                    \
  ((tab) == TAB_SMEM ? tile_launch<G, A, TAB_SMEM, X, B2V>(__VA_ARGS__) : tile_launch<G, A, TAB_NONE, X, B2V>(__VA_ARGS__))
#define KP_TILE_ACT(G, act, extra, tab, ...)                                                     \
  ((act) == KP_ACT_GELU ? KP_TILE_TAB(G, KP_ACT_GELU, false, false, tab, __VA_ARGS__)            \
   : (act) == KP_ACT_RELU ? KP_TILE_TAB(G, KP_ACT_RELU, true, false, tab, __VA_ARGS__)           \
   : (extra) ? KP_TILE_TAB(G, KP_ACT_NONE, true, false, tab, __VA_ARGS__)                        \
             : KP_TILE_TAB(G, KP_ACT_NONE, false, false, tab, __VA_ARGS__))

// forward (unfused): same template combinations as the fast path (agg_fast_host.h fast_combo)
int tile_fwd(const FastArgs& fa, int G, int act, int tab, bool extra, float* out, cudaStream_t st) {
  const kp_agg_desc& a = fa.d;
  TileArgs ta;
  ta.rowptr = a.rowptr; ta.col = a.col; ta.attr16 = a.attr16; ta.block_ptr = a.block_ptr;
  ta.block_stats = a.block_stats;
  ta.num_blocks = a.num_blocks; ta.self_src = nullptr;
  ta.chunks = tile_chunks(a, G);
  const size_t smem = tile_smem(a, tab, &ta.dpad, &ta.stage_floats);
  switch (G) {
    case 32: return KP_TILE_ACT(32, act, extra, tab, fa, ta, smem, out, st);
    case 16: return KP_TILE_ACT(16, act, extra, tab, fa, ta, smem, out, st);
    case 8: return KP_TILE_ACT(8, act, extra, tab, fa, ta, smem, out, st);
    default: return KP_TILE_ACT(4, act, extra, tab, fa, ta, smem, out, st);
  }
}

// B2: dX[u,h,:] = dinv[u,h] * sum_{v in rowT(u,h)} Gs[v,h,:] + (1+eps) dOut[u,h,:]   (unfused layers only)
int tile_b2(const FastArgs& fa, int G, bool extra, const float* Gs, const float* dOut, float* dX, cudaStream_t st) {
  const kp_agg_desc& a = fa.d;
  FastArgs t = fa;
  t.d.X = Gs;
  t.d.P = nullptr;
  t.d.T0 = t.d.Tk = nullptr;
  t.d.rows0 = t.d.rowsk = 0;
  t.d.indeg = nullptr;
  t.xs = (unsigned)(a.k * a.d);
  t.xh = (unsigned)a.d;
  TileArgs ta;
  ta.rowptr = a.rowptrT; ta.col = a.colT; ta.attr16 = nullptr; ta.block_ptr = a.block_ptr;
  ta.block_stats = a.block_stats;
  ta.num_blocks = a.num_blocks; ta.self_src = dOut;
  ta.chunks = tile_chunks(a, G);
  const size_t smem = tile_smem(t.d, TAB_NONE, &ta.dpad, &ta.stage_floats);
#define KP_TILE_B2(GG) \
  (extra ? tile_launch<GG, KP_ACT_NONE, TAB_NONE, true, true>(t, ta, smem, dX, st) \
         : tile_launch<GG, KP_ACT_NONE, TAB_NONE, false, true>(t, ta, smem, dX, st))
  switch (G) {
    case 32: return KP_TILE_B2(32);
    case 16: return KP_TILE_B2(16);
    case 8: return KP_TILE_B2(8);
    default: return KP_TILE_B2(4);
  }
#undef KP_TILE_B2
}

}  // namespace kp
