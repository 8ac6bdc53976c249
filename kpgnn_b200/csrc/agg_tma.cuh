// agg_tma.cuh -- fused K-hop aggregation forward with the gathered rows staged in shared memory by the TMA engine.
//
// Why: after the packed-math rewrite (agg_lean.cuh) the forward kernel is no longer issue-bound; ncu and the
// ablations in profiles/r1_fwd_ablation.txt show it waiting on the X gathers (removing the whole P stream saves
// 10 %, 40 % of the stall samples sit on the first use of a gathered row).  Every X row is first touched by
// exactly one gather, so nearly every (node,hop) row eats one DRAM round trip that only ~1 row of work per warp
// is there to hide; L2 prefetch hints and register prefetch both made it slower (profiles/r1_fwd_experiments.txt).
//
// Here no warp ever waits on DRAM for X.  A 1024-thread CTA owns a tile of 31 consecutive nodes (one per consumer
// warp); batched small graphs are contiguous node ranges, so all sources of a tile lie in a window [umin, umax] of
// a few dozen nodes.  Warp 31 is the producer: per (tile, hop) it copies the window's hop slice
// X[umin..umax, h, :] plus the tile's own P rows of that hop -- a few cp.async.bulk.tensor boxes of 32 / 16 rows
// (TMA: no registers, no L1; per-row bulk copies were issue-limited at ~25 M copies/s per SM) -- into one of S
// stage buffers and signals an mbarrier; the consumers gather from shared memory (LDS.128) exactly as the lean
// kernel gathers from global, and release the stage through a second mbarrier.  The producer runs S hops ahead,
// i.e. ~S * (window + 31) rows of DRAM latency are in flight per SM regardless of what the consumers are doing.
// The window of the NEXT tile is computed by the consumers themselves (two REDUX over the column ids they already
// hold in registers for their next node) and handed to the producer through a third mbarrier, so the producer
// never touches global memory.
// Tiles whose window does not fit a stage (arbitrary graphs) are gathered from global memory by the same code
// ("direct" tiles), so the kernel is correct for any plan; it is only profitable for locality-ordered batches.
#pragma once
#include <cuda.h>

#include "agg_lean.cuh"

namespace kp {

constexpr int TMA_STAGES = 3;
constexpr int TMA_THREADS = 1024;
constexpr int TMA_TILE = TMA_THREADS / 32 - 1;        // consumer warps = nodes per tile
constexpr int TMA_BOX_ROWS = 16;                      // stage granularity: source rows per small tensor copy (big = 32)

__device__ __forceinline__ void mbar_init(unsigned a, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned a, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned a, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(a), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a protocol bug must fault the launch, not hang the GPU
__device__ __forceinline__ void mbar_wait(unsigned a, unsigned parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(a, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
// one box {d floats, 1 hop, TMA_BOX_ROWS nodes} of X viewed as a 3-D tensor [N][k][d] -> dense rows in shared memory
__device__ __forceinline__ void tma_box_g2s(unsigned dst, const CUtensorMap* tm, int hop, int node, unsigned mbar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(tm), "r"(0), "r"(hop), "r"(node), "r"(mbar)
      : "memory");
}

// gather of NE window entries from a staged slice (entry.x = source node id, entry.y = table byte address)
template <int NE, int TAB>
__device__ __forceinline__ void tma_gather(P4& z, unsigned ent, unsigned sbase, unsigned d4, unsigned c4) {
  uint2 en[NE];
  P4 x[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) en[i] = lds2_sh(ent + 8u * i);
#pragma unroll
  for (int i = 0; i < NE; ++i) x[i] = lds4p(en[i].x * d4 + sbase);
  if (TAB == TAB_SMEM) {
#pragma unroll
    for (int i = 0; i < NE; ++i) z = add4p(z, lds4p(en[i].y + c4));
  }
#pragma unroll
  for (int i = 0; i < NE; ++i) z = add4p(z, x[i]);
}
template <int NE, int TAB>
__device__ __forceinline__ void direct_gather(P4& z, unsigned ent, const float* Xh, unsigned xs, unsigned c4) {
  uint2 en[NE];
  P4 x[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) en[i] = lds2_sh(ent + 8u * i);
#pragma unroll
  for (int i = 0; i < NE; ++i) x[i] = ldg4p(at_elem(Xh, en[i].x * xs));
  if (TAB == TAB_SMEM) {
#pragma unroll
    for (int i = 0; i < NE; ++i) z = add4p(z, lds4p(en[i].y + c4));
  }
#pragma unroll
  for (int i = 0; i < NE; ++i) z = add4p(z, x[i]);
}

// Shared memory: [tables | theta][31 windows x 384 B][stage headers S x 16 B][mbarriers 2S x 8 B][pad to 128]
//                [S stages x {stage_rows X rows, 32 P rows} x d*4 B]
template <int ACT, bool FUSE, int TAB>
__global__ void __launch_bounds__(TMA_THREADS, 1)
agg_fwd_tma_kernel(const FastArgs fa, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX32,
                   const __grid_constant__ CUtensorMap tmP, float* __restrict__ out, int stage_rows, int ntiles) {
  constexpr int G = 32;
  extern __shared__ __align__(16) float sm[];     // stage buffers are aligned by hand below
  const kp_agg_desc& a = fa.d;
  const int staged = stage_tables<TAB, FUSE>(a, sm);
  const int d = a.d, k = a.k, Kp = a.Kplan, N = a.N;
  const unsigned xs = fa.xs, d4 = (unsigned)d * 4u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned sm_base = sh_addr(sm);
  const unsigned win_all = sm_base + (unsigned)staged * 4u;
  const unsigned hdr_sh = win_all + (unsigned)TMA_TILE * 12u * G;          // per stage: {umin, direct, -, -}
  const unsigned bar_sh = hdr_sh + 16u * TMA_STAGES;                         // full[S], empty[S], range[2]
  const unsigned rbar_sh = bar_sh + 16u * TMA_STAGES;
  const unsigned rng_sh = rbar_sh + 16u;                                     // [2 tiles][32 warps] {min col, max col}
  const unsigned stage0 = (rng_sh + 2u * 32u * 8u + 127u) & ~127u;
  const bool hasP = a.P != nullptr;
  // a stage = the source window's hop slice of X (stage_rows rows) + the tile's own P rows of that hop (32 rows)
  const unsigned pofs = (unsigned)stage_rows * d4;
  const unsigned stage_bytes = pofs + (hasP ? 32u * d4 : 0u);
  if (threadIdx.x == 0) {
    for (int s = 0; s < TMA_STAGES; ++s) {
      mbar_init(bar_sh + 8u * s, 1);                                         // full: the producer's arrive (+ tx bytes)
      mbar_init(bar_sh + 8u * (TMA_STAGES + s), TMA_TILE);                   // empty: one arrive per consumer warp
    }
    mbar_init(rbar_sh, TMA_TILE);                                            // window of an even / odd tile published
    mbar_init(rbar_sh + 8u, TMA_TILE);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == TMA_TILE) {
    // ======================= producer =======================
    unsigned g = 0;                                                          // running (tile, hop) count of this CTA
    unsigned it = 0;                                                         // tile count of this CTA
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      // the tile's source window, published by the consumer warps (one {min, max} each)
      mbar_wait(rbar_sh + 8u * (it & 1u), (it >> 1) & 1u);
      const uint2 mm = lds2_sh(rng_sh + (it & 1u) * 256u + 8u * (unsigned)lane);
      const int mn = __reduce_min_sync(0xffffffffu, lane < TMA_TILE ? (int)mm.x : 0x7fffffff);
      const int mx = __reduce_max_sync(0xffffffffu, lane < TMA_TILE ? (int)mm.y : -1);
      const int lo = mx >= mn ? mn : 0;
      const int R = mx >= mn ? mx - mn + 1 : 0;
      const int nsmall = (R + TMA_BOX_ROWS - 1) / TMA_BOX_ROWS;              // 16-row units
      const bool direct = nsmall * TMA_BOX_ROWS > stage_rows;
      for (int h = 0; h < k; ++h, ++g) {
        const unsigned s = g % TMA_STAGES, ph = (g / TMA_STAGES) & 1u;
        const unsigned full = bar_sh + 8u * s, empty = bar_sh + 8u * (TMA_STAGES + s);
        mbar_wait(empty, ph ^ 1u);                                           // consumers released this stage
        if (lane == 0) {
          sts2_sh(hdr_sh + 16u * s, (unsigned)lo, direct ? 1u : 0u);
          mbar_arrive_expect_tx(full, (direct ? 0u : (unsigned)(nsmall * TMA_BOX_ROWS) * d4) + (hasP ? 32u * d4 : 0u));
          const unsigned dst0 = stage0 + s * stage_bytes;
          if (hasP) tma_box_g2s(dst0 + pofs, &tmP, h, t * TMA_TILE, full);
          if (!direct) {
            int r = 0;
            for (; r + 2 <= nsmall; r += 2)                                  // 32-row boxes, then at most one 16-row box
              tma_box_g2s(dst0 + (unsigned)(r * TMA_BOX_ROWS) * d4, &tmX32, h, lo + r * TMA_BOX_ROWS, full);
            if (r < nsmall) tma_box_g2s(dst0 + (unsigned)(r * TMA_BOX_ROWS) * d4, &tmX, h, lo + r * TMA_BOX_ROWS, full);
          }
        }
        __syncwarp();
      }
    }
    return;
  }

  // ======================= consumers: one node per warp and tile =======================
  const bool active = lane * 4 < d;
  const unsigned c = (unsigned)min(lane * 4, d - 4);
  const unsigned c4 = c * 4u;
  const unsigned tab0_sh = sm_base;
  const unsigned tabk_sh = tab0_sh + ((TAB == TAB_SMEM) ? (unsigned)(a.rows0 * d) * 4u : 0u);
  const unsigned theta_sh = sm_base + ((TAB == TAB_SMEM) ? (unsigned)((a.rows0 + a.rowsk) * d) : 0u) * 4u + c4;
  const unsigned win_sh = win_all + (unsigned)warp * (12u * G);
  const unsigned rp_sh = win_sh + 8u * G;
  const float* Xc = opaque_ptr(a.X + c);
  const int vstride = gridDim.x * TMA_TILE;

  int vn = blockIdx.x * TMA_TILE + warp;                                    // node of the first tile (may be >= N)
  int rpn = (vn < N && lane <= k) ? __ldg(a.rowptr + (size_t)vn * Kp + lane) : 0;
  int ncol = 0, nattr = 0;
  {
    const int nb = __shfl_sync(0xffffffffu, rpn, 0), ne = __shfl_sync(0xffffffffu, rpn, k);
    if (nb + lane < ne) {
      ncol = __ldg(a.col + nb + lane);
      if (TAB != TAB_NONE) nattr = (int)__ldg(a.attr16 + nb + lane);
    }
  }
  int rpnn = 0;
  {
    const int v2 = vn + vstride;
    if (v2 < N && lane <= k) rpnn = __ldg(a.rowptr + (size_t)v2 * Kp + lane);
  }
  // window hand-off: min / max source of this warp's node of tile `it`, then one arrive on the tile's barrier
  auto publish_range = [&](unsigned it, int col, bool has) {
    const int mn = __reduce_min_sync(0xffffffffu, has ? col : 0x7fffffff);
    const int mx = __reduce_max_sync(0xffffffffu, has ? col : -1);
    if (lane == 0) {
      sts2_sh(rng_sh + (it & 1u) * 256u + 8u * (unsigned)warp, (unsigned)mn, (unsigned)mx);
      mbar_arrive(rbar_sh + 8u * (it & 1u));
    }
  };
  {
    const int nb = __shfl_sync(0xffffffffu, rpn, 0), ne = __shfl_sync(0xffffffffu, rpn, k);
    publish_range(0u, ncol, nb + lane < ne);
  }
  const int h_pub = max(0, k - 1 - TMA_STAGES);      // early enough for the producer's run-ahead into the next tile
  unsigned g = 0, it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int v = vn;
    const bool live = v < N;                                                 // warps past the last node still follow
    const int nbeg = __shfl_sync(0xffffffffu, rpn, 0), nend = __shfl_sync(0xffffffffu, rpn, k);   // the barriers
    const int e1 = __shfl_sync(0xffffffffu, rpn, 1);
    __syncwarp();
    sts1_sh(rp_sh + 4u * lane, rpn - nbeg);
    sts2_sh(win_sh + 8u * lane, (unsigned)ncol, (nbeg + lane < e1 ? tab0_sh : tabk_sh) + (unsigned)nattr * d4);
    __syncwarp();
    const bool big = (nend - nbeg) > G;
    vn = v + vstride;
    rpn = rpnn;
    ncol = 0; nattr = 0;
    rpnn = 0;
    bool nhas = false;                                                       // this lane holds an entry of the next node
    const bool more = t + (int)gridDim.x < ntiles;
    if (vn < N) {
      const int nb = __shfl_sync(0xffffffffu, rpn, 0), ne = __shfl_sync(0xffffffffu, rpn, k);
      nhas = nb + lane < ne;
      if (nb + lane < ne) {
        ncol = __ldg(a.col + nb + lane);
        if (TAB != TAB_NONE) nattr = (int)__ldg(a.attr16 + nb + lane);
      }
      const int v2 = vn + vstride;
      if (v2 < N && lane <= k) rpnn = __ldg(a.rowptr + (size_t)v2 * Kp + lane);
    }

    const float* Xh = Xc;
    float* outv = out + (FUSE ? (size_t)v * d : (size_t)v * k * d) + c;
    P4 o = p4zero();
    unsigned ent = win_sh;
    unsigned th = theta_sh;
    int b = 0;
    for (int h = 0; h < k; ++h, ++g) {
      const unsigned s = g % TMA_STAGES, ph = (g / TMA_STAGES) & 1u;
      const int e = live ? lds1_sh(rp_sh + 4u * (h + 1)) : 0;
      if (h == h_pub && more) publish_range(it + 1u, ncol, nhas);
      mbar_wait(bar_sh + 8u * s, ph);                                        // the hop slice has landed
      const uint2 hdr = lds2_sh(hdr_sh + 16u * s);
      P4 p = p4zero();
      if (hasP) p = lds4p(stage0 + s * stage_bytes + pofs + (unsigned)warp * d4 + c4);
      P4 z = p4zero();
      if (hdr.y == 0u && !big) {
        const unsigned sbase = stage0 + s * stage_bytes + c4 - hdr.x * d4;   // + source id * d4 = the row's bytes
        int n = e - b;
        while (n >= 4) {
          tma_gather<4, TAB>(z, ent, sbase, d4, c4);
          ent += 32u;
          n -= 4;
        }
        if (n & 2) {
          tma_gather<2, TAB>(z, ent, sbase, d4, c4);
          ent += 16u;
        }
        if (n & 1) {
          tma_gather<1, TAB>(z, ent, sbase, d4, c4);
          ent += 8u;
        }
      } else if (!big) {
        int n = e - b;
        while (n >= 2) {
          direct_gather<2, TAB>(z, ent, Xh, xs, c4);
          ent += 16u;
          n -= 2;
        }
        if (n & 1) {
          direct_gather<1, TAB>(z, ent, Xh, xs, c4);
          ent += 8u;
        }
      } else {
        for (int j = nbeg + b; j < nbeg + e; ++j) {                          // more than 32 entries on this node
          const int cj = __ldg(a.col + j);
          P4 x = ldg4p(at_elem(Xh, (unsigned)cj * xs));
          if (TAB == TAB_SMEM) x = add4p(x, lds4p((h == 0 ? tab0_sh : tabk_sh) + c4 + (unsigned)__ldg(a.attr16 + j) * d4));
          z = add4p(z, x);
        }
      }
      b = e;
      // the gathered values are consumed by the adds above; make every lane's reads precede the release
      asm volatile("" ::"l"(z.lo), "l"(z.hi) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sh + 8u * (TMA_STAGES + s));
      z.lo = act_fwd2<ACT>(z.lo);
      z.hi = act_fwd2<ACT>(z.hi);
      z = add4p(z, p);
      if (FUSE) {
        const P4 tq = lds4p(th);
        o.lo = fma2(tq.lo, z.lo, o.lo); o.hi = fma2(tq.hi, z.hi, o.hi);
        th += d4;
      } else {
        if (active && live) stg4p_stream(outv + h * d, z);
      }
      Xh += fa.xh;
    }
    if (FUSE && active && live) stg4p_stream(outv, o);
  }
}

}  // namespace kp
