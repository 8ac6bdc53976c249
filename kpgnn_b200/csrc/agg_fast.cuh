// agg_fast.cuh -- instruction-lean float4 kernels for the K-hop aggregation (the common case: d % 4 == 0,
// d <= 128, 16-byte aligned operands, tensors below 2^32 elements).  The generic kernels in agg.cu cover
// everything else.
//
// Why: the first ncu captures (profiles/r1_agg_fwd.md) showed DRAM traffic == algorithmic bytes but the kernel
// ISSUE-bound (70-75 % issue-active; 360-430 warp instructions per (node,hop) row): erff / __frcp_rn / __expf
// expansions, 64-bit address arithmetic, predicated-off code for optional features, divergence bookkeeping
// around the 6 idle lanes of a 104-wide row.  This version:
//   * a group of G lanes (G = 4/8/16/32, 4*G >= d) owns one destination node; lane l owns channels 4l..4l+3;
//     lanes past the row's width recompute the last chunk (same addresses, no extra traffic) so the hot loop
//     has no divergent regions -- only the final stores are predicated;
//   * the hop segment's (col, attr) entries are loaded ONCE, coalesced, by the group's lanes and broadcast with
//     width-G shuffles; the next hop's entries, its row pointer and the P row are prefetched before the
//     current hop's gathers are consumed (software pipelining across hops);
//   * embedding tables and theta are staged in shared memory once per persistent CTA: a lookup is one LDS.128;
//   * every gather address is base + 32-bit element offset (one IMAD + one IMAD.WIDE);
//   * gathers are issued two at a time; X/Gs rows use the default (L1-allocating) path because a row is reused
//     by ~2.5 destination rows of the same small graph, P/dOut/outputs use streaming loads/stores;
//   * optional features (GCN norm, SAGE mean, GIN self term) are compiled out unless EXTRA;
//   * GELU through the erfc form with raw MUFU.RCP / MUFU.EX2 (agg_common.cuh).
#pragma once
#include <stdlib.h>
#include "agg_common.cuh"

namespace kp {

#ifndef KP_CHUNK_ITERS
#define KP_CHUNK_ITERS 1
#endif
#ifndef KP_FWD_MINB
#define KP_FWD_MINB 4
#endif

enum { TAB_NONE = 0, TAB_SMEM = 1, TAB_GLOBAL = 2 };

struct FastArgs {
  kp_agg_desc d;
  unsigned xs, xh, ps, ph;   // element strides of X and P (host-validated: every offset < 2^32)
  // unfused [N,k,d] output of the lean kernel (forward without combine, and B2's dX): node / hop strides in elements
  // (0 = contiguous) and accumulate-into instead of overwrite (dX of the layer-history buffer, kpgnn_b200/stack.py)
  unsigned os = 0, oh = 0;
  int oacc = 0;
};

// Lean kernels (agg_lean.cuh): entries of one destination node held in the group's shared-memory window.  The first
// G entries arrive through the one-node-ahead register prefetch, the rest (a 34-atom molecule at K = 8 has nodes with
// 33 in-entries) are fetched with coalesced loads when the node is published; only nodes with more than kLeanWin
// entries take the entry-by-entry path.  Per group: kLeanWin x {X element offset, table byte address} + G row pointers.
constexpr int kLeanWin = 64;
__host__ __device__ constexpr unsigned lean_group_scratch_bytes(int G) { return 8u * kLeanWin + 4u * (unsigned)G; }

// Launch geometry of the lean kernels for batches that fit ONE wave (N <= kNumSMs * 1024 / G nodes): CTAs sized so
// that every SM gets exactly one CTA with the same number of lane groups -- ceil(N / kNumSMs) groups, threads a
// multiple of 32, at least 256.  At the bench batch (2 986 nodes) that is 143 CTAs x 672 threads instead of 374 x 256
// (2 or 3 CTAs per SM: the 3-CTA SMs set the kernel time) or 94 x 1024 (54 SMs idle).  Returns 0 when the batch does
// not fit one wave (the caller keeps its persistent configuration).  OPT-IN through the environment variable
// KP_LEAN_BALANCED (bit mask: 1 = forward / B2 launches, 2 = B1 launches; default 0): measured 0.8 % faster on the
// training step (profiles/r1zzz_small_batch.txt), but the GPU budget ran out before
// tests/test_layers_gpu.py::test_bench_size_parity could be run with it, and nothing unverified ships as a default.
inline int lean_balanced_threads(long long N, int G, int which) {
  static const int mask = getenv("KP_LEAN_BALANCED") ? atoi(getenv("KP_LEAN_BALANCED")) : 0;
  if (G != 32) return 0;      // measured and parity-tested at one-wave sizes for 32-lane groups only (64 < d <= 128)
  if (!(mask & which) || N <= 0 || N > (long long)kNumSMs * (1024 / G)) return 0;
  long long gpb = (N + kNumSMs - 1) / kNumSMs;
  int threads = (int)((gpb * G + 31) / 32 * 32);
  if (threads < 256) threads = 256;
  return threads > 1024 ? 1024 : threads;
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4s(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 lds4_sh(unsigned addr) {   // explicit LDS.128 with a 32-bit shared address
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts4_sh(unsigned addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ unsigned sh_addr(const void* p) {
  unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("" : "+r"(a));        // opaque: keep it in a register instead of re-deriving it (S2R + LEA) per use
  return a;
}
__device__ __forceinline__ void st4s(float* p, const float4& v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void add4(float4& a, const float4& b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
__device__ __forceinline__ void fma4(float4& a, float w, const float4& b) {
  a.x = fmaf(w, b.x, a.x); a.y = fmaf(w, b.y, a.y); a.z = fmaf(w, b.z, a.z); a.w = fmaf(w, b.w, a.w);
}

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if (G == 32) return 0xffffffffu;
  return ((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) & ~(G - 1));
}

// Stages tables (TAB_SMEM) and theta (NEED_THETA) at the start of dynamic shared memory.
// Layout: [T0 rows0*d][Tk rowsk*d][theta k*d].  Returns the number of floats staged.
// The three sources are walked as ONE virtual array and every thread issues up to KP_STAGE_BATCH 16-byte loads before
// its first store: at the bench batch (27 KB per CTA, 256 threads) the copy is one L2 round trip instead of eight
// dependent ones (22 % of the forward kernel's stall samples, profiles/r1zzz_small_batch.txt).
#ifndef KP_STAGE_BATCH
#define KP_STAGE_BATCH 8
#endif
template <int TAB, bool NEED_THETA>
__device__ __forceinline__ int stage_tables(const kp_agg_desc& a, float* sm) {
  const int n0 = (TAB == TAB_SMEM) ? a.rows0 * a.d : 0;
  const int n1 = n0 + ((TAB == TAB_SMEM) ? a.rowsk * a.d : 0);
  const int total = n1 + (NEED_THETA ? a.k * a.d : 0);
  const int step = blockDim.x * 4;
  for (int base = threadIdx.x * 4; base < total; base += step * KP_STAGE_BATCH) {
    float4 v[KP_STAGE_BATCH];
#pragma unroll
    for (int j = 0; j < KP_STAGE_BATCH; ++j) {
      const int i = base + j * step;
      if (i < total) v[j] = ld4(i < n0 ? a.T0 + i : (i < n1 ? a.Tk + (i - n0) : a.theta + (i - n1)));
    }
#pragma unroll
    for (int j = 0; j < KP_STAGE_BATCH; ++j) {
      const int i = base + j * step;
      if (i < total) st4(sm + i, v[j]);
    }
  }
  __syncthreads();
  return total;
}

// Gathers one node's hop segments.  The node's entry list (all hops, contiguous in the plan) is read through a
// running window of G entries held one-per-lane in registers (for G = 32 a ZINC-shaped node's ~20 entries are
// one coalesced load); entries are broadcast with width-G shuffles.  The first two rows of the NEXT segment are
// requested before the current segment's epilogue runs, so their latency hides behind the activation math.
template <int G, int TAB, bool NORM>
struct SegGather {
  // per kernel
  const float* Xb;
  const int* col;
  const uint16_t* attr;
  const float* dinv;
  unsigned xs, gm;
  int Kp, d, lane;
  // per node
  int wb, pc, pa, nend;
  float4 xa, xb;
  int npre;

  __device__ __forceinline__ void load_window() {
    pc = 0;
    pa = 0;
    const unsigned i = (unsigned)(wb + lane);
    if ((int)i < nend) {
      pc = __ldg(col + i);
      if (TAB != TAB_NONE) pa = (int)__ldg(attr + i);
    }
  }
  __device__ __forceinline__ void open(int start, int node_end) {
    wb = start;
    nend = node_end;
    npre = 0;
    load_window();
  }
  __device__ __forceinline__ unsigned src(int j) const { return (unsigned)__shfl_sync(gm, pc, j - wb, G); }
  __device__ __forceinline__ int att(int j) const { return __shfl_sync(gm, pa, j - wb, G); }
  __device__ __forceinline__ float4 table(int a, const float* Tg, unsigned Tsh) const {
    if (TAB == TAB_SMEM) return lds4_sh(Tsh + (unsigned)(a * d) * 4u);
    return ld4(Tg + a * d);
  }
  // request the first (up to) two rows of segment [b,e); group-uniform control flow
  __device__ __forceinline__ void prefetch(int b, int e, unsigned xoff) {
    npre = 0;
    if (b < e) {
      if (b - wb >= G) {
        wb += G;
        load_window();
      }
      xa = ld4(Xb + (src(b) * xs + xoff));
      npre = 1;
      if (b + 1 < e && b + 1 - wb < G) {
        xb = ld4(Xb + (src(b + 1) * xs + xoff));
        npre = 2;
      }
    }
  }
  __device__ __forceinline__ void accum(float4& acc, float4 x, int j, int h, const float* Tg, unsigned Tsh) const {
    if (TAB != TAB_NONE) add4(x, table(att(j), Tg, Tsh));
    if (NORM && dinv) fma4(acc, __ldg(dinv + (size_t)src(j) * Kp + h), x);
    else add4(acc, x);
  }
  // sum of segment [b,e) (the one last passed to prefetch)
  __device__ __forceinline__ float4 consume(int b, int e, unsigned xoff, int h, const float* Tg, unsigned Tsh) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int j = b;
    if (npre >= 1) {
      if (NORM && dinv) {
        accum(acc, xa, j, h, Tg, Tsh);
      } else {
        acc = xa;
        if (TAB != TAB_NONE) add4(acc, table(att(j), Tg, Tsh));
      }
      ++j;
      if (npre == 2) {
        accum(acc, xb, j, h, Tg, Tsh);
        ++j;
      }
    }
    while (j < e) {
      if (j - wb >= G) {
        wb += G;
        load_window();
      }
      if (j + 1 < e && j + 1 - wb < G) {
        const float4 x0 = ld4(Xb + (src(j) * xs + xoff));
        const float4 x1 = ld4(Xb + (src(j + 1) * xs + xoff));
        accum(acc, x0, j, h, Tg, Tsh);
        accum(acc, x1, j + 1, h, Tg, Tsh);
        j += 2;
      } else {
        const float4 x0 = ld4(Xb + (src(j) * xs + xoff));
        accum(acc, x0, j, h, Tg, Tsh);
        ++j;
      }
    }
    npre = 0;
    return acc;
  }
};

template <bool EXTRA>
__device__ __forceinline__ float fast_row_scale(const kp_agg_desc& a, int v, int h) {
  float s = 1.f;
  if (EXTRA) {
    if (a.dinv) s *= __ldg(a.dinv + (size_t)v * a.Kplan + h);
    if (a.indeg) s *= 1.f / (float)max(__ldg(a.indeg + v), 1);
  }
  return s;
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int G, int ACT, bool FUSE, int TAB, bool EXTRA>
__global__ void __launch_bounds__(256, KP_FWD_MINB) agg_fwd_fast_kernel(const FastArgs fa, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const kp_agg_desc& a = fa.d;
  stage_tables<TAB, FUSE>(a, sm);
  const int d = a.d, k = a.k, Kp = a.Kplan;
  const unsigned xh = fa.xh;
  const int lane = threadIdx.x & (G - 1);
  const bool active = lane * 4 < d;
  const unsigned c = (unsigned)min(lane * 4, d - 4);          // idle lanes shadow the last chunk
  constexpr int gpb = 256 / G;
  const int gib = threadIdx.x / G;
  const unsigned sm_base = sh_addr(sm);
  const unsigned n0f = (TAB == TAB_SMEM) ? (unsigned)(a.rows0 * d) : 0u;
  const unsigned theta_sh = sm_base + ((TAB == TAB_SMEM) ? (unsigned)((a.rows0 + a.rowsk) * d) : 0u) * 4u + c * 4u;
  float self_c = 0.f;
  if (EXTRA && a.eps) self_c = 1.f + __ldg(a.eps);
  SegGather<G, TAB, EXTRA> sg;
  sg.Xb = a.X; sg.col = a.col; sg.attr = a.attr16; sg.dinv = EXTRA ? a.dinv : nullptr;
  sg.xs = fa.xs; sg.gm = group_mask<G>(); sg.Kp = Kp; sg.d = d; sg.lane = lane;
  // each CTA owns a CONTIGUOUS node range: the nodes of one small graph (whose rows gather each other) stay on
  // one SM, so re-gathered X rows hit that SM's L1 and the index arrays are read sequentially
  // CTAs walk the node list in chunks of KP_CHUNK consecutive nodes (about one small graph, whose rows gather
  // each other -> L1 hits), chunks are dealt round-robin so all SMs sweep HBM in step
  constexpr int chunk = KP_CHUNK_ITERS * gpb;
  for (int v0 = blockIdx.x * chunk; v0 < a.N; v0 += gridDim.x * chunk)
  for (int v = v0 + gib; v < min(a.N, v0 + chunk); v += gpb) {
    const int* rp = a.rowptr + (size_t)v * Kp;
    int e0 = __ldg(rp), e1 = __ldg(rp + 1), e2 = __ldg(rp + min(2, k));
    sg.open(e0, __ldg(rp + k));
    sg.prefetch(e0, e1, c);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* Pv = a.P ? a.P + ((size_t)v * fa.ps + c) : nullptr;
    float* outv = out + (FUSE ? (size_t)v * d : (size_t)v * k * d) + c;
    float4 pn = make_float4(0.f, 0.f, 0.f, 0.f);
    if (Pv) pn = ld4s(Pv);
    for (int h = 0; h < k; ++h) {
      const int e3 = __ldg(rp + min(h + 3, k));                // row pointers run two hops ahead
      const float4 p = pn;                                     // the streamed P row runs one hop ahead
      if (Pv && h + 1 < k) pn = ld4s(Pv + (h + 1) * fa.ph);
      const unsigned xoff = h * xh + c;
      const float* Tg = (TAB == TAB_GLOBAL) ? (h == 0 ? a.T0 : a.Tk) + c : nullptr;
      const unsigned Tsh = sm_base + ((h == 0 ? 0u : n0f) + c) * 4u;
      float4 z = sg.consume(e0, e1, xoff, h, Tg, Tsh);
      if (h + 1 < k) sg.prefetch(e1, e2, xoff + xh);          // next segment's rows fly during the epilogue
      if (EXTRA) {
        const float s = fast_row_scale<EXTRA>(a, v, h);
        z.x *= s; z.y *= s; z.z *= s; z.w *= s;
      }
      z.x = act_fwd<ACT>(z.x); z.y = act_fwd<ACT>(z.y); z.z = act_fwd<ACT>(z.z); z.w = act_fwd<ACT>(z.w);
      if (EXTRA && a.eps) fma4(z, self_c, ld4(a.X + ((unsigned)v * fa.xs + xoff)));
      if (FUSE) {
        const float4 th = lds4_sh(theta_sh + (unsigned)(h * d) * 4u);
        o.x = fmaf(th.x, z.x, fmaf(th.x, p.x, o.x)); o.y = fmaf(th.y, z.y, fmaf(th.y, p.y, o.y));
        o.z = fmaf(th.z, z.z, fmaf(th.z, p.z, o.z)); o.w = fmaf(th.w, z.w, fmaf(th.w, p.w, o.w));
      } else {
        add4(z, p);
        if (active) st4s(outv + h * d, z);
      }
      e0 = e1; e1 = e2; e2 = e3;
    }
    if (FUSE && active) st4s(outv, o);
  }
}

// ------------------------------------------------------------------------------------------------------------
// forward, asynchronous-copy variant (G = 32, i.e. 68 <= d <= 128).  The registers of the plain kernel cap the
// SM at 32 warps, and with ~1 us loaded-memory latency the two rows + one P row a warp could keep in flight in
// registers left 37 % of issue slots empty (profiles/r1g_agg_fwd.txt).  Here every warp owns a 9-slot ring in
// shared memory (3 hops in flight x {row 0, row 1, P row}) filled by cp.async (LDGSTS, no registers held): the
// P row and the first two gathered rows of hop h+2 are requested while hop h is being reduced.  Each lane copies
// and later reads back only its own 16 bytes, so cp.async.wait_group is the only synchronisation.
// ------------------------------------------------------------------------------------------------------------
#define KP_RING_SLOT_BYTES 416u      /* 26 lanes x 16 B: a 104-float row; rows up to 128 floats use 512 */
__device__ __forceinline__ void cp_async16_ca(unsigned dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async16_cg(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

template <int ACT, bool FUSE, int TAB, bool EXTRA>
__global__ void __launch_bounds__(256, 4) agg_fwd_ring_kernel(const FastArgs fa, float* __restrict__ out, int stage_floats,
                                                              unsigned slot_bytes) {
  constexpr int G = 32;
  extern __shared__ __align__(16) float sm[];
  const kp_agg_desc& a = fa.d;
  stage_tables<TAB, FUSE>(a, sm);
  const int d = a.d, k = a.k, Kp = a.Kplan;
  const unsigned xs = fa.xs, xh = fa.xh;
  const int lane = threadIdx.x & 31;
  const bool active = lane * 4 < d;
  const unsigned c = (unsigned)min(lane * 4, d - 4);
  const int wib = threadIdx.x >> 5;
  const unsigned sm_base = sh_addr(sm);
  const unsigned n0f = (TAB == TAB_SMEM) ? (unsigned)(a.rows0 * d) : 0u;
  const unsigned theta_sh = sm_base + ((TAB == TAB_SMEM) ? (unsigned)((a.rows0 + a.rowsk) * d) : 0u) * 4u + c * 4u;
  // this lane's byte inside slot s of the warp's ring: ring + (wib*9 + s)*slot_bytes + lane*16
  const unsigned ring = sm_base + (unsigned)stage_floats * 4u + (unsigned)(wib * 9) * slot_bytes +
                        min((unsigned)lane * 16u, slot_bytes - 16u);   // idle lanes alias the last chunk (never past the ring)
  float self_c = 0.f;
  if (EXTRA && a.eps) self_c = 1.f + __ldg(a.eps);
  SegGather<G, TAB, EXTRA> sg;
  sg.Xb = a.X; sg.col = a.col; sg.attr = a.attr16; sg.dinv = EXTRA ? a.dinv : nullptr;
  sg.xs = xs; sg.gm = 0xffffffffu; sg.Kp = Kp; sg.d = d; sg.lane = lane;
  for (int v = blockIdx.x * 8 + wib; v < a.N; v += gridDim.x * 8) {
    const int* rp = a.rowptr + (size_t)v * Kp;
    const int rpv = (lane <= k) ? __ldg(rp + lane) : 0;             // all k+1 row pointers, one coalesced load
    const int nbeg = __shfl_sync(0xffffffffu, rpv, 0), nend = __shfl_sync(0xffffffffu, rpv, k);
    sg.open(nbeg, nend);
    const bool fits = (nend - nbeg) <= G;                            // whole node inside one entry window
    const float* Pv = a.P ? a.P + ((size_t)v * fa.ps + c) : nullptr;
    float* outv = out + (FUSE ? (size_t)v * d : (size_t)v * k * d) + c;
    auto issue = [&](int h) {                                        // requests for hop h -> ring slots (h%3)*3+{0,1,2}
      const unsigned slot = ring + (unsigned)((h % 3) * 3) * slot_bytes;
      if (fits) {
        const int b = __shfl_sync(0xffffffffu, rpv, h), e = __shfl_sync(0xffffffffu, rpv, h + 1);
        const unsigned xoff = h * xh + c;
        if (b < e) {
          const unsigned u0 = sg.src(b);
          if (active) cp_async16_ca(slot, a.X + (u0 * xs + xoff));
          if (b + 1 < e) {
            const unsigned u1 = sg.src(b + 1);
            if (active) cp_async16_ca(slot + slot_bytes, a.X + (u1 * xs + xoff));
          }
        }
      }
      if (Pv && active) cp_async16_cg(slot + 2u * slot_bytes, Pv + h * fa.ph);
      cp_async_commit();
    };
    issue(0);
    if (k > 1) issue(1); else cp_async_commit();
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int h = 0; h < k; ++h) {
      if (h + 2 < k) issue(h + 2); else cp_async_commit();
      cp_async_wait<2>();                                            // hop h's group has landed
      const int b = __shfl_sync(0xffffffffu, rpv, h), e = __shfl_sync(0xffffffffu, rpv, h + 1);
      const unsigned slot = ring + (unsigned)((h % 3) * 3) * slot_bytes;
      const unsigned xoff = h * xh + c;
      const float* Tg = (TAB == TAB_GLOBAL) ? (h == 0 ? a.T0 : a.Tk) + c : nullptr;
      const unsigned Tsh = sm_base + ((h == 0 ? 0u : n0f) + c) * 4u;
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      int j = b;
      if (fits) {
        if (b < e) {
          sg.accum(z, lds4_sh(slot), b, h, Tg, Tsh);
          j = b + 1;
          if (b + 1 < e) {
            sg.accum(z, lds4_sh(slot + slot_bytes), b + 1, h, Tg, Tsh);
            j = b + 2;
          }
        }
      }
      if (j < e) {
        sg.npre = 0;
        float4 rest = sg.consume(j, e, xoff, h, Tg, Tsh);
        add4(z, rest);
      }
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (Pv) p = lds4_sh(slot + 2u * slot_bytes);
      if (EXTRA) {
        const float s = fast_row_scale<EXTRA>(a, v, h);
        z.x *= s; z.y *= s; z.z *= s; z.w *= s;
      }
      z.x = act_fwd<ACT>(z.x); z.y = act_fwd<ACT>(z.y); z.z = act_fwd<ACT>(z.z); z.w = act_fwd<ACT>(z.w);
      if (EXTRA && a.eps) fma4(z, self_c, ld4(a.X + ((unsigned)v * xs + xoff)));
      if (FUSE) {
        const float4 th = lds4_sh(theta_sh + (unsigned)(h * d) * 4u);
        o.x = fmaf(th.x, z.x, fmaf(th.x, p.x, o.x)); o.y = fmaf(th.y, z.y, fmaf(th.y, p.y, o.y));
        o.z = fmaf(th.z, z.z, fmaf(th.z, p.z, o.z)); o.w = fmaf(th.w, z.w, fmaf(th.w, p.w, o.w));
      } else {
        add4(z, p);
        if (active) st4s(outv + h * d, z);
      }
    }
    cp_async_wait<0>();
    if (FUSE && active) st4s(outv, o);
  }
}

// ------------------------------------------------------------------------------------------------------------
// B1: per destination row -- recompute, Gs, dP, dtheta / deps partials
// ------------------------------------------------------------------------------------------------------------
template <int G, int ACT, bool FUSE, int TAB, bool EXTRA>
__global__ void __launch_bounds__(256, 3)
agg_bwd_dst_fast_kernel(const FastArgs fa, const float* __restrict__ dOut, float* __restrict__ Gs,
                        float* __restrict__ dP, float* __restrict__ dtheta_part, float* __restrict__ deps_part) {
  extern __shared__ __align__(16) float sm[];
  const kp_agg_desc& a = fa.d;
  const int staged = stage_tables<TAB, FUSE>(a, sm);
  const int d = a.d, k = a.k, Kp = a.Kplan;
  const unsigned xh = fa.xh;
  const int lane = threadIdx.x & (G - 1);
  const bool active = lane * 4 < d;
  const unsigned c = (unsigned)min(lane * 4, d - 4);
  constexpr int gpb = 256 / G;
  constexpr int dpad = 4 * G;
  const int gib = threadIdx.x / G;
  const unsigned sm_base = sh_addr(sm);
  const unsigned n0f = (TAB == TAB_SMEM) ? (unsigned)(a.rows0 * d) : 0u;
  const unsigned theta_sh = sm_base + ((TAB == TAB_SMEM) ? (unsigned)((a.rows0 + a.rowsk) * d) : 0u) * 4u + c * 4u;
  float self_c = 0.f;
  if (EXTRA && a.eps) self_c = 1.f + __ldg(a.eps);
  const bool need_z = FUSE && dtheta_part != nullptr;
  const bool recompute = (ACT != KP_ACT_NONE) || need_z;
  // per-group private dtheta accumulators [k][dpad] behind the staged data; lane l owns columns 4l..4l+3 of its
  // group's copy (idle lanes write padding columns that are never read back)
  float* th_all = sm + staged;
  const unsigned th_sh = sh_addr(th_all) + (unsigned)((gib * k) * dpad + lane * 4) * 4u;
  if (dtheta_part) {
    for (int i = threadIdx.x; i < gpb * k * dpad; i += blockDim.x) th_all[i] = 0.f;
    __syncthreads();
  }
  SegGather<G, TAB, EXTRA> sg;
  sg.Xb = a.X; sg.col = a.col; sg.attr = a.attr16; sg.dinv = EXTRA ? a.dinv : nullptr;
  sg.xs = fa.xs; sg.gm = group_mask<G>(); sg.Kp = Kp; sg.d = d; sg.lane = lane;
  double eps_acc = 0.0;   // scalar reduction over N*k*d products: double keeps it within the 1e-5 parity bar
  constexpr int chunk = KP_CHUNK_ITERS * gpb;
  for (int v0 = blockIdx.x * chunk; v0 < a.N; v0 += gridDim.x * chunk)
  for (int v = v0 + gib; v < min(a.N, v0 + chunk); v += gpb) {
    const int* rp = a.rowptr + (size_t)v * Kp;
    int e0 = 0, e1 = 0, e2 = 0;
    if (recompute) {
      e0 = __ldg(rp); e1 = __ldg(rp + 1); e2 = __ldg(rp + min(2, k));
      sg.open(e0, __ldg(rp + k));
      sg.prefetch(e0, e1, c);
    }
    float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
    if (FUSE) go = ld4s(dOut + ((size_t)v * d + c));
    const float* Pv = (need_z && a.P) ? a.P + ((size_t)v * fa.ps + c) : nullptr;
    const size_t row0 = (size_t)v * k * d + c;
    for (int h = 0; h < k; ++h) {
      int e3 = 0;
      if (recompute) e3 = __ldg(rp + min(h + 3, k));
      const size_t row = row0 + (size_t)h * d;
      float4 dy;
      if (FUSE) {
        const float4 th = lds4_sh(theta_sh + (unsigned)(h * d) * 4u);
        dy = make_float4(th.x * go.x, th.y * go.y, th.z * go.z, th.w * go.w);
      } else {
        dy = ld4s(dOut + row);
      }
      if (dP && active) st4s(dP + row, dy);
      const unsigned xoff = h * xh + c;
      float4 g = dy;
      float s = 1.f;
      if (EXTRA) s = fast_row_scale<EXTRA>(a, v, h);
      if (recompute) {
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (Pv) p = ld4s(Pv + h * fa.ph);
        const float* Tg = (TAB == TAB_GLOBAL) ? (h == 0 ? a.T0 : a.Tk) + c : nullptr;
        const unsigned Tsh = sm_base + ((h == 0 ? 0u : n0f) + c) * 4u;
        float4 pre = sg.consume(e0, e1, xoff, h, Tg, Tsh);
        if (h + 1 < k) sg.prefetch(e1, e2, xoff + xh);
        if (EXTRA) {
          pre.x *= s; pre.y *= s; pre.z *= s; pre.w *= s;
        }
        if (need_z) {
          float4 z = make_float4(act_fwd<ACT>(pre.x) + p.x, act_fwd<ACT>(pre.y) + p.y, act_fwd<ACT>(pre.z) + p.z,
                                 act_fwd<ACT>(pre.w) + p.w);
          if (EXTRA && a.eps) fma4(z, self_c, ld4(a.X + ((unsigned)v * fa.xs + xoff)));
          float4 t = lds4_sh(th_sh + (unsigned)(h * dpad) * 4u);
          t.x = fmaf(go.x, z.x, t.x); t.y = fmaf(go.y, z.y, t.y);
          t.z = fmaf(go.z, z.z, t.z); t.w = fmaf(go.w, z.w, t.w);
          sts4_sh(th_sh + (unsigned)(h * dpad) * 4u, t);
        }
        g.x *= act_bwd<ACT>(pre.x); g.y *= act_bwd<ACT>(pre.y);
        g.z *= act_bwd<ACT>(pre.z); g.w *= act_bwd<ACT>(pre.w);
        e0 = e1; e1 = e2; e2 = e3;
      }
      if (EXTRA && deps_part && active) {
        const float4 x = ld4(a.X + ((unsigned)v * fa.xs + xoff));
        eps_acc += (double)(dy.x * x.x + dy.y * x.y) + (double)(dy.z * x.z + dy.w * x.w);
      }
      if (Gs && active) {
        if (EXTRA) {
          g.x *= s; g.y *= s; g.z *= s; g.w *= s;
        }
        st4(Gs + row, g);
      }
    }
  }
  if (dtheta_part) {
    __syncthreads();
    for (int i = threadIdx.x; i < k * d; i += blockDim.x) {   // fixed-order reduction over the CTA's groups
      const int h = i / d, cc = i - h * d;
      float s = 0.f;
      for (int g = 0; g < gpb; ++g) s += th_all[((size_t)g * k + h) * dpad + cc];
      dtheta_part[(size_t)blockIdx.x * k * d + i] = s;
    }
  }
  if (EXTRA && deps_part) {
    __shared__ double red[256];
    __syncthreads();
    red[threadIdx.x] = eps_acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) deps_part[blockIdx.x] = (float)red[0];
  }
}

// ------------------------------------------------------------------------------------------------------------
// B2: per source row, gather Gs ([N,k,d] contiguous) through the transposed CSR
// ------------------------------------------------------------------------------------------------------------
template <int G, bool FUSE, bool EXTRA>
__global__ void __launch_bounds__(256, 4)
agg_bwd_src_fast_kernel(const FastArgs fa, const float* __restrict__ Gs, const float* __restrict__ dOut,
                        float* __restrict__ dX) {
  const kp_agg_desc& a = fa.d;
  const int d = a.d, k = a.k, Kp = a.Kplan;
  const int lane = threadIdx.x & (G - 1);
  const bool active = lane * 4 < d;
  const unsigned c = (unsigned)min(lane * 4, d - 4);
  constexpr int gpb = 256 / G;
  const int gib = threadIdx.x / G;
  float self_c = 0.f;
  if (EXTRA && a.eps) self_c = 1.f + __ldg(a.eps);
  SegGather<G, TAB_NONE, false> sg;
  sg.Xb = Gs; sg.col = a.colT; sg.attr = nullptr; sg.dinv = nullptr;
  sg.xs = (unsigned)(k * d); sg.gm = group_mask<G>(); sg.Kp = Kp; sg.d = d; sg.lane = lane;
  constexpr int chunk = KP_CHUNK_ITERS * gpb;
  for (int u0 = blockIdx.x * chunk; u0 < a.N; u0 += gridDim.x * chunk)
  for (int u = u0 + gib; u < min(a.N, u0 + chunk); u += gpb) {
    const int* rp = a.rowptrT + (size_t)u * Kp;
    int e0 = __ldg(rp), e1 = __ldg(rp + 1), e2 = __ldg(rp + min(2, k));
    sg.open(e0, __ldg(rp + k));
    sg.prefetch(e0, e1, c);
    float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EXTRA && FUSE && a.eps) go = ld4(dOut + ((size_t)u * d + c));
    const size_t row0 = (size_t)u * k * d + c;
    for (int h = 0; h < k; ++h) {
      const int e3 = __ldg(rp + min(h + 3, k));
      const unsigned xoff = (unsigned)(h * d) + c;
      float4 acc = sg.consume(e0, e1, xoff, h, nullptr, 0u);
      if (h + 1 < k) sg.prefetch(e1, e2, xoff + d);
      const size_t row = row0 + (size_t)h * d;
      if (EXTRA) {
        if (a.dinv) {
          const float w = __ldg(a.dinv + (size_t)u * Kp + h);
          acc.x *= w; acc.y *= w; acc.z *= w; acc.w *= w;
        }
        if (a.eps) {
          float4 dy;
          if (FUSE) {
            const float4 th = ld4(a.theta + h * d + c);
            dy = make_float4(th.x * go.x, th.y * go.y, th.z * go.z, th.w * go.w);
          } else {
            dy = ld4(dOut + row);
          }
          fma4(acc, self_c, dy);
        }
      }
      if (active) st4s(dX + row, acc);
      e0 = e1; e1 = e2; e2 = e3;
    }
  }
}

}  // namespace kp
