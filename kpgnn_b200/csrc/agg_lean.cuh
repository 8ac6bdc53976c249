// agg_lean.cuh -- third generation of the fused K-hop aggregation kernels (sm_100a).
//
// Why: ncu on the register-prefetch and cp.async-ring kernels (profiles/r1g, r1i) showed DRAM traffic == the
// algorithmic bytes but the SMs ISSUE-bound: 63-73 % issue-active at 208-233 warp instructions per (node,hop) row,
// of which only ~76 were the row's arithmetic; the rest was software-pipelining bookkeeping (ring slots, "rows
// already requested" state, window checks per entry, 64-bit addressing).  This version spends instructions only on
// the row itself:
//   * Blackwell packed fp32 (FADD2 / FMUL2 / FFMA2 on register pairs) for the accumulation, the GELU polynomial
//     and the theta-combine: 16 bytes of a row are two 64-bit registers from load to store;
//   * DRAM latency is taken off the instruction stream altogether: one lane per node issues
//     cp.async.bulk.prefetch.L2 for the X and P rows of the node its group will process `pf` iterations later
//     (3.3 KB per instruction), so every gather and every P read below is an L1/L2 hit;
//   * the node's k+1 row pointers are one coalesced load (lane h holds rowptr[v*K+h]) and its entry list one more
//     (lane i holds entry nbeg+i, already multiplied out to an X element offset and a table byte offset); both are
//     loaded one node ahead, so the hop loop starts with everything in registers and broadcasts with SHFL;
//   * the P row runs one hop ahead in registers; gathers are issued two at a time.
// Requirements beyond the fast path's (d % 4 == 0, d <= 128, 32-bit offsets): k + 1 <= G.
#pragma once
#include "agg_fast.cuh"

namespace kp {

typedef unsigned long long u64;

struct P4 {      // four consecutive floats as two packed pairs: lo = (x, y), hi = (z, w)
  u64 lo, hi;
};

__device__ __forceinline__ u64 pk2(float a, float b) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpk2(u64 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 splat2(float a) { return pk2(a, a); }
__device__ __forceinline__ P4 p4zero() { return P4{0ull, 0ull}; }
__device__ __forceinline__ P4 add4p(const P4& a, const P4& b) { return P4{add2(a.lo, b.lo), add2(a.hi, b.hi)}; }

// gathered rows: read-only path, L1-allocating (a row is re-gathered by ~2.5 destination rows of the same graph)
__device__ __forceinline__ P4 ldg4p(const float* p) {
  P4 v;
  asm("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "l"(p));
  return v;
}
// rows touched exactly once (P, dOut): no L1 allocation
__device__ __forceinline__ P4 ldg4p_stream(const float* p) {
  P4 v;
  asm("ld.global.nc.L1::no_allocate.v2.b64 {%0, %1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "l"(p));
  return v;
}
__device__ __forceinline__ P4 lds4p(unsigned addr) {
  P4 v;
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts4p(unsigned addr, const P4& v) {
  asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(v.lo), "l"(v.hi) : "memory");
}
__device__ __forceinline__ void stg4p_stream(float* p, const P4& v) {
  asm volatile("st.global.cs.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(v.lo), "l"(v.hi) : "memory");
}
// coherent load (no read-only / L1-texture path): for memory this kernel or an overlapping one also writes
__device__ __forceinline__ P4 ld4p_coherent(const float* p) {
  P4 v;
  asm volatile("ld.global.v2.b64 {%0, %1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stg4p(float* p, const P4& v) {
  asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(v.lo), "l"(v.hi) : "memory");
}
// asynchronous L2 prefetch of `bytes` (multiple of 16) starting at a 16-byte aligned global address
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// exact-erf GELU on a packed pair; same Abramowitz-Stegun 7.1.26 erfc form as agg_common.cuh (|err| < 5e-7), with
// the polynomial, the exponent argument and the products on the packed pipe.  The polynomial coefficients carry
// the minus sign, so  gelu(x) = max(x,0) + |x| * (-half_erfc(|x|)).
#define KP_GELU_C0 (0.3275911f * 0.70710678118654752440f)
#define KP_GELU_S 0.84932180028801904272f /* sqrt(0.5 * log2(e)):  exp(-x^2/2) = 2^-(S x)^2 */
__device__ __forceinline__ void gelu_pair_parts(u64 x, u64& neg_half_erfc, u64& gauss) {
  float x0, x1;
  unpk2(x, x0, x1);
  const float t0 = rcp_approx(fmaf(KP_GELU_C0, fabsf(x0), 1.0f));
  const float t1 = rcp_approx(fmaf(KP_GELU_C0, fabsf(x1), 1.0f));
  const u64 t = pk2(t0, t1);
  u64 p = fma2(t, splat2(-0.5f * 1.061405429f), splat2(-0.5f * -1.453152027f));
  p = fma2(t, p, splat2(-0.5f * 1.421413741f));
  p = fma2(t, p, splat2(-0.5f * -0.284496736f));
  p = fma2(t, p, splat2(-0.5f * 0.254829592f));
  const u64 y = mul2(x, splat2(KP_GELU_S));
  const u64 q = mul2(y, y);
  float q0, q1;
  unpk2(q, q0, q1);
  gauss = pk2(ex2_approx(-q0), ex2_approx(-q1));
  neg_half_erfc = mul2(mul2(p, t), gauss);
}
template <int ACT>
__device__ __forceinline__ u64 act_fwd2(u64 x) {
  if (ACT == KP_ACT_GELU) {
    u64 nh, g;
    gelu_pair_parts(x, nh, g);
    float x0, x1, h0, h1;
    unpk2(x, x0, x1);
    unpk2(nh, h0, h1);
    return pk2(fmaf(fabsf(x0), h0, fmaxf(x0, 0.f)), fmaf(fabsf(x1), h1, fmaxf(x1, 0.f)));
  }
  if (ACT == KP_ACT_RELU) {
    float x0, x1;
    unpk2(x, x0, x1);
    return pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
  }
  return x;
}
// d act / d x on a packed pair
template <int ACT>
__device__ __forceinline__ u64 act_bwd2(u64 x) {
  if (ACT == KP_ACT_GELU) {
    u64 nh, g;
    gelu_pair_parts(x, nh, g);
    float x0, x1, h0, h1;
    unpk2(x, x0, x1);
    unpk2(nh, h0, h1);
    const u64 cdf = pk2(x0 > 0.f ? 1.0f + h0 : -h0, x1 > 0.f ? 1.0f + h1 : -h1);
    return fma2(mul2(x, splat2(0.39894228040143267794f)), g, cdf);
  }
  if (ACT == KP_ACT_RELU) {
    float x0, x1;
    unpk2(x, x0, x1);
    return pk2(x0 > 0.f ? 1.f : 0.f, x1 > 0.f ? 1.f : 0.f);
  }
  return splat2(1.f);
}

__device__ __forceinline__ void l2_prefetch_line(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ uint2 lds2_sh(unsigned addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds1_sh(unsigned addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts2_sh(unsigned addr, unsigned x, unsigned y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts1_sh(unsigned addr, int x) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(x) : "memory");
}
template <int G>
__device__ __forceinline__ void group_sync(unsigned gm) {
  __syncwarp(gm);
}
template <typename T>
__device__ __forceinline__ T* opaque_ptr(T* p) {      // keeps a 64-bit base in registers: address = IMAD.WIDE.U32
  asm volatile("" : "+l"(p));
  return p;
}
__device__ __forceinline__ const float* at_elem(const float* base, unsigned elem) {
  return reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + (size_t)elem * 4u);
}

#ifndef KP_LEAN_MINB
#define KP_LEAN_MINB 4
#endif
#ifndef KP_LEAN_PIPE
#define KP_LEAN_PIPE 0
#endif

// Sum of NE consecutive entries of the group's shared-memory entry window starting at byte address `ent`:
// the NE gathers are issued back to back, the table rows (shared memory) are added while they fly.
template <int NE, int TAB>
__device__ __forceinline__ void lean_gather(P4& z, unsigned ent, const float* Xh, unsigned c4) {
  uint2 en[NE];
  P4 x[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) en[i] = lds2_sh(ent + 8u * i);
#pragma unroll
  for (int i = 0; i < NE; ++i) x[i] = ldg4p(at_elem(Xh, en[i].x));
  if (TAB == TAB_SMEM) {
#pragma unroll
    for (int i = 0; i < NE; ++i) z = add4p(z, lds4p(en[i].y + c4));
  }
#pragma unroll
  for (int i = 0; i < NE; ++i) z = add4p(z, x[i]);
}

// Entries G .. kLeanWin-1 of the node being published (the first G came through the register prefetch): coalesced
// loads straight from the plan into the group's window.  Uniform per group; a no-op for nodes with <= G entries.
template <int G, int TAB>
__device__ __forceinline__ void lean_publish_rest(const kp_agg_desc& a, unsigned win_sh, int nbeg, int nend, int e1,
                                                  int lane, unsigned xs, unsigned d4, unsigned tab0_sh,
                                                  unsigned tabk_sh) {
  const int cnt = min(nend - nbeg, kLeanWin);
  for (int j = G + lane; j < cnt; j += G) {
    const int cj = __ldg(a.col + nbeg + j);
    int aj = 0;
    if (TAB != TAB_NONE) aj = (int)__ldg(a.attr16 + nbeg + j);
    sts2_sh(win_sh + 8u * (unsigned)j, (unsigned)cj * xs, (nbeg + j < e1 ? tab0_sh : tabk_sh) + (unsigned)aj * d4);
  }
}

// ------------------------------------------------------------------------------------------------------------
// forward.  Everything except the KP-GCN per-entry norm (dinv) and tables too large for shared memory, which stay
// on the kernels of agg_fast.cuh.
// ------------------------------------------------------------------------------------------------------------
template <int G, int ACT, bool FUSE, int TAB, bool EXTRA>
__global__ void __launch_bounds__(1024, 1)      // <= 64 registers; launched with 256..1024 threads
agg_fwd_lean_kernel(const FastArgs fa, float* __restrict__ out, unsigned pf_x_lines, unsigned pf_p_lines, int pf_dist,
                    unsigned pf_bulk_bytes) {
  extern __shared__ __align__(16) float sm[];
  const kp_agg_desc& a = fa.d;
  // Programmatic dependent launch: the tables / theta (parameters and step-head products) are staged while the
  // preceding kernel (the dense block) is still running; X and P are touched behind kp_pdl_wait().  Dependents are
  // released only after this CTA's own wait, so a dependent's prologue never overlaps more than one kernel back and
  // never competes with CTAs of this grid that are not resident yet.
  const int staged = stage_tables<TAB, FUSE>(a, sm);
  kp_pdl_wait();
  kp_pdl_trigger();
  const int d = a.d, k = a.k, Kp = a.Kplan, N = a.N;
  const unsigned xs = fa.xs, d4 = (unsigned)d * 4u;
  const int lane = threadIdx.x & (G - 1);
  const bool active = lane * 4 < d;
  const unsigned c = (unsigned)min(lane * 4, d - 4);            // idle lanes shadow the last chunk
  const int gpb = blockDim.x / G;
  const int gib = threadIdx.x / G;
  const unsigned gm = group_mask<G>();
  const unsigned sm_base = sh_addr(sm);
  // table base addresses (hop 0 uses T0, hops >= 1 use Tk); window entries are shared by the group's lanes, so the
  // lane's own column offset c4 is added at the lookup
  const unsigned c4 = c * 4u;
  const unsigned tab0_sh = sm_base;
  const unsigned tabk_sh = tab0_sh + ((TAB == TAB_SMEM) ? (unsigned)(a.rows0 * d) * 4u : 0u);
  const unsigned theta_sh = sm_base + ((TAB == TAB_SMEM) ? (unsigned)((a.rows0 + a.rowsk) * d) : 0u) * 4u + c * 4u;
  // per-group scratch behind the staged tables: entry window [kLeanWin] x {X element offset, table byte address} and
  // the node's row pointers [G]
  const unsigned win_sh = sm_base + (unsigned)staged * 4u + (unsigned)gib * lean_group_scratch_bytes(G);
  const unsigned rp_sh = win_sh + 8u * kLeanWin;
  const float* Xc = opaque_ptr(a.X + c);
  const bool hasP = a.P != nullptr;
  const float* Pc = opaque_ptr(hasP ? a.P + c : a.X + c);
  float self_c = 0.f;
  if (EXTRA && a.eps) self_c = 1.f + __ldg(a.eps);

  const int vstride = gridDim.x * gpb;
  int v = blockIdx.x * gpb + gib;
  if (v >= N) return;
  // L2 prefetch: lane l asks for the l-th 128-byte line of a node's X and P rows
  auto prefetch_node = [&](long long vp) {
    if (pf_bulk_bytes) {                     // experiment: one TMA-engine prefetch per node instead of per-line hints
      if (vp < N && lane == 0) {
        l2_prefetch_bulk(a.X + (size_t)vp * xs, pf_bulk_bytes);
        if (hasP) l2_prefetch_bulk(a.P + (size_t)vp * fa.ps, pf_bulk_bytes);
      }
      return;
    }
    if (vp < N) {
      if ((unsigned)lane < pf_x_lines) l2_prefetch_line(reinterpret_cast<const char*>(a.X + (size_t)vp * xs) + lane * 128);
      if ((unsigned)lane < pf_p_lines) l2_prefetch_line(reinterpret_cast<const char*>(a.P + (size_t)vp * fa.ps) + lane * 128);
    }
  };
  for (int i = 0; i < pf_dist; ++i) prefetch_node((long long)v + (long long)i * vstride);

  // software pipeline over nodes: row pointers two nodes ahead, entries one node ahead (registers), current node in
  // the group's shared-memory window
  int rpn = (lane <= k) ? __ldg(a.rowptr + (size_t)v * Kp + lane) : 0;            // "next" = the first node
  int ncol = 0, nattr = 0;
  {
    const int nb = __shfl_sync(gm, rpn, 0, G), ne = __shfl_sync(gm, rpn, k, G);
    if (nb + lane < ne) {
      ncol = __ldg(a.col + nb + lane);
      if (TAB != TAB_NONE) nattr = (int)__ldg(a.attr16 + nb + lane);
    }
  }
  int vn = v;
  int rpnn = 0;
  {
    const int v2 = v + vstride;
    if (v2 < N && lane <= k) rpnn = __ldg(a.rowptr + (size_t)v2 * Kp + lane);
  }
  while (true) {
    // ---- publish node vn (registers -> the group's window), then request the node after it
    v = vn;
    const int nbeg = __shfl_sync(gm, rpn, 0, G), nend = __shfl_sync(gm, rpn, k, G);
    const int e1 = __shfl_sync(gm, rpn, 1, G);
    group_sync<G>(gm);                                           // everyone is done reading the previous window
    sts1_sh(rp_sh + 4u * lane, rpn - nbeg);                      // window-relative row pointers
    {
      const unsigned xo = (unsigned)ncol * xs;
      const unsigned ta = (nbeg + lane < e1 ? tab0_sh : tabk_sh) + (unsigned)nattr * d4;
      sts2_sh(win_sh + 8u * lane, xo, ta);
    }
    lean_publish_rest<G, TAB>(a, win_sh, nbeg, nend, e1, lane, xs, d4, tab0_sh, tabk_sh);
    group_sync<G>(gm);
    const bool big = (nend - nbeg) > kLeanWin;                   // entry list longer than the window: slow path
    vn = v + vstride;
    rpn = rpnn;
    ncol = 0; nattr = 0;
    if (vn < N) {
      const int nb = __shfl_sync(gm, rpn, 0, G), ne = __shfl_sync(gm, rpn, k, G);
      if (nb + lane < ne) {
        ncol = __ldg(a.col + nb + lane);
        if (TAB != TAB_NONE) nattr = (int)__ldg(a.attr16 + nb + lane);
      }
      const int v2 = vn + vstride;
      rpnn = (v2 < N && lane <= k) ? __ldg(a.rowptr + (size_t)v2 * Kp + lane) : 0;
    }
    prefetch_node((long long)v + (long long)pf_dist * vstride);

    // ---- this node
    const float* Xh = Xc;                                        // + h * xh per hop
    const float* Pv = Pc + (size_t)v * fa.ps;
    float* outv = out + (FUSE ? (size_t)v * d : (size_t)v * (fa.os ? fa.os : (unsigned)(k * d))) + c;
    P4 o = p4zero();
    unsigned ent = win_sh;                                       // byte address of the segment's first entry
    unsigned th = theta_sh;
#if KP_LEAN_PIPE
    // the first two gathered rows of every hop are requested one hop early, so their latency hides behind the
    // previous hop's activation math; rows 3.. of a segment are gathered on demand
    P4 xa = p4zero(), xb = p4zero();
    unsigned ta = 0, tb = 0;
    int e = lds1_sh(rp_sh + 4u);
    int n = e;
    if (!big) {
      if (n >= 1) {
        const uint2 en = lds2_sh(ent);
        xa = ldg4p(at_elem(Xh, en.x));
        ta = en.y;
      }
      if (n >= 2) {
        const uint2 en = lds2_sh(ent + 8u);
        xb = ldg4p(at_elem(Xh, en.x));
        tb = en.y;
      }
    }
    int b = 0;
    for (int h = 0; h < k; ++h) {
      const int e2 = (h + 1 < k) ? lds1_sh(rp_sh + 4u * (h + 2)) : e;
      P4 p = p4zero();
      if (hasP) p = ldg4p_stream(Pv);
      P4 z = p4zero();
      const int nn = e2 - e;
      if (!big) {
        if (n >= 1) {
          z = xa;
          if (TAB == TAB_SMEM) z = add4p(z, lds4p(ta + c4));
        }
        if (n >= 2) {
          if (TAB == TAB_SMEM) xb = add4p(xb, lds4p(tb + c4));
          z = add4p(z, xb);
        }
        if (n > 2) {
          int m = n - 2;
          unsigned er = ent + 16u;
          while (m >= 4) {
            lean_gather<4, TAB>(z, er, Xh, c4);
            er += 32u;
            m -= 4;
          }
          if (m & 2) {
            lean_gather<2, TAB>(z, er, Xh, c4);
            er += 16u;
          }
          if (m & 1) lean_gather<1, TAB>(z, er, Xh, c4);
        }
        ent += 8u * (unsigned)n;
        const float* Xn = Xh + fa.xh;
        if (nn >= 1) {
          const uint2 en = lds2_sh(ent);
          xa = ldg4p(at_elem(Xn, en.x));
          ta = en.y;
        }
        if (nn >= 2) {
          const uint2 en = lds2_sh(ent + 8u);
          xb = ldg4p(at_elem(Xn, en.x));
          tb = en.y;
        }
      } else {
        for (int j = nbeg + b; j < nbeg + e; ++j) {
          const int cj = __ldg(a.col + j);
          P4 x = ldg4p(at_elem(Xh, (unsigned)cj * xs));
          if (TAB == TAB_SMEM) x = add4p(x, lds4p((h == 0 ? tab0_sh : tabk_sh) + c4 + (unsigned)__ldg(a.attr16 + j) * d4));
          z = add4p(z, x);
        }
      }
      b = e;
      e = e2;
      n = nn;
#else
    int b = 0;
    for (int h = 0; h < k; ++h) {
      const int e = lds1_sh(rp_sh + 4u * (h + 1));
      P4 p = p4zero();
      if (hasP) p = ldg4p_stream(Pv);
      P4 z = p4zero();
      if (!big) {
        int n = e - b;
        while (n >= 4) {
          lean_gather<4, TAB>(z, ent, Xh, c4);
          ent += 32u;
          n -= 4;
        }
        if (n & 2) {
          lean_gather<2, TAB>(z, ent, Xh, c4);
          ent += 16u;
        }
        if (n & 1) {
          lean_gather<1, TAB>(z, ent, Xh, c4);
          ent += 8u;
        }
      } else {
        // rare: more than G entries on one node -- walk the plan arrays directly
        for (int j = nbeg + b; j < nbeg + e; ++j) {
          const int cj = __ldg(a.col + j);
          P4 x = ldg4p(at_elem(Xh, (unsigned)cj * xs));
          if (TAB == TAB_SMEM) x = add4p(x, lds4p((h == 0 ? tab0_sh : tabk_sh) + c4 + (unsigned)__ldg(a.attr16 + j) * d4));
          z = add4p(z, x);
        }
      }
      b = e;
#endif
      if (EXTRA) {
        const u64 s = splat2(fast_row_scale<EXTRA>(a, v, h));
        z.lo = mul2(z.lo, s); z.hi = mul2(z.hi, s);
      }
      z.lo = act_fwd2<ACT>(z.lo);
      z.hi = act_fwd2<ACT>(z.hi);
      if (EXTRA && a.eps) {
        const P4 xv = ldg4p(at_elem(Xh, (unsigned)v * xs));
        const u64 sc = splat2(self_c);
        z.lo = fma2(sc, xv.lo, z.lo); z.hi = fma2(sc, xv.hi, z.hi);
      }
      z = add4p(z, p);
      if (FUSE) {
        const P4 t = lds4p(th);
        o.lo = fma2(t.lo, z.lo, o.lo); o.hi = fma2(t.hi, z.hi, o.hi);
        th += d4;
      } else if (active) {
        float* po = outv + (size_t)h * (fa.oh ? fa.oh : (unsigned)d);
        if (fa.oacc) z = add4p(z, ld4p_coherent(po));     // read-modify-write of the caller's buffer: never through .nc
        stg4p_stream(po, z);
      }
      Xh += fa.xh;
      Pv += fa.ph;
    }
    if (FUSE && active) stg4p_stream(outv, o);
    if (vn >= N) break;
  }
}

// gelu(x) and gelu'(x) of a packed pair from one evaluation of the erfc form
__device__ __forceinline__ void gelu_fwd_bwd2(u64 x, u64& f, u64& df) {
  u64 nh, g;
  gelu_pair_parts(x, nh, g);
  float x0, x1, h0, h1;
  unpk2(x, x0, x1);
  unpk2(nh, h0, h1);
  f = pk2(fmaf(fabsf(x0), h0, fmaxf(x0, 0.f)), fmaf(fabsf(x1), h1, fmaxf(x1, 0.f)));
  const u64 cdf = pk2(x0 > 0.f ? 1.0f + h0 : -h0, x1 > 0.f ? 1.0f + h1 : -h1);
  df = fma2(mul2(x, splat2(0.39894228040143267794f)), g, cdf);
}
template <int ACT>
__device__ __forceinline__ void act_fwd_bwd2(u64 x, u64& f, u64& df) {
  if (ACT == KP_ACT_GELU) {
    gelu_fwd_bwd2(x, f, df);
  } else {
    f = act_fwd2<ACT>(x);
    df = act_bwd2<ACT>(x);
  }
}

// ------------------------------------------------------------------------------------------------------------
// B1 (backward by destination row), lean version for the layers without self term / norm / mean:
//   pre = sum_j (X[col_j,h] + T_h[attr_j])                         (recomputed, never stored by the forward)
//   dy  = fuse ? theta[h] * dOut[v] : dOut[v,h]       -> dP
//   Gs  = dy * act'(pre)                               -> workspace, consumed by B2 (dX) and B3 (dT0/dTk)
//   dtheta[h] += dOut[v] * (act(pre) + P[v,h])         (per-group accumulators in shared memory, fixed-order
//                                                       reduction per CTA -> dtheta_part[blockIdx.x])
// Same node pipeline, entry window and packed arithmetic as agg_fwd_lean_kernel.
// ------------------------------------------------------------------------------------------------------------
template <int G, int ACT, bool FUSE, int TAB>
__global__ void __launch_bounds__(1024, 1)
agg_bwd_dst_lean_kernel(const FastArgs fa, const float* __restrict__ dOut, float* __restrict__ Gs,
                        float* __restrict__ dP, float* __restrict__ dtheta_part) {
  extern __shared__ __align__(16) float sm[];
  const kp_agg_desc& a = fa.d;
  const int staged = stage_tables<TAB, FUSE>(a, sm);
  const int d = a.d, k = a.k, Kp = a.Kplan, N = a.N;
  const unsigned xs = fa.xs, d4 = (unsigned)d * 4u;
  const int lane = threadIdx.x & (G - 1);
  const bool active = lane * 4 < d;
  const unsigned c = (unsigned)min(lane * 4, d - 4);
  const int gpb = blockDim.x / G;
  const int gib = threadIdx.x / G;
  const unsigned gm = group_mask<G>();
  const unsigned sm_base = sh_addr(sm);
  const unsigned c4 = c * 4u;
  const unsigned tab0_sh = sm_base;
  const unsigned tabk_sh = tab0_sh + ((TAB == TAB_SMEM) ? (unsigned)(a.rows0 * d) * 4u : 0u);
  const unsigned theta_sh = sm_base + ((TAB == TAB_SMEM) ? (unsigned)((a.rows0 + a.rowsk) * d) : 0u) * 4u + c4;
  const unsigned win_sh = sm_base + (unsigned)staged * 4u + (unsigned)gib * lean_group_scratch_bytes(G);
  const unsigned rp_sh = win_sh + 8u * kLeanWin;
  // dtheta accumulators: [gpb][k][4G] floats behind the windows; lane l of group g owns columns 4l..4l+3 of copy g
  constexpr unsigned dpad = 4u * G;
  const bool need_z = FUSE && dtheta_part != nullptr;
  float* acc_all = sm + staged + gpb * (int)(lean_group_scratch_bytes(G) / 4u);
  const unsigned acc_sh = sm_base + (unsigned)staged * 4u + (unsigned)gpb * lean_group_scratch_bytes(G) + (unsigned)(gib * k) * dpad * 4u + (unsigned)lane * 16u;
  if (need_z) {
    for (int i = threadIdx.x * 4; i < gpb * k * (int)dpad; i += blockDim.x * 4) st4(acc_all + i, make_float4(0.f, 0.f, 0.f, 0.f));
    __syncthreads();
  }
  kp_pdl_wait();         // prologue (tables, accumulators) overlapped the preceding kernel; dOut / X / P from here on
  kp_pdl_trigger();
  const float* Xc = opaque_ptr(a.X + c);
  const bool hasP = need_z && a.P != nullptr;
  const float* Pc = opaque_ptr(hasP ? a.P + c : a.X + c);

  const int vstride = gridDim.x * gpb;
  int v = blockIdx.x * gpb + gib;
  if (v < N) {
    int rpn = (lane <= k) ? __ldg(a.rowptr + (size_t)v * Kp + lane) : 0;
    int ncol = 0, nattr = 0;
    {
      const int nb = __shfl_sync(gm, rpn, 0, G), ne = __shfl_sync(gm, rpn, k, G);
      if (nb + lane < ne) {
        ncol = __ldg(a.col + nb + lane);
        if (TAB != TAB_NONE) nattr = (int)__ldg(a.attr16 + nb + lane);
      }
    }
    int vn = v;
    int rpnn = 0;
    {
      const int v2 = v + vstride;
      if (v2 < N && lane <= k) rpnn = __ldg(a.rowptr + (size_t)v2 * Kp + lane);
    }
    while (true) {
      v = vn;
      const int nbeg = __shfl_sync(gm, rpn, 0, G), nend = __shfl_sync(gm, rpn, k, G);
      const int e1 = __shfl_sync(gm, rpn, 1, G);
      group_sync<G>(gm);
      sts1_sh(rp_sh + 4u * lane, rpn - nbeg);
      {
        const unsigned xo = (unsigned)ncol * xs;
        const unsigned ta = (nbeg + lane < e1 ? tab0_sh : tabk_sh) + (unsigned)nattr * d4;
        sts2_sh(win_sh + 8u * lane, xo, ta);
      }
      lean_publish_rest<G, TAB>(a, win_sh, nbeg, nend, e1, lane, xs, d4, tab0_sh, tabk_sh);
      group_sync<G>(gm);
      const bool big = (nend - nbeg) > kLeanWin;
      vn = v + vstride;
      rpn = rpnn;
      ncol = 0; nattr = 0;
      if (vn < N) {
        const int nb = __shfl_sync(gm, rpn, 0, G), ne = __shfl_sync(gm, rpn, k, G);
        if (nb + lane < ne) {
          ncol = __ldg(a.col + nb + lane);
          if (TAB != TAB_NONE) nattr = (int)__ldg(a.attr16 + nb + lane);
        }
        const int v2 = vn + vstride;
        rpnn = (v2 < N && lane <= k) ? __ldg(a.rowptr + (size_t)v2 * Kp + lane) : 0;
      }

      const float* Xh = Xc;
      const float* Pv = Pc + (size_t)v * fa.ps;
      const size_t row0 = (size_t)v * k * d + c;
      P4 go = p4zero();
      if (FUSE) go = ldg4p_stream(dOut + ((size_t)v * d + c));
      unsigned ent = win_sh;
      unsigned th = theta_sh;
      unsigned ac = acc_sh;
      int b = 0;
      for (int h = 0; h < k; ++h) {
        const int e = lds1_sh(rp_sh + 4u * (h + 1));
        const size_t row = row0 + (size_t)h * d;
        P4 p = p4zero();
        if (hasP) p = ldg4p_stream(Pv);
        P4 dy;
        if (FUSE) {
          const P4 t = lds4p(th);
          dy.lo = mul2(t.lo, go.lo); dy.hi = mul2(t.hi, go.hi);
          th += d4;
        } else {
          dy = ldg4p_stream(dOut + row);
        }
        if (dP && active) stg4p_stream(dP + row, dy);
        P4 z = p4zero();
        if (!big) {
          int n = e - b;
          while (n >= 4) {
            lean_gather<4, TAB>(z, ent, Xh, c4);
            ent += 32u;
            n -= 4;
          }
          if (n & 2) {
            lean_gather<2, TAB>(z, ent, Xh, c4);
            ent += 16u;
          }
          if (n & 1) {
            lean_gather<1, TAB>(z, ent, Xh, c4);
            ent += 8u;
          }
        } else {
          for (int j = nbeg + b; j < nbeg + e; ++j) {
            const int cj = __ldg(a.col + j);
            P4 x = ldg4p(at_elem(Xh, (unsigned)cj * xs));
            if (TAB == TAB_SMEM) x = add4p(x, lds4p((h == 0 ? tab0_sh : tabk_sh) + c4 + (unsigned)__ldg(a.attr16 + j) * d4));
            z = add4p(z, x);
          }
        }
        b = e;
        P4 f, df;
        act_fwd_bwd2<ACT>(z.lo, f.lo, df.lo);
        act_fwd_bwd2<ACT>(z.hi, f.hi, df.hi);
        if (need_z) {
          f = add4p(f, p);
          P4 t = lds4p(ac);
          t.lo = fma2(go.lo, f.lo, t.lo); t.hi = fma2(go.hi, f.hi, t.hi);
          sts4p(ac, t);
          ac += dpad * 4u;
        }
        if (Gs && active) {
          P4 g;
          g.lo = mul2(dy.lo, df.lo); g.hi = mul2(dy.hi, df.hi);
          stg4p(Gs + row, g);
        }
        Xh += fa.xh;
        Pv += fa.ph;
      }
      if (vn >= N) break;
    }
  }
  if (dtheta_part) {
    __syncthreads();
    for (int i = threadIdx.x; i < k * d; i += blockDim.x) {   // fixed-order reduction over the CTA's groups
      const int h = i / d, cc = i - h * d;
      float s = 0.f;
#pragma unroll 8
      for (int g = 0; g < gpb; ++g) s += acc_all[((size_t)g * k + h) * dpad + cc];      // loads batched, adds in order
      dtheta_part[(size_t)blockIdx.x * k * d + i] = s;
    }
  }
}

}  // namespace kp
