// agg_fast.cuh -- instruction-lean float4 kernels for the K-hop aggregation (the common case: d % 4 == 0,
// d <= 128, 16-byte aligned operands, int32 strides).  The generic kernels in agg.cu cover everything else.
//
// What the first ncu capture of the generic forward kernel showed (profiles/r1a_agg_fwd_generic.md): DRAM
// traffic already equals the algorithmic bytes, but the kernel was ISSUE-bound (75 % issue-active, 648 M warp
// instructions for 1.5 M (node,hop) rows): erff GELU, 64-bit address arithmetic, predicated 4-way unrolling
// and constant-bank reloads.  This version:
//   * a group of G lanes (G = 4/8/16/32, 4*G >= d) owns one destination node; lane l owns channels 4l..4l+3;
//   * the hop segment's (col, attr) entries are loaded ONCE, coalesced, by the group's lanes and broadcast with
//     width-G shuffles; the next hop's entries, its row pointer and the P row are prefetched before the
//     current hop's gathers are consumed (software pipelining across hops);
//   * embedding tables and theta are staged in shared memory once per CTA (persistent grid), so a table lookup
//     is one LDS.128 with 32-bit addressing;
//   * gathers are issued two at a time; X/Gs rows use the default (L1-allocating) path because a row is reused
//     by ~2.5 destination rows of the same small graph, P/dOut/outputs use streaming loads/stores;
//   * GELU through the erfc form in agg_common.cuh (2 MUFU + 12 FP32 ops).
#pragma once
#include "agg_common.cuh"

namespace kp {

struct FastArgs {
  kp_agg_desc d;
  int xs, xh, ps, ph;   // element strides of X and P (validated to fit int32 by the host)
  int tab_floats;       // floats of embedding table staged in smem (0: read tables from global)
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4s(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 lds4_sh(unsigned addr) {   // explicit LDS.128 with a 32-bit shared address
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned sh_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st4s(float* p, const float4& v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if (G == 32) return 0xffffffffu;
  return ((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1));
}

// Copies the embedding tables (and theta) into shared memory; returns pointers.
template <bool TSMEM, bool NEED_THETA>
__device__ __forceinline__ void stage_tables(const FastArgs& fa, float* sm, const float*& T0, const float*& Tk,
                                             const float*& theta) {
  const kp_agg_desc& a = fa.d;
  T0 = a.T0;
  Tk = a.Tk;
  theta = a.theta;
  int off = 0;
  if (TSMEM && a.T0) {
    const int n0 = a.rows0 * a.d, nk = a.rowsk * a.d;
    for (int i = threadIdx.x * 4; i < n0; i += blockDim.x * 4) st4(sm + i, ld4(a.T0 + i));
    for (int i = threadIdx.x * 4; i < nk; i += blockDim.x * 4) st4(sm + n0 + i, ld4(a.Tk + i));
    T0 = sm;
    Tk = sm + n0;
    off = n0 + nk;
  }
  if (NEED_THETA && a.theta) {
    const int nt = a.k * a.d;
    for (int i = threadIdx.x * 4; i < nt; i += blockDim.x * 4) st4(sm + off + i, ld4(a.theta + i));
    theta = sm + off;
  }
  __syncthreads();
}

// Walks one destination node's hop segments with the entry prefetch described above.
template <int G>
struct HopCursor {
  const int* rp;        // row pointers of this node
  const int* col;
  const uint16_t* attr;
  int b, e;             // current segment
  int pc, pa;           // this lane's prefetched entry of the current window
  int nb, ne, npc, npa;
  int lane;

  __device__ __forceinline__ void load_window(int start, int end, int& c, int& a) const {
    c = 0;
    a = 0;
    if (start + lane < end) {
      c = __ldg(col + start + lane);
      if (attr) a = (int)__ldg(attr + start + lane);
    }
  }
  __device__ __forceinline__ void begin(const int* rowptr_row, const int* col_, const uint16_t* attr_, int lane_) {
    rp = rowptr_row;
    col = col_;
    attr = attr_;
    lane = lane_;
    b = __ldg(rp);
    e = __ldg(rp + 1);
    load_window(b, e, pc, pa);
  }
  // issue the loads for hop h+1 (call at the top of hop h)
  __device__ __forceinline__ void prefetch_next(int h, int k) {
    nb = e;
    ne = e;
    npc = 0;
    npa = 0;
    if (h + 1 < k) {
      ne = __ldg(rp + h + 2);
      load_window(nb, ne, npc, npa);
    }
  }
  __device__ __forceinline__ void advance() {
    b = nb;
    e = ne;
    pc = npc;
    pa = npa;
  }
};

// acc = sum over the current segment of w_j * (X[col_j] + T[attr_j]); all lanes of the group must call.
template <int G, bool HAS_T, bool TSMEM>
__device__ __forceinline__ float4 gather_segment(HopCursor<G>& cur, unsigned gm, bool active, const float* Xh, int xs,
                                                 const float* Th, unsigned Tsh, int d, const float* dinv_h, int Kp) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int j = cur.b;
  int pc = cur.pc, pa = cur.pa;
  while (true) {
    const int cnt = min(G, cur.e - j);
    int q = 0;
    for (; q + 1 < cnt; q += 2) {
      const int u0 = __shfl_sync(gm, pc, q, G), u1 = __shfl_sync(gm, pc, q + 1, G);
      int a0 = 0, a1 = 0;
      if (HAS_T) {
        a0 = __shfl_sync(gm, pa, q, G);
        a1 = __shfl_sync(gm, pa, q + 1, G);
      }
      if (active) {
        const float4 x0 = ld4(Xh + (long long)u0 * xs);
        const float4 x1 = ld4(Xh + (long long)u1 * xs);
        float4 s0 = x0, s1 = x1;
        if (HAS_T) {
          const float4 t0 = TSMEM ? lds4_sh(Tsh + a0 * d * 4) : ld4(Th + a0 * d);
          const float4 t1 = TSMEM ? lds4_sh(Tsh + a1 * d * 4) : ld4(Th + a1 * d);
          s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
          s1.x += t1.x; s1.y += t1.y; s1.z += t1.z; s1.w += t1.w;
        }
        if (dinv_h) {
          const float w0 = __ldg(dinv_h + (long long)u0 * Kp), w1 = __ldg(dinv_h + (long long)u1 * Kp);
          acc.x = fmaf(w0, s0.x, acc.x); acc.y = fmaf(w0, s0.y, acc.y);
          acc.z = fmaf(w0, s0.z, acc.z); acc.w = fmaf(w0, s0.w, acc.w);
          acc.x = fmaf(w1, s1.x, acc.x); acc.y = fmaf(w1, s1.y, acc.y);
          acc.z = fmaf(w1, s1.z, acc.z); acc.w = fmaf(w1, s1.w, acc.w);
        } else {
          acc.x += s0.x; acc.y += s0.y; acc.z += s0.z; acc.w += s0.w;
          acc.x += s1.x; acc.y += s1.y; acc.z += s1.z; acc.w += s1.w;
        }
      }
    }
    if (q < cnt) {
      const int u0 = __shfl_sync(gm, pc, q, G);
      int a0 = 0;
      if (HAS_T) a0 = __shfl_sync(gm, pa, q, G);
      if (active) {
        float4 s0 = ld4(Xh + (long long)u0 * xs);
        if (HAS_T) {
          const float4 t0 = TSMEM ? lds4_sh(Tsh + a0 * d * 4) : ld4(Th + a0 * d);
          s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
        }
        if (dinv_h) {
          const float w0 = __ldg(dinv_h + (long long)u0 * Kp);
          acc.x = fmaf(w0, s0.x, acc.x); acc.y = fmaf(w0, s0.y, acc.y);
          acc.z = fmaf(w0, s0.z, acc.z); acc.w = fmaf(w0, s0.w, acc.w);
        } else {
          acc.x += s0.x; acc.y += s0.y; acc.z += s0.z; acc.w += s0.w;
        }
      }
    }
    j += cnt > 0 ? cnt : 0;
    if (j >= cur.e) break;
    cur.load_window(j, cur.e, pc, pa);     // segments longer than G entries
  }
  return acc;
}

__device__ __forceinline__ float fast_row_scale(const kp_agg_desc& a, int v, int h) {
  float s = 1.f;
  if (a.dinv) s *= __ldg(a.dinv + (long long)v * a.Kplan + h);
  if (a.indeg) s *= 1.f / (float)max(__ldg(a.indeg + v), 1);
  return s;
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int G, int ACT, bool FUSE, bool TSMEM>
__global__ void __launch_bounds__(256, 4) agg_fwd_fast_kernel(const FastArgs fa, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const kp_agg_desc& a = fa.d;
  const float *T0, *Tk, *theta;
  stage_tables<TSMEM, FUSE>(fa, sm, T0, Tk, theta);
  const int d = a.d, k = a.k, Kp = a.Kplan, xs = fa.xs, xh = fa.xh;
  const int lane = threadIdx.x & (G - 1);
  const int c = lane * 4;
  const bool active = c < d;
  const unsigned gm = group_mask<G>();
  constexpr int gpb = 256 / G;
  const int gib = threadIdx.x / G;
  const float self_c = a.eps ? 1.f + __ldg(a.eps) : 0.f;
  const bool has_t = a.T0 != nullptr;
  const unsigned theta_sh = FUSE ? sh_addr(theta) : 0u;
  for (int v = blockIdx.x * gpb + gib; v < a.N; v += gridDim.x * gpb) {
    HopCursor<G> cur;
    cur.begin(a.rowptr + (long long)v * Kp, a.col, has_t ? a.attr16 : nullptr, lane);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* Xc = a.X + c;
    const float* Pv = a.P ? a.P + (long long)v * fa.ps + c : nullptr;
    for (int h = 0; h < k; ++h) {
      cur.prefetch_next(h, k);
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (Pv && active) p = ld4s(Pv + h * fa.ph);
      const float* Xh = Xc + (long long)h * xh;
      const float* Th = (h == 0 ? T0 : Tk) + c;
      const unsigned Tsh = TSMEM ? sh_addr(Th) : 0u;
      const float* dinv_h = a.dinv ? a.dinv + h : nullptr;
      float4 z = has_t ? gather_segment<G, true, TSMEM>(cur, gm, active, Xh, xs, Th, Tsh, d, dinv_h, Kp)
                       : gather_segment<G, false, TSMEM>(cur, gm, active, Xh, xs, Th, Tsh, d, dinv_h, Kp);
      if (active) {
        const float s = fast_row_scale(a, v, h);
        z.x = act_fwd<ACT>(z.x * s) + p.x;
        z.y = act_fwd<ACT>(z.y * s) + p.y;
        z.z = act_fwd<ACT>(z.z * s) + p.z;
        z.w = act_fwd<ACT>(z.w * s) + p.w;
        if (a.eps) {
          const float4 x = ld4(Xh + (long long)v * xs);
          z.x = fmaf(self_c, x.x, z.x); z.y = fmaf(self_c, x.y, z.y);
          z.z = fmaf(self_c, x.z, z.z); z.w = fmaf(self_c, x.w, z.w);
        }
        if (FUSE) {
          const float4 th = lds4_sh(theta_sh + (h * d + c) * 4);
          o.x = fmaf(th.x, z.x, o.x); o.y = fmaf(th.y, z.y, o.y);
          o.z = fmaf(th.z, z.z, o.z); o.w = fmaf(th.w, z.w, o.w);
        } else {
          st4s(out + ((long long)v * k + h) * d + c, z);
        }
      }
      cur.advance();
    }
    if (FUSE && active) st4s(out + (long long)v * d + c, o);
  }
}

// ------------------------------------------------------------------------------------------------------------
// B1: per destination row -- recompute, Gs, dP, dtheta / deps partials
// ------------------------------------------------------------------------------------------------------------
template <int G, int ACT, bool FUSE, bool TSMEM>
__global__ void __launch_bounds__(256, 3)
agg_bwd_dst_fast_kernel(const FastArgs fa, const float* __restrict__ dOut, float* __restrict__ Gs,
                        float* __restrict__ dP, float* __restrict__ dtheta_part, float* __restrict__ deps_part,
                        int sm_tab_floats) {
  extern __shared__ __align__(16) float sm[];
  const kp_agg_desc& a = fa.d;
  const float *T0, *Tk, *theta;
  stage_tables<TSMEM, FUSE>(fa, sm, T0, Tk, theta);
  const int d = a.d, k = a.k, Kp = a.Kplan, xs = fa.xs, xh = fa.xh;
  const int lane = threadIdx.x & (G - 1);
  const int c = lane * 4;
  const bool active = c < d;
  const unsigned gm = group_mask<G>();
  constexpr int gpb = 256 / G;
  constexpr int dpad = 4 * G;
  const int gib = threadIdx.x / G;
  const float self_c = a.eps ? 1.f + __ldg(a.eps) : 0.f;
  const bool has_t = a.T0 != nullptr;
  const bool recompute = (ACT != KP_ACT_NONE) || (FUSE && dtheta_part != nullptr);
  // per-group private dtheta accumulators [k][dpad] behind the staged tables; each lane owns its columns
  float* th_all = sm + sm_tab_floats;
  float* th_acc = th_all + (size_t)gib * k * dpad + c;
  if (dtheta_part) {
    for (int i = threadIdx.x; i < gpb * k * dpad; i += blockDim.x) th_all[i] = 0.f;
    __syncthreads();
  }
  float eps_acc = 0.f;
  for (int v = blockIdx.x * gpb + gib; v < a.N; v += gridDim.x * gpb) {
    HopCursor<G> cur;
    if (recompute) cur.begin(a.rowptr + (long long)v * Kp, a.col, has_t ? a.attr16 : nullptr, lane);
    float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
    if (FUSE && active) go = ld4s(dOut + (long long)v * d + c);
    const float* Xc = a.X + c;
    const float* Pv = a.P ? a.P + (long long)v * fa.ps + c : nullptr;
    for (int h = 0; h < k; ++h) {
      if (recompute) cur.prefetch_next(h, k);
      const long long row = ((long long)v * k + h) * d + c;
      float4 dy = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active) {
        if (FUSE) {
          const float4 th = lds4_sh(sh_addr(theta) + (h * d + c) * 4);
          dy = make_float4(th.x * go.x, th.y * go.y, th.z * go.z, th.w * go.w);
        } else {
          dy = ld4s(dOut + row);
        }
        if (dP) st4s(dP + row, dy);
      }
      const float* Xh = Xc + (long long)h * xh;
      float4 g = dy;
      const float s = fast_row_scale(a, v, h);
      if (recompute) {
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool need_z = FUSE && dtheta_part != nullptr;
        if (need_z && Pv && active) p = ld4s(Pv + h * fa.ph);
        const float* Th = (h == 0 ? T0 : Tk) + c;
        const unsigned Tsh = TSMEM ? sh_addr(Th) : 0u;
        const float* dinv_h = a.dinv ? a.dinv + h : nullptr;
        float4 pre = has_t ? gather_segment<G, true, TSMEM>(cur, gm, active, Xh, xs, Th, Tsh, d, dinv_h, Kp)
                           : gather_segment<G, false, TSMEM>(cur, gm, active, Xh, xs, Th, Tsh, d, dinv_h, Kp);
        if (active) {
          pre.x *= s; pre.y *= s; pre.z *= s; pre.w *= s;
          if (need_z) {
            float4 z = make_float4(act_fwd<ACT>(pre.x) + p.x, act_fwd<ACT>(pre.y) + p.y,
                                   act_fwd<ACT>(pre.z) + p.z, act_fwd<ACT>(pre.w) + p.w);
            if (a.eps) {
              const float4 x = ld4(Xh + (long long)v * xs);
              z.x = fmaf(self_c, x.x, z.x); z.y = fmaf(self_c, x.y, z.y);
              z.z = fmaf(self_c, x.z, z.z); z.w = fmaf(self_c, x.w, z.w);
            }
            float4 t = lds4(th_acc + h * dpad);
            t.x = fmaf(go.x, z.x, t.x); t.y = fmaf(go.y, z.y, t.y);
            t.z = fmaf(go.z, z.z, t.z); t.w = fmaf(go.w, z.w, t.w);
            st4(th_acc + h * dpad, t);
          }
          g.x *= act_bwd<ACT>(pre.x); g.y *= act_bwd<ACT>(pre.y);
          g.z *= act_bwd<ACT>(pre.z); g.w *= act_bwd<ACT>(pre.w);
        }
        cur.advance();
      }
      if (active) {
        if (deps_part) {
          const float4 x = ld4(Xh + (long long)v * xs);
          eps_acc += dy.x * x.x + dy.y * x.y + dy.z * x.z + dy.w * x.w;
        }
        if (Gs) {
          g.x *= s; g.y *= s; g.z *= s; g.w *= s;
          st4(Gs + row, g);
        }
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
  if (deps_part) {
    __shared__ float red[256];
    __syncthreads();
    red[threadIdx.x] = eps_acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) deps_part[blockIdx.x] = red[0];
  }
}

// ------------------------------------------------------------------------------------------------------------
// B2: per source row, gather Gs through the transposed CSR
// ------------------------------------------------------------------------------------------------------------
template <int G, bool FUSE>
__global__ void __launch_bounds__(256, 4)
agg_bwd_src_fast_kernel(const FastArgs fa, const float* __restrict__ Gs, const float* __restrict__ dOut,
                        float* __restrict__ dX) {
  const kp_agg_desc& a = fa.d;
  const int d = a.d, k = a.k, Kp = a.Kplan;
  const int lane = threadIdx.x & (G - 1);
  const int c = lane * 4;
  const bool active = c < d;
  const unsigned gm = group_mask<G>();
  constexpr int gpb = 256 / G;
  const int gib = threadIdx.x / G;
  const float self_c = a.eps ? 1.f + __ldg(a.eps) : 0.f;
  const int gs = k * d;
  for (int u = blockIdx.x * gpb + gib; u < a.N; u += gridDim.x * gpb) {
    HopCursor<G> cur;
    cur.begin(a.rowptrT + (long long)u * Kp, a.colT, nullptr, lane);
    float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
    if (FUSE && a.eps && active) go = ld4(dOut + (long long)u * d + c);
    for (int h = 0; h < k; ++h) {
      cur.prefetch_next(h, k);
      const float* Gh = Gs + h * d + c;
      float4 acc = gather_segment<G, false, false>(cur, gm, active, Gh, gs, nullptr, 0u, d, nullptr, Kp);
      if (active) {
        if (a.dinv) {
          const float w = __ldg(a.dinv + (long long)u * Kp + h);
          acc.x *= w; acc.y *= w; acc.z *= w; acc.w *= w;
        }
        const long long row = ((long long)u * k + h) * d + c;
        if (a.eps) {
          float4 dy;
          if (FUSE) {
            const float4 th = ld4(a.theta + h * d + c);
            dy = make_float4(th.x * go.x, th.y * go.y, th.z * go.z, th.w * go.w);
          } else {
            dy = ld4(dOut + row);
          }
          acc.x = fmaf(self_c, dy.x, acc.x); acc.y = fmaf(self_c, dy.y, acc.y);
          acc.z = fmaf(self_c, dy.z, acc.z); acc.w = fmaf(self_c, dy.w, acc.w);
        }
        st4s(dX + row, acc);
      }
      cur.advance();
    }
  }
}

}  // namespace kp
