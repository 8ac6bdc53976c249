// agg.cu -- per-hop masked K-hop aggregation with fused epilogue, forward and backward (sm_100a).
//
// Reference semantics (see include/kpgnn.h for the formula): the message/aggregate/update triple of
// layers/KPGIN.py:100-121, KPGINplus.py:74-88, KPGCN.py:107-126, KPGraphSAGE.py:86-106, gine.py:52-59 and
// GeometricCombine (combine.py:43-58).  The reference materialises [E_K,k,d] message tensors; here each
// (dst,hop) row of the plan is a segment of a CSR and a GROUP of G lanes (G*VEC >= d, G<=32) owns one
// destination node: it walks the node's hop segments, gathers source rows with VEC-wide loads (float4 when
// aligned), applies activation / peripheral / self terms in registers and, when `fuse`, reduces over hops with
// theta so the [N,k,d] intermediate never reaches HBM.
//
// Backward is three deterministic passes (no float atomics):
//   B1 (dst rows)  recompute the pre-activation, emit Gs = dZ * act' * scale  [N,k,d], dP, dtheta/deps partials
//   B2 (src rows)  dX = wsrc * sum over the transposed CSR of Gs  (+ self term)
//   B3 (rows)      dT0/dTk by owner-computes partial tables in shared memory, then a fixed-order reduction
// HBM-bound gather work: no tensor cores on purpose.
#include <stdlib.h>

#include "agg_common.cuh"
#include "agg_fast_host.h"

namespace kp {
// count-matrix table-gradient pass (agg_b3_count.cu)
bool b3_count_ok(const kp_agg_desc& a, int G);
size_t b3_count_part_floats(const kp_agg_desc& a);
int b3_count(const kp_agg_desc& a, int G, const float* Gs, float* part, float* dT0, float* dTk, cudaStream_t st);

struct AggArgs {
  kp_agg_desc d;
  int G;        // lanes per node group (power of two <= 32)
  int gshift;   // log2(G)
};

// acc = sum_j wsrc * (X[col_j,h,c..] + T_h[attr_j,c..]) over row (v,h) of the plan, 4 gathers in flight
template <int VEC>
__device__ __forceinline__ Vf<VEC> gather_row(const kp_agg_desc& a, int v, int h, int c) {
  Vf<VEC> acc;
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc.v[i] = 0.f;
  const long long r = (long long)v * a.Kplan + h;
  const int b = __ldg(a.rowptr + r), e = __ldg(a.rowptr + r + 1);
  const float* __restrict__ T = a.T0 ? (h == 0 ? a.T0 : a.Tk) : nullptr;
  const float* __restrict__ Xh = a.X + (long long)h * a.x_hop_stride + c;
  for (int j = b; j < e; j += 4) {
    int u[4];
    int at[4];
    Vf<VEC> xv[4];
    const int n = min(4, e - j);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q < n) {
        u[q] = __ldg(a.col + j + q);
        at[q] = T ? (int)__ldg(a.attr16 + j + q) : 0;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < n) xv[q] = vload<VEC>(Xh + (long long)u[q] * a.x_node_stride);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q < n) {
        float w = a.dinv ? __ldg(a.dinv + (long long)u[q] * a.Kplan + h) : 1.f;
        if (T) {
          Vf<VEC> tv = vload<VEC>(T + (long long)at[q] * a.d + c);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc.v[i] += w * (xv[q].v[i] + tv.v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc.v[i] += w * xv[q].v[i];
        }
      }
    }
  }
  return acc;
}

__device__ __forceinline__ float row_scale(const kp_agg_desc& a, int v, int h) {
  float s = 1.f;
  if (a.dinv) s *= __ldg(a.dinv + (long long)v * a.Kplan + h);
  if (a.indeg) s *= 1.f / (float)max(__ldg(a.indeg + v), 1);
  return s;
}

template <int VEC, int ACT, bool FUSE>
__global__ void __launch_bounds__(256) agg_fwd_kernel(const AggArgs args, float* __restrict__ out) {
  const kp_agg_desc& a = args.d;
  const int G = args.G;
  const int lane = threadIdx.x & (G - 1);
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> args.gshift;
  const long long ngroups = ((long long)gridDim.x * blockDim.x) >> args.gshift;
  const float self_c = a.eps ? 1.f + __ldg(a.eps) : 0.f;
  for (long long v = group; v < a.N; v += ngroups) {
    for (int c = lane * VEC; c < a.d; c += G * VEC) {
      Vf<VEC> o;
#pragma unroll
      for (int i = 0; i < VEC; ++i) o.v[i] = 0.f;
      for (int h = 0; h < a.k; ++h) {
        Vf<VEC> z = gather_row<VEC>(a, (int)v, h, c);
        const float s = row_scale(a, (int)v, h);
#pragma unroll
        for (int i = 0; i < VEC; ++i) z.v[i] = act_fwd<ACT>(z.v[i] * s);
        if (a.P) {
          Vf<VEC> p = vload_stream<VEC>(a.P + v * a.p_node_stride + (long long)h * a.p_hop_stride + c);
#pragma unroll
          for (int i = 0; i < VEC; ++i) z.v[i] += p.v[i];
        }
        if (a.eps) {
          Vf<VEC> x = vload<VEC>(a.X + v * a.x_node_stride + (long long)h * a.x_hop_stride + c);
#pragma unroll
          for (int i = 0; i < VEC; ++i) z.v[i] += self_c * x.v[i];
        }
        if (FUSE) {
          Vf<VEC> th = vload<VEC>(a.theta + (long long)h * a.d + c);
#pragma unroll
          for (int i = 0; i < VEC; ++i) o.v[i] += th.v[i] * z.v[i];
        } else {
          vstore<VEC>(out + (v * a.k + h) * a.d + c, z);
        }
      }
      if (FUSE) vstore<VEC>(out + v * a.d + c, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// B1: per destination row.  RECOMPUTE=false is the linear, unfused case where only dP/deps are wanted.
// ---------------------------------------------------------------------------------------------------------
template <int VEC, int ACT, bool FUSE>
__global__ void __launch_bounds__(256)
agg_bwd_dst_kernel(const AggArgs args, const float* __restrict__ dOut, float* __restrict__ Gs,
                   float* __restrict__ dP, float* __restrict__ dtheta_part, float* __restrict__ deps_part) {
  extern __shared__ float smem[];
  const kp_agg_desc& a = args.d;
  const int G = args.G;
  const int lane = threadIdx.x & (G - 1);
  const int gib = threadIdx.x >> args.gshift;                 // group index inside the block
  const int gpb = blockDim.x >> args.gshift;                  // groups per block
  const long long group = (long long)blockIdx.x * gpb + gib;
  const long long ngroups = (long long)gridDim.x * gpb;
  const float self_c = a.eps ? 1.f + __ldg(a.eps) : 0.f;
  const int dpad = ((a.d + G * VEC - 1) / (G * VEC)) * (G * VEC);
  // per-group private dtheta accumulators [k][dpad]; every lane owns its columns -> no races
  float* th_acc = smem + (size_t)gib * a.k * dpad;
  if (dtheta_part) {
    for (int i = threadIdx.x; i < gpb * a.k * dpad; i += blockDim.x) smem[i] = 0.f;
    __syncthreads();
  }
  double eps_acc = 0.0;   // scalar reduction over N*k*d products: double keeps it within the 1e-5 parity bar
  for (long long v = group; v < a.N; v += ngroups) {
    for (int c = lane * VEC; c < a.d; c += G * VEC) {
      Vf<VEC> go;
      if (FUSE) go = vload_stream<VEC>(dOut + v * a.d + c);
      for (int h = 0; h < a.k; ++h) {
        Vf<VEC> dy;
        if (FUSE) {
          Vf<VEC> th = vload<VEC>(a.theta + (long long)h * a.d + c);
#pragma unroll
          for (int i = 0; i < VEC; ++i) dy.v[i] = th.v[i] * go.v[i];
        } else {
          dy = vload_stream<VEC>(dOut + (v * a.k + h) * a.d + c);
        }
        if (dP) vstore<VEC>(dP + (v * a.k + h) * a.d + c, dy);
        Vf<VEC> g = dy;
        const float s = row_scale(a, (int)v, h);
        if (ACT != KP_ACT_NONE || (FUSE && dtheta_part)) {
          Vf<VEC> pre = gather_row<VEC>(a, (int)v, h, c);
#pragma unroll
          for (int i = 0; i < VEC; ++i) pre.v[i] *= s;
          if (FUSE && dtheta_part) {
            Vf<VEC> z;
#pragma unroll
            for (int i = 0; i < VEC; ++i) z.v[i] = act_fwd<ACT>(pre.v[i]);
            if (a.P) {
              Vf<VEC> p = vload_stream<VEC>(a.P + v * a.p_node_stride + (long long)h * a.p_hop_stride + c);
#pragma unroll
              for (int i = 0; i < VEC; ++i) z.v[i] += p.v[i];
            }
            if (a.eps) {
              Vf<VEC> x = vload<VEC>(a.X + v * a.x_node_stride + (long long)h * a.x_hop_stride + c);
#pragma unroll
              for (int i = 0; i < VEC; ++i) z.v[i] += self_c * x.v[i];
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) th_acc[h * dpad + c + i] += go.v[i] * z.v[i];
          }
#pragma unroll
          for (int i = 0; i < VEC; ++i) g.v[i] *= act_bwd<ACT>(pre.v[i]);
        }
        if (deps_part) {
          Vf<VEC> x = vload<VEC>(a.X + v * a.x_node_stride + (long long)h * a.x_hop_stride + c);
#pragma unroll
          for (int i = 0; i < VEC; ++i) eps_acc += (double)(dy.v[i] * x.v[i]);
        }
        if (Gs) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) g.v[i] *= s;
          vstore<VEC>(Gs + (v * a.k + h) * a.d + c, g);
        }
      }
    }
  }
  if (dtheta_part) {
    __syncthreads();
    // fixed-order reduction over the block's groups
    for (int i = threadIdx.x; i < a.k * a.d; i += blockDim.x) {
      int h = i / a.d, c = i - h * a.d;
      float s = 0.f;
      for (int g = 0; g < gpb; ++g) s += smem[((size_t)g * a.k + h) * dpad + c];
      dtheta_part[(size_t)blockIdx.x * a.k * a.d + i] = s;
    }
  }
  if (deps_part) {
    __shared__ double red[256];
    __syncthreads();
    red[threadIdx.x] = eps_acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o && threadIdx.x + o < blockDim.x) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) deps_part[blockIdx.x] = (float)red[0];
  }
}

// ---------------------------------------------------------------------------------------------------------
// B2: per source row, gather Gs through the transposed CSR.  Gs is [N,k,d] contiguous (or dOut itself).
// ---------------------------------------------------------------------------------------------------------
template <int VEC, bool FUSE>
__global__ void __launch_bounds__(256)
agg_bwd_src_kernel(const AggArgs args, const float* __restrict__ Gs, const float* __restrict__ dOut,
                   float* __restrict__ dX) {
  const kp_agg_desc& a = args.d;
  const int G = args.G;
  const int lane = threadIdx.x & (G - 1);
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> args.gshift;
  const long long ngroups = ((long long)gridDim.x * blockDim.x) >> args.gshift;
  const float self_c = a.eps ? 1.f + __ldg(a.eps) : 0.f;
  for (long long u = group; u < a.N; u += ngroups) {
    for (int c = lane * VEC; c < a.d; c += G * VEC) {
      for (int h = 0; h < a.k; ++h) {
        Vf<VEC> acc;
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc.v[i] = 0.f;
        const long long r = u * a.Kplan + h;
        const int b = __ldg(a.rowptrT + r), e = __ldg(a.rowptrT + r + 1);
        for (int j = b; j < e; j += 4) {
          int vv[4];
          Vf<VEC> gv[4];
          const int n = min(4, e - j);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < n) vv[q] = __ldg(a.colT + j + q);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < n) gv[q] = vload<VEC>(Gs + ((long long)vv[q] * a.k + h) * a.d + c);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < n) {
#pragma unroll
              for (int i = 0; i < VEC; ++i) acc.v[i] += gv[q].v[i];
            }
        }
        if (a.dinv) {
          const float w = __ldg(a.dinv + r);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc.v[i] *= w;
        }
        if (a.eps) {
          Vf<VEC> dy;
          if (FUSE) {
            Vf<VEC> go = vload<VEC>(dOut + u * a.d + c);
            Vf<VEC> th = vload<VEC>(a.theta + (long long)h * a.d + c);
#pragma unroll
            for (int i = 0; i < VEC; ++i) dy.v[i] = go.v[i] * th.v[i];
          } else {
            dy = vload<VEC>(dOut + (u * a.k + h) * a.d + c);
          }
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc.v[i] += self_c * dy.v[i];
        }
        vstore<VEC>(dX + (u * a.k + h) * a.d + c, acc);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// B3: embedding-table gradients.  Thread (rl, t) owns VEC columns of a row-lane-private table copy in shared
// memory, walks its rows in a fixed order, so the float sums are reproducible without atomics.
// ---------------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256)
agg_bwd_table_kernel(const kp_agg_desc a, const float* __restrict__ Gs, int cw, int rl_count, int rows_per_block,
                     float* __restrict__ part) {
  extern __shared__ float smem[];
  const int trows = a.rows0 + a.rowsk;
  const size_t tsz = (size_t)trows * a.d;
  for (size_t i = threadIdx.x; i < tsz * rl_count; i += blockDim.x) smem[i] = 0.f;
  __syncthreads();
  const int rl = threadIdx.x / cw, t = threadIdx.x - rl * cw;
  const int c = t * VEC;
  const long long R = (long long)a.N * a.k;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(R, r0 + rows_per_block);
  if (rl < rl_count && c < a.d) {
    float* tab = smem + tsz * rl;
    for (long long row = r0 + rl; row < r1; row += rl_count) {
      const int v = (int)(row / a.k), h = (int)(row - (long long)v * a.k);
      const long long pr = (long long)v * a.Kplan + h;
      const int b = __ldg(a.rowptr + pr), e = __ldg(a.rowptr + pr + 1);
      if (b == e) continue;
      const Vf<VEC> g = vload_stream<VEC>(Gs + row * a.d + c);
      const int base = (h == 0) ? 0 : a.rows0;
      for (int j = b; j < e; ++j) {
        const int at = (int)__ldg(a.attr16 + j);
        const float w = a.dinv ? __ldg(a.dinv + (long long)__ldg(a.col + j) * a.Kplan + h) : 1.f;
        float* dst = tab + (size_t)(base + at) * a.d + c;
#pragma unroll
        for (int i = 0; i < VEC; ++i) dst[i] += w * g.v[i];
      }
    }
  }
  __syncthreads();
  for (size_t i = threadIdx.x; i < tsz; i += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < rl_count; ++q) s += smem[tsz * q + i];
    part[(size_t)blockIdx.x * tsz + i] = s;
  }
}

// Fast variant (float4 rows, d <= 128): a group of G lanes owns a private sub-table in shared memory and a
// contiguous range of (node,hop) rows.  Only ~8 warps fit next to the sub-tables, so nothing hides latency but
// the warp itself: the loads of batch i+1 (row pointers + Gs rows, RB rows) are issued before batch i is folded
// into the table (two register sets), and the fold is instruction-lean -- the first version spent 64 warp
// instructions per entry on 64-bit index arithmetic (profiles/r1q_b3.txt): (node,hop) advance incrementally,
// table addresses are 32-bit shared addresses, the read-modify-write is LDS.128 / 2 FADD2 / STS.128.
// Row order inside a group and group order in the final reduction are fixed -> bit-reproducible.
template <int G, bool NORM>
__global__ void __launch_bounds__(256)
agg_bwd_table_fast_kernel(const kp_agg_desc a, const float* __restrict__ Gs, int ngroups_cta, int rows_per_group,
                          float* __restrict__ part) {
  extern __shared__ __align__(16) float smem[];
  constexpr int RB = 8;
  const int trows = a.rows0 + a.rowsk;
  const int tsz = trows * a.d;
  for (int i = threadIdx.x * 4; i < tsz * ngroups_cta; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(smem + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int lane = threadIdx.x & (G - 1);
  const int gib = threadIdx.x / G;
  const int c = min(lane * 4, a.d - 4);
  const bool active = lane * 4 < a.d;
  const long long R = (long long)a.N * a.k;
  const long long gid = (long long)blockIdx.x * ngroups_cta + gib;
  const long long r0 = gid * rows_per_group;
  const long long r1 = min(R, r0 + rows_per_group);
  const unsigned d4 = (unsigned)a.d * 4u;
  const unsigned tab0 = (unsigned)__cvta_generic_to_shared(smem + (size_t)tsz * gib + c);
  const unsigned tabk = tab0 + (unsigned)a.rows0 * d4;
  if (gib < ngroups_cta && active && r0 < r1) {
    struct Batch {
      int b[RB], e[RB];
      unsigned tb[RB];           // table base of the row's hop (0 = no row)
      int hh[RB];
      float4 g[RB];
    };
    // running (node, hop) of the next row to load
    int lv = (int)(r0 / a.k), lh = (int)(r0 - (long long)lv * a.k);
    const int* rp = a.rowptr + (size_t)lv * a.Kplan + lh;
    const float* gp = Gs + (size_t)r0 * a.d + c;
    long long lrow = r0;
    auto load = [&](Batch& t) {
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        t.b[u] = t.e[u] = 0;
        t.tb[u] = tab0;
        t.hh[u] = 0;
        t.g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lrow < r1) {
          t.b[u] = __ldg(rp);
          t.e[u] = __ldg(rp + 1);
          t.tb[u] = lh == 0 ? tab0 : tabk;
          t.hh[u] = lh;
          t.g[u] = __ldcs(reinterpret_cast<const float4*>(gp));
          ++lrow;
          gp += a.d;
          ++rp;
          if (++lh == a.k) {
            lh = 0;
            rp += a.Kplan - a.k;
          }
        }
      }
    };
    auto fold = [&](const Batch& t) {
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const unsigned long long glo = (unsigned long long)__float_as_uint(t.g[u].x) |
                                       ((unsigned long long)__float_as_uint(t.g[u].y) << 32);
        const unsigned long long ghi = (unsigned long long)__float_as_uint(t.g[u].z) |
                                       ((unsigned long long)__float_as_uint(t.g[u].w) << 32);
        const uint16_t* ap = a.attr16 + t.b[u];
        const int n = t.e[u] - t.b[u];
        for (int j = 0; j < n; ++j) {
          const unsigned addr = t.tb[u] + (unsigned)__ldg(ap + j) * d4;
          unsigned long long xlo, xhi;
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(xlo), "=l"(xhi) : "r"(addr));
          if (NORM) {
            const float w = __ldg(a.dinv + (long long)__ldg(a.col + t.b[u] + j) * a.Kplan + t.hh[u]);
            unsigned long long ww;
            asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(xlo) : "l"(ww), "l"(glo));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(xhi) : "l"(ww), "l"(ghi));
          } else {
            asm("add.rn.f32x2 %0, %0, %1;" : "+l"(xlo) : "l"(glo));
            asm("add.rn.f32x2 %0, %0, %1;" : "+l"(xhi) : "l"(ghi));
          }
          asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(xlo), "l"(xhi) : "memory");
        }
      }
    };
    Batch A, B;
    load(A);
    for (long long rb = r0; rb < r1; rb += 2 * RB) {
      load(B);
      fold(A);
      load(A);
      fold(B);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < tsz; i += blockDim.x) {      // fixed-order reduction over the CTA's groups
    float s = 0.f;
    for (int q = 0; q < ngroups_cta; ++q) s += smem[(size_t)tsz * q + i];
    part[(size_t)blockIdx.x * tsz + i] = s;
  }
}

// fallback for tables that do not fit in shared memory: global float atomics (NOT bitwise reproducible)
__global__ void agg_bwd_table_atomic_kernel(const kp_agg_desc a, const float* __restrict__ Gs,
                                            float* __restrict__ dT0, float* __restrict__ dTk) {
  const long long R = (long long)a.N * a.k;
  const long long total = R * a.d;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / a.d;
    const int c = (int)(i - row * a.d);
    const int v = (int)(row / a.k), h = (int)(row - (long long)v * a.k);
    const long long pr = (long long)v * a.Kplan + h;
    const int b = __ldg(a.rowptr + pr), e = __ldg(a.rowptr + pr + 1);
    if (b == e) continue;
    const float g = Gs[i];
    float* T = (h == 0) ? dT0 : dTk;
    for (int j = b; j < e; ++j) {
      const float w = a.dinv ? __ldg(a.dinv + (long long)__ldg(a.col + j) * a.Kplan + h) : 1.f;
      atomicAdd(T + (size_t)__ldg(a.attr16 + j) * a.d + c, w * g);
    }
  }
}

// out[i] = sum_b part[b*n + i].  One warp per output element: lane l adds the partials b = l, l+32, ... in order,
// then a fixed shuffle tree combines the 32 lane sums -> the summation order is a function of (nblocks) only, so
// the result is bit-reproducible; optional second destination split at n0 (dT0 | dTk).
__global__ void reduce_partials_kernel(const float* __restrict__ part, int nblocks, int n, int n0,
                                       float* __restrict__ out0, float* __restrict__ out1) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  float s = 0.f;
  for (int b = lane; b < nblocks; b += 32) s += __ldcs(part + (size_t)b * n + i);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (i < n0) {
      if (out0) out0[i] = s;
    } else {
      if (out1) out1[i - n0] = s;
    }
  }
}

// dtheta[h,c] = sum_b part[b*n + h*d + c] (same order as reduce_partials_kernel) fused with GeometricCombine's
// backward (combine.py:51-58): one warp per channel c; dalphas[c] = a(1-a) * sum_h theta_h (dtheta_h - dot)
// ((1-a)^h - a h (1-a)^(h-1)),  dot = sum_h theta_h dtheta_h,  a = sigmoid(alphas[c]).  dtheta itself is optional.
__global__ void __launch_bounds__(256)
dtheta_geo_bwd_kernel(const float* __restrict__ part, int nblocks, int k, int d, const float* __restrict__ alphas,
                      const float* __restrict__ theta, float* __restrict__ dtheta, float* __restrict__ dalphas) {
  // one CTA per channel; warp w sums hop h = w, w+8, ... (all hops' loads in flight together: one L2 round trip)
  __shared__ float sd[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x;
  const int n = k * d;
  for (int h = w; h < k; h += 8) {
    float s = 0.f;
    for (int b = lane; b < nblocks; b += 32) s += __ldcs(part + (size_t)b * n + (size_t)h * d + c);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      sd[h] = s;
      if (dtheta) dtheta[(size_t)h * d + c] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float a = 1.f / (1.f + expf(-alphas[c]));
    float dot = 0.f;
    for (int h = 0; h < k; ++h) dot += theta[(size_t)h * d + c] * sd[h];
    float da = 0.f, pw = 1.f, pwm1 = 0.f;
    for (int h = 0; h < k; ++h) {
      const float dt = theta[(size_t)h * d + c] * (sd[h] - dot);
      da += dt * (pw - a * (float)h * pwm1);
      pwm1 = pw;
      pw *= (1.f - a);
    }
    dalphas[c] = da * a * (1.f - a);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side: configuration shared by forward, backward and the workspace-size query
// ---------------------------------------------------------------------------------------------------------
struct Config {
  int vec, G, gshift;
  int grid;         // forward / B2 grid (256 threads)
  int grid_b1;      // B1 grid (bounded so the dtheta partials stay small)
  size_t smem_b1;   // dynamic smem of B1 when dtheta is wanted
  bool need_gs;     // Gs workspace needed (otherwise Gs == dOut)
  int cw, rl, grid_b3, rows_per_block;
  size_t smem_b3;
  bool table_atomic;
  bool b3_fast;
  bool b3_count;    // count-matrix kernel with register accumulators (agg_b3_count.cu)
  size_t b3_part_floats;
  int b3_G, b3_groups, b3_threads, b3_rows_per_group;
  // fast float4 path (agg_fast.cuh)
  bool fast, fextra;
  int ftab, fG, fgrid, fgrid_b1, stage_floats;   // stage_floats = tables (+ theta) staged in smem, in floats
  size_t fsmem_fwd;
  // lean B1 (agg_lean.cuh): packed-math kernel for the layers without self term / norm / mean
  bool lean_b1;
  int lean_b1_threads;
  size_t lean_b1_smem;
  // the whole backward as one block-resident kernel (agg_block_bwd.cu)
  bool fb;
};

static int g_force_generic = 0;   // test hook (kp_agg_set_force_generic): exercise the generic kernels
int g_geom_max_ctas = 0, g_geom_lean_threads = 0;   // test hook (kp_agg_set_launch_geometry), see agg_fast_host.h

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
static bool aligned8(const void* p) { return ((uintptr_t)p & 7) == 0; }

static int make_config(const kp_agg_desc& a, Config* c) {
  KP_CHECK_ARG(a.N >= 0 && a.k >= 1 && a.k <= a.Kplan && a.d >= 1, "kp_agg: bad sizes N=%d k=%d Kplan=%d d=%d",
               a.N, a.k, a.Kplan, a.d);
  KP_CHECK_ARG(a.rowptr && a.col && a.X, "kp_agg: null plan or X");
  KP_CHECK_ARG(!a.T0 || (a.attr16 && a.rows0 > 0 && (a.k == 1 || (a.Tk && a.rowsk > 0))),
               "kp_agg: embedding tables incomplete");
  KP_CHECK_ARG(!a.fuse || a.theta, "kp_agg: fuse requires theta");
  KP_CHECK_ARG(a.act >= 0 && a.act <= 2, "kp_agg: bad activation %d", a.act);
  auto ok = [&](int v) {
    if (a.d % v) return false;
    if (a.x_node_stride % v || a.x_hop_stride % v) return false;
    if (a.P && (a.p_node_stride % v || a.p_hop_stride % v)) return false;
    const void* ptrs[] = {a.X, a.P, a.T0, a.Tk, a.theta};
    for (const void* p : ptrs)
      if (p && !(v == 4 ? aligned16(p) : aligned8(p))) return false;
    return true;
  };
  c->vec = ok(4) ? 4 : (ok(2) ? 2 : 1);
  int lanes = (a.d + c->vec - 1) / c->vec;
  int G = 1, sh = 0;
  while (G < lanes && G < 32) {
    G <<= 1;
    ++sh;
  }
  c->G = G;
  c->gshift = sh;
  const int gpb = 256 / G;
  long long want = ((long long)a.N + gpb - 1) / gpb;
  c->grid = geom_cap(want > kNumSMs * 16 ? kNumSMs * 16 : want);
  c->grid_b1 = geom_cap(want > kNumSMs * 2 ? kNumSMs * 2 : want);
  const int dpad = ((a.d + G * c->vec - 1) / (G * c->vec)) * (G * c->vec);
  c->smem_b1 = sizeof(float) * (size_t)gpb * a.k * dpad;
  c->need_gs = (a.act != KP_ACT_NONE) || a.fuse || a.dinv || a.indeg;
  // fast path eligibility: float4 everywhere, one column chunk, every element offset below 2^32
  const unsigned long long lim = 0xffffffffULL;
  const unsigned long long n1 = (unsigned long long)(a.N > 0 ? a.N : 1);
  bool fextra = false;
  c->fast = !g_force_generic && c->vec == 4 && a.d <= 128 &&
            fast_combo(a.act, a.fuse != 0, a.dinv || a.indeg || a.eps, &fextra) &&
            n1 * (unsigned long long)a.x_node_stride + (unsigned long long)a.k * a.x_hop_stride < lim &&
            (!a.P || n1 * (unsigned long long)a.p_node_stride + (unsigned long long)a.k * a.p_hop_stride < lim) &&
            n1 * (unsigned long long)a.k * a.d < lim;
  c->fextra = fextra;
  c->lean_b1 = false;
  if (c->fast) {
    int fl = a.d / 4, fG = 4;
    while (fG < fl) fG <<= 1;
    c->fG = fG;
    const int fgpb = 256 / fG;
    const long long fwant = ((long long)a.N + fgpb - 1) / fgpb;
    c->fgrid = geom_cap(fwant > kNumSMs * KP_FWD_MINB ? kNumSMs * KP_FWD_MINB : fwant);
    c->fgrid_b1 = geom_cap(fwant > kNumSMs * 3 ? kNumSMs * 3 : fwant);
    const long long tabf = a.T0 ? (long long)(a.rows0 + a.rowsk) * a.d : 0;
    c->ftab = !a.T0 ? TAB_NONE : (tabf * 4 <= 56 * 1024 ? TAB_SMEM : TAB_GLOBAL);
    c->stage_floats = (int)((c->ftab == TAB_SMEM ? tabf : 0) + (a.fuse ? a.k * a.d : 0));
    c->fsmem_fwd = sizeof(float) * (size_t)c->stage_floats;
    c->grid_b1 = c->fgrid_b1;
    c->smem_b1 = sizeof(float) * ((size_t)c->stage_floats + (size_t)fgpb * a.k * 4 * fG);
    c->lean_b1 = fast_lean_enabled() && (fG == 32 || fG == 16) && a.k + 1 <= fG && !a.dinv && !a.indeg && !a.eps &&
                 c->ftab != TAB_GLOBAL && (a.act != KP_ACT_NONE || a.fuse);
    if (c->lean_b1) {
      // largest CTA whose dtheta accumulators fit next to the tables; small batches keep 256 threads (more CTAs)
      static const int env_b1_threads_ = getenv("KP_LEAN_B1_THREADS") ? atoi(getenv("KP_LEAN_B1_THREADS")) : 0;
      const int env_b1_threads = g_geom_lean_threads ? g_geom_lean_threads : env_b1_threads_;
      int threads = 1024;
      const int balanced = env_b1_threads ? 0 : lean_balanced_threads(a.N, fG, 2);
      if (balanced) {      // one CTA per SM, the same number of groups each (see lean_balanced_threads)
        const int gpb = balanced / fG;
        const size_t sm = sizeof(float) * ((size_t)c->stage_floats + (size_t)gpb * (lean_group_scratch_bytes(fG) / 4) +
                                           (a.fuse ? (size_t)gpb * a.k * 4 * fG : 0));
        if (sm <= 200 * 1024) {
          c->lean_b1_threads = balanced;
          c->lean_b1_smem = sm;
          threads = 0;
        }
      }
      while (threads) {
        const int gpb = threads / fG;
        const size_t sm = sizeof(float) * ((size_t)c->stage_floats + (size_t)gpb * (lean_group_scratch_bytes(fG) / 4) +
                                           (a.fuse ? (size_t)gpb * a.k * 4 * fG : 0));
        const bool enough = (long long)a.N * 2 >= (long long)kNumSMs * gpb;   // half a wave of CTAs is enough (measured)
        if (env_b1_threads ? threads <= env_b1_threads && (sm <= 200 * 1024 || threads == 256)
                           : ((sm <= 200 * 1024 && enough) || threads == 256)) {
          c->lean_b1_threads = threads;
          c->lean_b1_smem = sm;
          break;
        }
        threads >>= 1;
      }
      if (c->lean_b1_smem > 200 * 1024) {
        c->lean_b1 = false;
      } else {
        const int gpb = c->lean_b1_threads / fG;
        const long long want = ((long long)a.N + gpb - 1) / gpb;
        const long long cap = (long long)kNumSMs * (1024 / c->lean_b1_threads);
        c->grid_b1 = geom_cap(want > cap ? cap : want);
      }
    }
  }
  c->fb = c->fast && fast_lean_enabled() && block_bwd_eligible(a, c->fG, c->ftab, true);
  if (c->fb && c->grid_b1 < block_bwd_grid()) c->grid_b1 = block_bwd_grid();    // sizes the dtheta partials
  // table pass
  const int trows = a.T0 ? a.rows0 + (a.k > 1 ? a.rowsk : 0) : 0;
  c->cw = lanes;
  c->table_atomic = false;
  c->rl = 0;
  c->grid_b3 = 0;
  c->rows_per_block = 0;
  c->smem_b3 = 0;
  c->b3_fast = false;
  c->b3_count = false;
  c->b3_part_floats = 0;
  // count-matrix kernel: needs the plan's largest-embedding-row statistics (desc.amax0 / amaxk) to be <= 31;
  // KP_B3_COUNT=0 keeps the sub-table kernel (A/B measurements)
  static const bool use_count = !(getenv("KP_B3_COUNT") && atoi(getenv("KP_B3_COUNT")) == 0);
  if (use_count && trows > 0 && c->fast && fast_lean_enabled() && b3_count_ok(a, c->fG)) {
    c->b3_count = true;
    c->grid_b3 = 1;
    c->b3_part_floats = b3_count_part_floats(a);
  } else if (trows > 0) {
    size_t tsz = sizeof(float) * (size_t)(a.rows0 + a.rowsk) * a.d;
    const long long R = (long long)a.N * a.k;
    if (c->fast && tsz <= 200 * 1024) {
      // fast table pass: groups of fG lanes, one private sub-table each
      int groups = 256 / c->fG;
      while (groups > 1 && tsz * groups > 200 * 1024) groups >>= 1;
      int threads = groups * c->fG;
      if (threads < 32) threads = 32;
      c->b3_fast = true;
      c->b3_G = c->fG;
      c->b3_groups = groups;
      c->b3_threads = threads;
      c->smem_b3 = tsz * groups;
      // enough rows per group to amortise zeroing/flushing the sub-table, at most one CTA per SM
      long long want_groups = (R + 31) / 32;
      long long blocks = (want_groups + groups - 1) / groups;
      if (blocks > kNumSMs) blocks = kNumSMs;
      if (blocks < 1) blocks = 1;
      c->grid_b3 = (int)blocks;
      const long long tg = blocks * groups;
      c->b3_rows_per_group = (int)((R + tg - 1) / tg);
    } else {
      int rl = 256 / lanes;
      if (rl > 8) rl = 8;
      while (rl > 1 && tsz * rl > 96 * 1024) rl >>= 1;
      if (tsz * rl > 200 * 1024 || lanes > 256) {
        c->table_atomic = true;
      } else {
        c->rl = rl;
        c->smem_b3 = tsz * rl;
        long long blocks = (R + 255) / 256;  // at least 256 rows per block keeps the partials small
        if (blocks > kNumSMs * 2) blocks = kNumSMs * 2;
        if (blocks < 1) blocks = 1;
        c->grid_b3 = (int)blocks;
        c->rows_per_block = (int)((R + blocks - 1) / blocks);
      }
    }
  }
  return 0;
}

struct WsLayout {
  size_t gs, dtheta, deps, table, total;
};

static WsLayout ws_layout(const kp_agg_desc& a, const Config& c) {
  WsLayout w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 256);
    return o;
  };
  w.gs = take(c.need_gs ? sizeof(float) * (size_t)a.N * a.k * a.d : 0);
  w.dtheta = take(sizeof(float) * (size_t)c.grid_b1 * a.k * a.d);
  w.deps = take(sizeof(float) * (size_t)c.grid_b1);
  size_t tab_bytes = c.b3_count ? sizeof(float) * c.b3_part_floats
                                : (c.grid_b3 ? sizeof(float) * (size_t)c.grid_b3 * (a.rows0 + a.rowsk) * a.d : 0);
  if (c.fb && a.T0) {
    const size_t fbb = sizeof(float) * (size_t)block_bwd_grid() * (a.rows0 + a.rowsk) * a.d;
    if (fbb > tab_bytes) tab_bytes = fbb;
  }
  w.table = take(tab_bytes);
  w.total = off;
  return w;
}

template <int VEC, int ACT, bool FUSE>
static int launch_fwd(const AggArgs& args, const Config& c, float* out, cudaStream_t st) {
  KP_LAUNCH((agg_fwd_kernel<VEC, ACT, FUSE>), c.grid, 256, 0, st, args, out);
  return 0;
}

template <int VEC, int ACT, bool FUSE>
static int launch_b1(const AggArgs& args, const Config& c, const float* dOut, float* Gs, float* dP,
                     float* dth, float* dep, cudaStream_t st) {
  size_t smem = dth ? c.smem_b1 : 0;
  if (smem > 32 * 1024) {
    KP_CUDA(cudaFuncSetAttribute(agg_bwd_dst_kernel<VEC, ACT, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  }
  KP_LAUNCH((agg_bwd_dst_kernel<VEC, ACT, FUSE>), c.grid_b1, 256, smem, st, args, dOut, Gs, dP, dth, dep);
  return 0;
}


// ---- fast-path arguments (kernels live in agg_fast_*.cu) ----
static FastArgs make_fast_args(const kp_agg_desc& a) {
  FastArgs fa;
  fa.d = a;
  fa.xs = (unsigned)a.x_node_stride;
  fa.xh = (unsigned)a.x_hop_stride;
  fa.ps = (unsigned)a.p_node_stride;
  fa.ph = (unsigned)a.p_hop_stride;
  return fa;
}

#define KP_DISPATCH_VAF(FN, vec, act, fuse, ...)                                          \
  do {                                                                                    \
    int _rc = 1;                                                                          \
    if (vec == 4) {                                                                       \
      KP_DISPATCH_AF(FN, 4, act, fuse, __VA_ARGS__);                                      \
    } else if (vec == 2) {                                                                \
      KP_DISPATCH_AF(FN, 2, act, fuse, __VA_ARGS__);                                      \
    } else {                                                                              \
      KP_DISPATCH_AF(FN, 1, act, fuse, __VA_ARGS__);                                      \
    }                                                                                     \
    if (_rc) return _rc;                                                                  \
  } while (0)
#define KP_DISPATCH_AF(FN, V, act, fuse, ...)                                             \
  do {                                                                                    \
    if (act == KP_ACT_GELU) {                                                             \
      _rc = fuse ? FN<V, KP_ACT_GELU, true>(__VA_ARGS__) : FN<V, KP_ACT_GELU, false>(__VA_ARGS__); \
    } else if (act == KP_ACT_RELU) {                                                      \
      _rc = fuse ? FN<V, KP_ACT_RELU, true>(__VA_ARGS__) : FN<V, KP_ACT_RELU, false>(__VA_ARGS__); \
    } else {                                                                              \
      _rc = fuse ? FN<V, KP_ACT_NONE, true>(__VA_ARGS__) : FN<V, KP_ACT_NONE, false>(__VA_ARGS__); \
    }                                                                                     \
  } while (0)

}  // namespace kp

extern "C" {

int kp_agg_forward(const kp_agg_desc* desc, float* out, void* stream) {
  KP_CHECK_ARG(desc && out, "kp_agg_forward: null argument");
  kp::Config c;
  if (kp::make_config(*desc, &c)) return 1;
  if (desc->N == 0) return 0;
  KP_CHECK_ARG(c.vec == 1 || (c.vec == 4 ? kp::aligned16(out) : kp::aligned8(out)),
               "kp_agg_forward: output not aligned for %d-wide stores", c.vec);
  kp::AggArgs args{*desc, c.G, c.gshift};
  cudaStream_t st = (cudaStream_t)stream;
  if (c.fast && kp::tile_eligible(*desc, c.ftab))
    return kp::tile_fwd(kp::make_fast_args(*desc), c.fG, desc->act, c.ftab, c.fextra, out, st);
  if (c.fast)
    return kp::fast_fwd(kp::make_fast_args(*desc), c.fG, desc->act, desc->fuse != 0, c.ftab, c.fextra, c.fgrid,
                        c.fsmem_fwd, out, st);
  KP_DISPATCH_VAF(kp::launch_fwd, c.vec, desc->act, desc->fuse, args, c, out, st);
  return 0;
}

int kp_agg_set_force_generic(int flag) {
  // bit 0: generic kernels instead of the float4 fast path; bit 1: fast path without the cp.async ring;
  // bit 2: fast path without the lean (packed-math, L2-prefetch) kernels
  kp::g_force_generic = (flag & 1) ? 1 : 0;
  kp::fast_fwd_set_ring((flag & 2) ? 0 : 1);
  kp::fast_fwd_set_lean((flag & 4) ? 0 : 1);
  // TMA-staged forward kernel (agg_tma.cuh): off by default (measured slightly slower than the lean kernel);
  // bit 4: use it for every eligible call, however small (tests); bit 5: use it for large batches only
  kp::fast_fwd_set_tma((flag & 16) ? 2 : ((flag & 32) ? 1 : 0));
  kp::tile_set_mode((flag & 64) ? 0 : 1);
  kp::block_bwd_set_mode((flag & 64) ? 0 : 1);
  return 0;
}

int kp_agg_set_launch_geometry(int max_ctas, int lean_threads) {
  if (max_ctas < 0 || (lean_threads != 0 && (lean_threads < 256 || lean_threads > 1024 || lean_threads % 32))) {
    kp::set_error("kp_agg_set_launch_geometry: max_ctas >= 0, lean_threads 0 or a multiple of 32 in [256,1024]");
    return 1;
  }
  kp::g_geom_max_ctas = max_ctas;
  kp::g_geom_lean_threads = lean_threads;
  return 0;
}

int kp_agg_backward_workspace_bytes(const kp_agg_desc* desc, size_t* bytes) {
  KP_CHECK_ARG(desc && bytes, "kp_agg_backward_workspace_bytes: null argument");
  kp::Config c;
  if (kp::make_config(*desc, &c)) return 1;
  *bytes = kp::ws_layout(*desc, c).total;
  return 0;
}

static bool chunkable(const kp_agg_desc& a, const kp::Config& c) {
  // the kernels whose per-node streams are indexed relative to the descriptor's pointers and whose gathers use the
  // batch-wide ids: lean B1, lean gather B2 (no extras), count-matrix or sub-table B3
  return c.fast && c.lean_b1 && c.need_gs && !c.fextra && !c.fb && !c.table_atomic && !a.block_ptr && !a.eps &&
         kp::fast_lean_enabled() && (c.fG == 32 || c.fG == 16) && a.k + 1 <= c.fG;
}

int kp_agg_backward_chunkable(const kp_agg_desc* desc, int32_t* ok) {
  KP_CHECK_ARG(desc && ok, "kp_agg_backward_chunkable: null argument");
  kp::Config c;
  if (kp::make_config(*desc, &c)) return 1;
  *ok = chunkable(*desc, c) ? 1 : 0;
  return 0;
}

int kp_agg_backward(const kp_agg_desc* desc, const float* dOut, float* dX, float* dP, float* dT0, float* dTk,
                    float* dtheta, float* deps, void* workspace, size_t workspace_bytes, void* stream) {
  KP_CHECK_ARG(desc && dOut, "kp_agg_backward: null argument");
  const kp_agg_desc& a = *desc;
  kp::Config c;
  if (kp::make_config(a, &c)) return 1;
  KP_CHECK_ARG(a.node_base == 0 || (a.node_base > 0 && chunkable(a, c)),
               "kp_agg_backward: node_base is only honoured where kp_agg_backward_chunkable() says so");
  KP_CHECK_ARG(a.rowptrT && a.colT, "kp_agg_backward: plan has no transposed CSR");
  KP_CHECK_ARG(!dtheta || a.fuse, "kp_agg_backward: dtheta requires fuse");
  const bool geo = a.fuse && a.geo_alphas && a.geo_dalphas;
  KP_CHECK_ARG(!geo || a.k <= 32, "kp_agg_backward: fused GeometricCombine backward needs k <= 32");
  KP_CHECK_ARG(!deps || a.eps, "kp_agg_backward: deps requires eps");
  KP_CHECK_ARG((!dT0 && !dTk) || a.T0, "kp_agg_backward: table gradients requested without tables");
  kp::WsLayout w = kp::ws_layout(a, c);
  KP_CHECK_ARG(workspace_bytes >= w.total && (workspace || w.total == 0), "kp_agg_backward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tn = (size_t)(a.rows0 + a.rowsk) * a.d;
  if (dX && (a.dx_node_stride || a.dx_hop_stride || a.dx_accumulate)) {
    // strided / accumulated dX: lean gather kernel only; refuse before anything is launched
    const bool lean = c.fast && !c.fextra && kp::fast_lean_enabled() && (c.fG == 32 || c.fG == 16) && a.k + 1 <= c.fG;
    if (!lean || a.dx_node_stride % 4 || a.dx_hop_stride % 4 || a.dx_node_stride < 0 || a.dx_hop_stride < 0 ||
        (unsigned long long)a.N * (unsigned long long)(a.dx_node_stride ? a.dx_node_stride : a.k * a.d) >= 0xffffffffull) {
      kp::set_error("kp_agg_backward: strided / accumulated dX is only available on the lean gather kernel");
      return 3;
    }
  }
  if (a.N == 0) {
    if (dT0) KP_CUDA(cudaMemsetAsync(dT0, 0, sizeof(float) * (size_t)a.rows0 * a.d, st));
    if (dTk) KP_CUDA(cudaMemsetAsync(dTk, 0, sizeof(float) * (size_t)a.rowsk * a.d, st));
    if (dtheta) KP_CUDA(cudaMemsetAsync(dtheta, 0, sizeof(float) * (size_t)a.k * a.d, st));
    if (deps) KP_CUDA(cudaMemsetAsync(deps, 0, sizeof(float), st));
    return 0;
  }
  if (c.fast)
    KP_CHECK_ARG(kp::aligned16(dOut) && kp::aligned16(dX) && kp::aligned16(dP) && kp::aligned16(workspace),
                 "kp_agg_backward: dOut/dX/dP/workspace must be 16-byte aligned");
  char* ws = (char*)workspace;
  float* Gs = c.need_gs ? (float*)(ws + w.gs) : nullptr;
  float* dth_part = (dtheta || geo) ? (float*)(ws + w.dtheta) : nullptr;
  float* dep_part = deps ? (float*)(ws + w.deps) : nullptr;
  kp::AggArgs args{a, c.G, c.gshift};
  const bool want_table = (dT0 || dTk);
  if (c.fb && !deps && (!want_table || (dT0 && (dTk || a.k == 1)))) {
    // ONE block-resident kernel: recompute + dP + dtheta partials, dX through shared memory, table partials
    float* dPk = dP;
    if (!a.fuse && dP == dOut) dPk = nullptr;
    float* tab_part = want_table ? (float*)(ws + w.table) : nullptr;
    int fgrid = 0;
    int rc = kp::block_bwd(kp::make_fast_args(a), dOut, dX, dPk, dth_part, tab_part, &fgrid, st);
    if (rc) return rc;
    cudaStream_t lst = st;
    if (a.leaf_stream && a.leaf_stream != stream) {
      lst = (cudaStream_t)a.leaf_stream;
      KP_CUDA(kp::fork_stream(st, lst));
    }
    if (want_table)
      KP_LAUNCH(kp::reduce_partials_kernel, kp::ceil_div((long long)tn * 32, 256), 256, 0, lst, tab_part, fgrid, (int)tn,
                a.rows0 * a.d, dT0, dTk);
    if (geo) {
      KP_LAUNCH(kp::dtheta_geo_bwd_kernel, a.d, 256, 0, lst, dth_part, fgrid, a.k, a.d, a.geo_alphas, a.theta, dtheta,
                a.geo_dalphas);
    } else if (dtheta) {
      const int n = a.k * a.d;
      KP_LAUNCH(kp::reduce_partials_kernel, kp::ceil_div((long long)n * 32, 256), 256, 0, lst, dth_part, fgrid, n, n,
                dtheta, (float*)nullptr);
    }
    return 0;
  }
  const bool want_b1 = c.need_gs ? (dX || want_table || dP || dth_part || dep_part) : (dP || dep_part);
  if (want_b1) {
    float* dPk = dP;
    if (!a.fuse && dP == dOut) dPk = nullptr;
    if (c.fast && c.lean_b1) {
      int rc = kp::lean_b1(kp::make_fast_args(a), c.fG, a.act, a.fuse != 0, c.ftab, c.grid_b1, c.lean_b1_threads,
                           c.lean_b1_smem, dOut, Gs, dPk, dth_part, st);
      if (rc) return rc;
    } else if (c.fast) {
      int rc = kp::fast_b1(kp::make_fast_args(a), c.fG, a.act, a.fuse != 0, c.ftab, c.fextra, c.grid_b1,
                           dth_part ? c.smem_b1 : c.fsmem_fwd, dOut, Gs, dPk, dth_part, dep_part, st);
      if (rc) return rc;
    } else {
      KP_DISPATCH_VAF(kp::launch_b1, c.vec, a.act, a.fuse, args, c, dOut, Gs, dPk, dth_part, dep_part, st);
    }
  }
  // Leaf gradients (tables, dtheta / dalphas, deps) feed nothing but the optimizer: with desc.leaf_stream they are
  // forked onto that stream right after B1, so B3 and the reductions overlap B2 and whatever the caller launches next
  // on the main stream.  The caller joins the leaf stream before reading them and keeps the workspace alive until then.
  cudaStream_t lst = st;
  if (a.leaf_stream && a.leaf_stream != stream) {
    lst = (cudaStream_t)a.leaf_stream;
    KP_CUDA(kp::fork_stream(st, lst));
  }
  const float* Gsrc = c.need_gs ? Gs : dOut;
  if (dX && c.fast && kp::tile_eligible(a, kp::TAB_NONE) && !a.dx_node_stride && !a.dx_hop_stride && !a.dx_accumulate) {
    int rc = kp::tile_b2(kp::make_fast_args(a), c.fG, c.fextra, Gsrc, dOut, dX, st);
    if (rc) return rc;
  } else if (dX && c.fast) {
    // chunked call: colT holds batch-wide node ids, the chunk's Gs starts at node node_base
    const float* Gb2 = Gsrc - (size_t)a.node_base * a.k * a.d;
    int rc = kp::fast_b2(kp::make_fast_args(a), c.fG, a.fuse != 0, c.fextra, c.fgrid, Gb2, dOut, dX, st);
    if (rc) return rc;
  } else if (dX) {
    if (c.vec == 4) {
      if (a.fuse) KP_LAUNCH((kp::agg_bwd_src_kernel<4, true>), c.grid, 256, 0, st, args, Gsrc, dOut, dX);
      else        KP_LAUNCH((kp::agg_bwd_src_kernel<4, false>), c.grid, 256, 0, st, args, Gsrc, dOut, dX);
    } else if (c.vec == 2) {
      if (a.fuse) KP_LAUNCH((kp::agg_bwd_src_kernel<2, true>), c.grid, 256, 0, st, args, Gsrc, dOut, dX);
      else        KP_LAUNCH((kp::agg_bwd_src_kernel<2, false>), c.grid, 256, 0, st, args, Gsrc, dOut, dX);
    } else {
      if (a.fuse) KP_LAUNCH((kp::agg_bwd_src_kernel<1, true>), c.grid, 256, 0, st, args, Gsrc, dOut, dX);
      else        KP_LAUNCH((kp::agg_bwd_src_kernel<1, false>), c.grid, 256, 0, st, args, Gsrc, dOut, dX);
    }
  }
  if (want_table) {
    if (c.table_atomic) {
      if (dT0) KP_CUDA(cudaMemsetAsync(dT0, 0, sizeof(float) * (size_t)a.rows0 * a.d, lst));
      if (dTk) KP_CUDA(cudaMemsetAsync(dTk, 0, sizeof(float) * (size_t)a.rowsk * a.d, lst));
      KP_CHECK_ARG(dT0 && (dTk || a.k == 1), "kp_agg_backward: atomic table path needs both dT0 and dTk");
      KP_LAUNCH(kp::agg_bwd_table_atomic_kernel, kp::kNumSMs * 8, 256, 0, lst, a, Gsrc, dT0, dTk);
    } else {
      float* part = (float*)(ws + w.table);
      const int threads = 256;
      if (c.b3_count) {
        int rc = kp::b3_count(a, c.fG, Gsrc, part, dT0, dTk, lst);
        if (rc) return rc;
      } else if (c.b3_fast) {
#define KP_B3F(GG)                                                                                              \
  do {                                                                                                          \
    if (c.smem_b3 > 32 * 1024) {                                                                                \
      KP_CUDA(cudaFuncSetAttribute(kp::agg_bwd_table_fast_kernel<GG, false>,                                    \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem_b3));               \
      KP_CUDA(cudaFuncSetAttribute(kp::agg_bwd_table_fast_kernel<GG, true>,                                     \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem_b3));               \
    }                                                                                                           \
    if (a.dinv)                                                                                                 \
      KP_LAUNCH((kp::agg_bwd_table_fast_kernel<GG, true>), c.grid_b3, c.b3_threads, c.smem_b3, lst, a, Gsrc,     \
                c.b3_groups, c.b3_rows_per_group, part);                                                        \
    else                                                                                                        \
      KP_LAUNCH((kp::agg_bwd_table_fast_kernel<GG, false>), c.grid_b3, c.b3_threads, c.smem_b3, lst, a, Gsrc,    \
                c.b3_groups, c.b3_rows_per_group, part);                                                        \
  } while (0)
        if (c.b3_G == 32) KP_B3F(32);
        else if (c.b3_G == 16) KP_B3F(16);
        else if (c.b3_G == 8) KP_B3F(8);
        else KP_B3F(4);
#undef KP_B3F
      } else {
#define KP_B3(V)                                                                                           \
  do {                                                                                                     \
    if (c.smem_b3 > 32 * 1024)                                                                             \
      KP_CUDA(cudaFuncSetAttribute(kp::agg_bwd_table_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   (int)c.smem_b3));                                                       \
    KP_LAUNCH((kp::agg_bwd_table_kernel<V>), c.grid_b3, threads, c.smem_b3, lst, a, Gsrc, c.cw, c.rl,       \
              c.rows_per_block, part);                                                                     \
  } while (0)
      if (c.vec == 4) KP_B3(4);
      else if (c.vec == 2) KP_B3(2);
      else KP_B3(1);
#undef KP_B3
      }
      if (!c.b3_count)
        KP_LAUNCH(kp::reduce_partials_kernel, kp::ceil_div((long long)tn * 32, 256), 256, 0, lst, part, c.grid_b3,
                  (int)tn, a.rows0 * a.d, dT0, dTk);
    }
  }
  if (geo) {
    KP_LAUNCH(kp::dtheta_geo_bwd_kernel, a.d, 256, 0, lst, dth_part, c.grid_b1, a.k, a.d,
              a.geo_alphas, a.theta, dtheta, a.geo_dalphas);
  } else if (dtheta) {
    const int n = a.k * a.d;
    KP_LAUNCH(kp::reduce_partials_kernel, kp::ceil_div((long long)n * 32, 256), 256, 0, lst, dth_part, c.grid_b1, n, n,
              dtheta, (float*)nullptr);
  }
  if (deps) {
    KP_LAUNCH(kp::reduce_partials_kernel, 1, 32, 0, lst, dep_part, c.grid_b1, 1, 1, deps, (float*)nullptr);
  }
  return 0;
}

}  // extern "C"
