// agg_b3_count.cu -- table-gradient pass (B3) as a count-matrix product with register accumulators:
//     dT[t,:] = sum over (node,hop) rows r of  C[r,t] * Gs[r,:],      C[r,t] = sum of the weights of row r's entries
//                                                                             whose embedding row is t
// (autograd of the edge-embedding lookups, KPGIN.py:115-118 / KPGINplus.py:82-85 / KPGCN.py:120-123 / gine.py:56-59).
//
// The sub-table kernel in agg.cu scatters every entry into a private [rows x d] table in shared memory; 190 KB of
// tables cap an SM at 8 warps and ncu shows it latency-bound at 11 % of DRAM throughput (profiles/r1q_b3.txt).
// The plan knows the largest embedding row actually present (kp_plan_count stats; 4 and <= 9 on ZINC-shape data
// although the tables have 5 and 52 rows), so here
//   * a thread per row builds that row's A counts in shared memory (256 rows x (A+4) floats per tile: no atomics,
//     the owner thread adds its entries in list order);
//   * a group of G lanes streams G rows of Gs (8 rows in flight) and does A x float4 FMAs per row into REGISTER
//     accumulators -- no per-entry loop, no read-modify-write on shared memory, empty rows are never loaded;
//   * hop-0 rows (table T0) and hop >= 1 rows (table Tk) go to different CTAs, so one accumulator set serves both.
// 16 warps/SM instead of 8, ~60 instead of ~160 warp instructions per row.  Accumulation order is a function of
// (grid, tile order, row order) only -> bit-reproducible; per-CTA partial tables are summed in a fixed order.
// Eligible: d % 4 == 0, d <= 128, largest embedding row <= 31 (desc.amax0 / amaxk).  Everything else stays on agg.cu.
#include "agg_common.cuh"

namespace kp { extern int g_geom_max_ctas; }   // test hook, agg_fast_host.h

namespace kp {

constexpr int B3C_THREADS = 256;
constexpr int B3C_CTAS_PER_SM = 2;

template <int G, int A4, bool NORM>
__global__ void __launch_bounds__(B3C_THREADS, (A4 <= 4) ? B3C_CTAS_PER_SM : 1)
agg_bwd_table_count_kernel(const kp_agg_desc a, const float* __restrict__ Gs, int g0, float* __restrict__ part) {
  constexpr int A = 4 * A4;
  constexpr int S = A + 4;                                   // count row: A counts | non-empty flag | Gs row | pad
  constexpr int RB = (G < 8) ? G : ((A4 <= 2 && G >= 16) ? 16 : ((A4 <= 4) ? 8 : 4));   // rows in flight per group
  extern __shared__ __align__(16) float smem[];
  float* Cm = smem;                                          // [B3C_THREADS][S]
  float* red = smem + B3C_THREADS * S;                       // [A][d]
  const int d = a.d, k = a.k, Kp = a.Kplan;
  const int lane = threadIdx.x & (G - 1);
  const int grp = threadIdx.x / G;
  const int c = min(lane * 4, d - 4);
  const bool active = lane * 4 < d;
  const bool cls1 = (int)blockIdx.x >= g0;                   // CTA class: 0 = hop-0 rows (T0), 1 = hops >= 1 (Tk)
  const int km1 = k - 1;
  const long long Q = cls1 ? (long long)a.N * km1 : (long long)a.N;
  const int ntiles = (int)((Q + B3C_THREADS - 1) / B3C_THREADS);
  const int first = cls1 ? (int)blockIdx.x - g0 : (int)blockIdx.x;
  const int step = cls1 ? (int)gridDim.x - g0 : g0;

  float4 acc[A];
#pragma unroll
  for (int i = 0; i < A; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int t = first; t < ntiles; t += step) {
    __syncthreads();                                         // previous tile's counts fully consumed
    {
      float* row = Cm + threadIdx.x * S;
#pragma unroll
      for (int i = 0; i < A4; ++i) *reinterpret_cast<float4*>(row + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
      const long long q = (long long)t * B3C_THREADS + threadIdx.x;
      int flag = 0, r = 0;
      if (q < Q) {
        int v, h;
        if (cls1) {
          v = (int)(q / km1);
          h = 1 + (int)(q - (long long)v * km1);
        } else {
          v = (int)q;
          h = 0;
        }
        r = v * k + h;
        const int* rp = a.rowptr + (size_t)v * Kp + h;
        const int b = __ldg(rp), e = __ldg(rp + 1);
        for (int j = b; j < e; ++j) {
          const unsigned at = __ldg(a.attr16 + j);
          float w = 1.f;
          if (NORM) w = __ldg(a.dinv + (size_t)__ldg(a.col + j) * Kp + h);
          if (at < (unsigned)A) row[at] += w;              // rows above the validated maximum: see GraphPlan.validate
        }
        flag = e > b;
      }
      row[A] = __int_as_float(flag);
      row[A + 1] = __int_as_float(r);
    }
    __syncthreads();
    const float* crow = Cm + (grp * G) * S;                  // this group's G consecutive rows of the tile
#pragma unroll 1
    for (int i0 = 0; i0 < G; i0 += RB) {
      float4 g[RB];
      int fl[RB];
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const float2 meta = *reinterpret_cast<const float2*>(crow + (i0 + u) * S + A);
        fl[u] = __float_as_int(meta.x);
        g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fl[u] && active)
          g[u] = __ldcs(reinterpret_cast<const float4*>(Gs + (size_t)__float_as_int(meta.y) * d + c));
      }
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        if (fl[u]) {
#pragma unroll
          for (int i = 0; i < A4; ++i) {
            const float4 cn = *reinterpret_cast<const float4*>(crow + (i0 + u) * S + 4 * i);
#define KP_B3C_FMA(J, W)                         \
  acc[4 * i + J].x = fmaf(W, g[u].x, acc[4 * i + J].x); \
  acc[4 * i + J].y = fmaf(W, g[u].y, acc[4 * i + J].y); \
  acc[4 * i + J].z = fmaf(W, g[u].z, acc[4 * i + J].z); \
  acc[4 * i + J].w = fmaf(W, g[u].w, acc[4 * i + J].w);
            KP_B3C_FMA(0, cn.x)
            KP_B3C_FMA(1, cn.y)
            KP_B3C_FMA(2, cn.z)
            KP_B3C_FMA(3, cn.w)
#undef KP_B3C_FMA
          }
        }
      }
    }
  }
  // groups of one warp: symmetric butterfly (a+b == b+a bitwise), then the CTA's warps in order through `red`
  if (G < 32) {
#pragma unroll
    for (int i = 0; i < A; ++i) {
#pragma unroll
      for (int o = G; o < 32; o <<= 1) {
        acc[i].x += __shfl_xor_sync(0xffffffffu, acc[i].x, o);
        acc[i].y += __shfl_xor_sync(0xffffffffu, acc[i].y, o);
        acc[i].z += __shfl_xor_sync(0xffffffffu, acc[i].z, o);
        acc[i].w += __shfl_xor_sync(0xffffffffu, acc[i].w, o);
      }
    }
  }
  const int warp = threadIdx.x >> 5;
  const bool writer = active && (threadIdx.x & 31) < G;
  for (int w = 0; w < B3C_THREADS / 32; ++w) {
    __syncthreads();
    if (warp == w && writer) {
#pragma unroll
      for (int i = 0; i < A; ++i) {
        float4* p = reinterpret_cast<float4*>(red + i * d + c);
        float4 s = acc[i];
        if (w > 0) {
          const float4 o = *p;
          s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
        }
        *p = s;
      }
    }
  }
  __syncthreads();
  float* out = part + (size_t)blockIdx.x * A * d;
  for (int i = threadIdx.x * 4; i < A * d; i += B3C_THREADS * 4)
    *reinterpret_cast<float4*>(out + i) = *reinterpret_cast<const float4*>(red + i);
}

// dT0[t,c] = sum over the class-0 CTAs' partials, dTk[t,c] over the class-1 CTAs'; embedding rows >= A got no
// entries -> 0.  One warp per output element, lane l adds partials l, l+32, ... then a fixed shuffle tree.
__global__ void b3_count_reduce_kernel(const float* __restrict__ part, int g0, int g1, int A, int d, int rows0,
                                       int rowsk, float* __restrict__ dT0, float* __restrict__ dTk) {
  const int lane = threadIdx.x & 31;
  int i = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int n0 = rows0 * d, n = n0 + rowsk * d;
  if (i >= n) return;
  const bool second = i >= n0;
  if (second) i -= n0;
  const int t = i / d, cc = i - t * d;
  const float* base = part + (second ? (size_t)g0 * A * d : 0) + (size_t)t * d + cc;
  const int nb = second ? g1 : g0;
  float s = 0.f;
  if (t < A)
    for (int b = lane; b < nb; b += 32) s += __ldcs(base + (size_t)b * A * d);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    float* out = second ? dTk : dT0;
    if (out) out[i] = s;
  }
}

static int b3c_need(const kp_agg_desc& a) {
  int m = a.amax0;
  if (a.k > 1 && a.amaxk > m) m = a.amaxk;
  return m + 1;
}

static int b3c_A4(const kp_agg_desc& a) {
  const int need = b3c_need(a);
  return need <= 8 ? 2 : (need <= 12 ? 3 : (need <= 16 ? 4 : 8));
}

bool b3_count_ok(const kp_agg_desc& a, int G) {
  if (!a.T0 || a.d % 4 || a.d > 128 || a.d < 4) return false;
  if (G != 32 && G != 16 && G != 8 && G != 4) return false;
  if (G * 4 < a.d) return false;
  if (a.amax0 < 0 || (a.k > 1 && a.amaxk < 0)) return false;          // plan statistics not supplied
  if (b3c_need(a) > 32) return false;
  if ((long long)a.N * a.k >= 0x7fffffffLL) return false;
  return true;
}

// CTAs of class 0 (hop-0 rows) and class 1 (hops >= 1), in proportion to their tiles
void b3_count_grid(const kp_agg_desc& a, int* g0, int* g1) {
  const long long t0 = ((long long)a.N + B3C_THREADS - 1) / B3C_THREADS;
  const long long t1 = ((long long)a.N * (a.k - 1) + B3C_THREADS - 1) / B3C_THREADS;
  long long cap = (long long)kNumSMs * (b3c_A4(a) <= 4 ? B3C_CTAS_PER_SM : 1);
  if (g_geom_max_ctas > 0 && cap > g_geom_max_ctas) cap = g_geom_max_ctas < 2 ? 2 : g_geom_max_ctas;   // test hook
  if (t0 + t1 <= cap) {
    *g0 = (int)t0;
    *g1 = (int)t1;
    return;
  }
  long long c0 = (cap * t0 + (t0 + t1) / 2) / (t0 + t1);
  if (c0 < 1) c0 = 1;
  if (c0 > t0) c0 = t0;
  long long c1 = cap - c0;
  if (c1 > t1) c1 = t1;
  *g0 = (int)c0;
  *g1 = (int)c1;
}

size_t b3_count_part_floats(const kp_agg_desc& a) {
  int g0, g1;
  b3_count_grid(a, &g0, &g1);
  return (size_t)(g0 + g1) * 4 * b3c_A4(a) * a.d;
}

template <int G, int A4>
static int b3c_launch(const kp_agg_desc& a, const float* Gs, int g0, int g1, float* part, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)B3C_THREADS * (4 * A4 + 4) + (size_t)4 * A4 * a.d);
  if (a.dinv) {
    if (smem > 32 * 1024)
      KP_CUDA(cudaFuncSetAttribute(agg_bwd_table_count_kernel<G, A4, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KP_LAUNCH((agg_bwd_table_count_kernel<G, A4, true>), g0 + g1, B3C_THREADS, smem, st, a, Gs, g0, part);
  } else {
    if (smem > 32 * 1024)
      KP_CUDA(cudaFuncSetAttribute(agg_bwd_table_count_kernel<G, A4, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KP_LAUNCH((agg_bwd_table_count_kernel<G, A4, false>), g0 + g1, B3C_THREADS, smem, st, a, Gs, g0, part);
  }
  return 0;
}

template <int G>
static int b3c_launch_g(const kp_agg_desc& a, const float* Gs, int g0, int g1, float* part, cudaStream_t st) {
  switch (b3c_A4(a)) {
    case 2: return b3c_launch<G, 2>(a, Gs, g0, g1, part, st);
    case 3: return b3c_launch<G, 3>(a, Gs, g0, g1, part, st);
    case 4: return b3c_launch<G, 4>(a, Gs, g0, g1, part, st);
    default: return b3c_launch<G, 8>(a, Gs, g0, g1, part, st);
  }
}

int b3_count(const kp_agg_desc& a, int G, const float* Gs, float* part, float* dT0, float* dTk, cudaStream_t st) {
  int g0, g1;
  b3_count_grid(a, &g0, &g1);
  int rc;
  if (G == 32) rc = b3c_launch_g<32>(a, Gs, g0, g1, part, st);
  else if (G == 16) rc = b3c_launch_g<16>(a, Gs, g0, g1, part, st);
  else if (G == 8) rc = b3c_launch_g<8>(a, Gs, g0, g1, part, st);
  else rc = b3c_launch_g<4>(a, Gs, g0, g1, part, st);
  if (rc) return rc;
  const int A = 4 * b3c_A4(a);
  const long long n = (long long)(a.rows0 + (dTk ? a.rowsk : 0)) * a.d;
  KP_LAUNCH(b3_count_reduce_kernel, ceil_div(n * 32, 256), 256, 0, st, part, g0, g1, A, a.d, a.rows0,
            dTk ? a.rowsk : 0, dT0, dTk);
  return 0;
}

}  // namespace kp
