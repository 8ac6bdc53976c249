// agg_b3_match.cu -- table-gradient pass (B3) at full occupancy: dT[t,:] = sum over entries with table row t of
// Gs[row(entry),:]   (autograd of the edge-embedding lookups, KPGIN.py:115-118 / KPGINplus.py:82-85 / gine.py:56-59).
//
// The sub-table kernel in agg.cu keeps private [rows x d] accumulators in shared memory, which caps an SM at ~8
// warps; ncu (profiles/r1q_b3.txt) shows it latency-bound at 11 % of DRAM throughput.  Here the accumulators live in
// REGISTERS: warp w of a 1024-thread CTA owns table rows w and w+32 (lane l holds columns 4l..4l+3 of both), so
// there is no read-modify-write on shared memory at all.  Per tile of consecutive nodes the CTA
//   1. streams the tile's Gs rows (contiguous, <= 52 KB) into shared memory with cp.async, double-buffered;
//   2. builds the tile's entry list {table row, local Gs row} from the plan (warp per node);
//   3. every warp scans the list 32 entries at a time, ballots "is this one of my rows" and adds the matching Gs
//      rows from shared memory, in list order.
// Accumulation order is fixed by (grid size, tile order, entry order) -> bit-reproducible; per-CTA partial tables
// go through the same fixed-order reduce_partials kernel as before.
// Eligible: 68 <= d <= 128, d % 4 == 0, rows0 + rowsk <= 64, no per-entry norm.  Everything else stays on agg.cu.
#include "agg_common.cuh"

namespace kp {

constexpr int B3M_THREADS = 1024;
constexpr int B3M_TILE_FLOATS = 13312;      // 52 KB per Gs buffer (two buffers)
constexpr int B3M_ECAP = 2048;              // entry slots per pass

__device__ __forceinline__ void b3m_cp16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(B3M_THREADS, 1)
agg_bwd_table_match_kernel(const kp_agg_desc a, const float* __restrict__ Gs, int nodes_per_tile, int ntiles,
                           float* __restrict__ part) {
  extern __shared__ __align__(16) float smem[];
  float* tile[2] = {smem, smem + B3M_TILE_FLOATS};
  unsigned* ents = reinterpret_cast<unsigned*>(smem + 2 * B3M_TILE_FLOATS);   // [B3M_ECAP] (table row << 16) | local row
  const int d = a.d, k = a.k, Kp = a.Kplan;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = min(lane * 4, d - 4);
  const bool active = lane * 4 < d;
  const int trows = a.rows0 + a.rowsk;
  const unsigned p0 = (unsigned)warp, p1 = (unsigned)warp + 32u;
  // two accumulators per owned row (even / odd matches): halves the dependent-add chain of a hot row
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0, acc0b = acc0, acc1b = acc0;
  // matches are taken four at a time: the four shuffles and the four shared-memory row loads are independent
  auto consume = [&](unsigned m, unsigned en, const float* G, float4& ea, float4& eb) {
    while (m) {
      unsigned row[4];
      bool ok[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ok[q] = m != 0u;
        const int bit = ok[q] ? __ffs(m) - 1 : 0;
        m &= m - 1u;
        row[q] = __shfl_sync(0xffffffffu, en, bit) & 0xffffu;
      }
      float4 g[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        g[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[q]) g[q] = *reinterpret_cast<const float4*>(G + row[q] * d);
      }
      ea.x += g[0].x; ea.y += g[0].y; ea.z += g[0].z; ea.w += g[0].w;
      eb.x += g[1].x; eb.y += g[1].y; eb.z += g[1].z; eb.w += g[1].w;
      ea.x += g[2].x; ea.y += g[2].y; ea.z += g[2].z; ea.w += g[2].w;
      eb.x += g[3].x; eb.y += g[3].y; eb.z += g[3].z; eb.w += g[3].w;
    }
  };

  auto issue = [&](int t, int buf) {                 // cp.async the Gs rows of tile t
    const int v0 = t * nodes_per_tile;
    const int nv = min(nodes_per_tile, a.N - v0);
    const int n4 = nv * k * d / 4;
    const float4* src = reinterpret_cast<const float4*>(Gs + (size_t)v0 * k * d);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(tile[buf]);
    for (int i = threadIdx.x; i < n4; i += B3M_THREADS) b3m_cp16(dst + 16u * i, src + i);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int t = blockIdx.x, buf = 0;
  if (t < ntiles) issue(t, 0);
  for (; t < ntiles; t += gridDim.x, buf ^= 1) {
    const int tn = t + gridDim.x;
    if (tn < ntiles) issue(tn, buf ^ 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const int v0 = t * nodes_per_tile;
    const int nv = min(nodes_per_tile, a.N - v0);
    const int j0 = __ldg(a.rowptr + (size_t)v0 * Kp);
    const int j1 = __ldg(a.rowptr + (size_t)(v0 + nv - 1) * Kp + k);      // end of the last node's used hops
    asm volatile("cp.async.wait_group 1;" ::: "memory");                  // this tile's rows have landed
    const float* G = tile[buf] + c;
    for (int s0 = j0; s0 < j1; s0 += B3M_ECAP) {                           // usually one pass
      const int s1 = min(j1, s0 + B3M_ECAP);
      __syncthreads();                                                     // previous pass / tile fully consumed
      for (int i = threadIdx.x; i < s1 - s0; i += B3M_THREADS) ents[i] = 0xffffffffu;   // slots of unused hops
      __syncthreads();
      for (int vl = warp; vl < nv; vl += B3M_THREADS / 32) {               // warp per node: fill its hops' slots
        const int* rp = a.rowptr + (size_t)(v0 + vl) * Kp;
        int b = __ldg(rp);
        for (int h = 0; h < k; ++h) {
          const int e = __ldg(rp + h + 1);
          const unsigned base = (h == 0) ? 0u : (unsigned)a.rows0;
          const unsigned lrow = (unsigned)(vl * k + h);
          for (int j = max(b, s0) + lane; j < min(e, s1); j += 32)
            ents[j - s0] = ((base + (unsigned)__ldg(a.attr16 + j)) << 16) | lrow;
          b = e;
        }
      }
      __syncthreads();
      const int ne = s1 - s0;
      for (int i0 = 0; i0 < ne; i0 += 32) {
        const unsigned en = (i0 + lane < ne) ? ents[i0 + lane] : 0xffffffffu;
        const unsigned key = en >> 16;
        unsigned m0 = __ballot_sync(0xffffffffu, key == p0);
        unsigned m1 = __ballot_sync(0xffffffffu, key == p1);
        consume(m0, en, G, acc0, acc0b);
        consume(m1, en, G, acc1, acc1b);
      }
    }
    __syncthreads();                                                       // tile[buf] may be refilled next round
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  float* out = part + (size_t)blockIdx.x * trows * d + c;
  acc0.x += acc0b.x; acc0.y += acc0b.y; acc0.z += acc0b.z; acc0.w += acc0b.w;
  acc1.x += acc1b.x; acc1.y += acc1b.y; acc1.z += acc1b.z; acc1.w += acc1b.w;
  if (active) {
    if ((int)p0 < trows) *reinterpret_cast<float4*>(out + (size_t)p0 * d) = acc0;
    if ((int)p1 < trows) *reinterpret_cast<float4*>(out + (size_t)p1 * d) = acc1;
  }
}

bool b3_match_ok(const kp_agg_desc& a, const float* Gs) {
  return a.d % 4 == 0 && a.d > 64 && a.d <= 128 && a.rows0 + a.rowsk <= 64 && !a.dinv && a.T0 &&
         (((uintptr_t)Gs) & 15) == 0 && (long long)a.k * a.d <= B3M_TILE_FLOATS &&
         (long long)(B3M_TILE_FLOATS / (a.k * a.d)) * a.k <= 65535;
}

int b3_match_grid(const kp_agg_desc& a) {
  const int npt = B3M_TILE_FLOATS / (a.k * a.d);
  const int ntiles = (a.N + npt - 1) / npt;
  return ntiles < kNumSMs ? (ntiles < 1 ? 1 : ntiles) : kNumSMs;
}

int b3_match(const kp_agg_desc& a, const float* Gs, float* part, int grid, cudaStream_t st) {
  const int npt = B3M_TILE_FLOATS / (a.k * a.d);
  const int ntiles = (a.N + npt - 1) / npt;
  const size_t smem = sizeof(float) * 2 * B3M_TILE_FLOATS + sizeof(unsigned) * B3M_ECAP;
  KP_CUDA(cudaFuncSetAttribute(agg_bwd_table_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KP_LAUNCH(agg_bwd_table_match_kernel, grid, B3M_THREADS, smem, st, a, Gs, npt, ntiles, part);
  return 0;
}

}  // namespace kp
