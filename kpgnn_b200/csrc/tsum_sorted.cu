// tsum_sorted.cu -- gradient of the multi-table gather-sum (kp_table_sum_backward) by tile-local counting sort:
//     dTable[t,:] = sum_{(r,s): slot_off[s]+idx[r,s] == t} dOut[r,:]
// (autograd of the peripheral-attribute / input embedding lookups, models/GNNs.py:393-400, feature_encoder.py:37-67).
//
// The sub-table kernel in tsum.cu privatises [rows x d] tables in shared memory: 8 groups x 23 KB fill an SM, every
// group walks ~170 rows one L2 round trip per 8 rows, and consecutive rows that hit the same table row (index 0 is by
// far the most common) serialise on a shared-memory read-modify-write: 150 us for a 128-molecule batch, the largest
// kernel of the training step (profiles/r1y_step_kineto.txt).  Here
//   * a CTA stages a tile of <= 192 rows of dOut in shared memory with ONE batch of cp.async (one round trip);
//   * each warp takes one embedding table, counting-sorts the tile's (row, slot) entries of that table by index
//     (match_any + popc ranks: no atomics, positions are a pure function of the data) and
//   * sums every non-empty bin's rows from shared memory into REGISTERS, in sorted order, writing one partial row per
//     (CTA, table row) plus a presence flag; a second kernel adds the partials of the CTAs that have the row, in CTA
//     order.  Fixed order everywhere -> bit-reproducible; no float atomics.
// Eligible: d % 4 == 0, 64 <= d <= 128, every table <= 256 rows, <= 32 tables.  Everything else stays on tsum.cu.
#include "agg_common.cuh"

namespace kp {

constexpr int TS_THREADS = 512;
constexpr int TS_WARPS = TS_THREADS / 32;
constexpr int TS_MAXT = 32;          // tables
constexpr int TS_MAXROWS = 256;      // rows per table

struct TsTables {
  int ntables;
  int row0[TS_MAXT + 1];             // first table row of table j
  int slot0[TS_MAXT + 1];            // first slot of table j
};

__device__ __forceinline__ void ts_cp16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}

__global__ void __launch_bounds__(TS_THREADS, 1)
tsum_bwd_sorted_kernel(const kp_tsum_desc t, const TsTables tb, const float* __restrict__ dOut, int TR, int max_ns,
                       float* __restrict__ part, unsigned char* __restrict__ flags) {
  extern __shared__ __align__(16) float smem[];
  const int d = t.d, S = t.S;
  const int r0 = blockIdx.x * TR;
  const int nr = min(TR, t.R - r0);
  float* G = smem;                                                        // [TR][d]
  int* hist = reinterpret_cast<int*>(G + (size_t)TR * d);                 // [TS_WARPS][TS_MAXROWS]
  unsigned short* ix = reinterpret_cast<unsigned short*>(hist + TS_WARPS * TS_MAXROWS);   // [S][TR]
  unsigned short* perm = ix + (size_t)S * TR;                             // [TS_WARPS][TR * max_ns]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const float* src = dOut + (size_t)r0 * d;
    for (int i = threadIdx.x * 4; i < nr * d; i += TS_THREADS * 4) ts_cp16(G + i, src + i);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int e = threadIdx.x; e < nr * S; e += TS_THREADS) {                // coalesced int64 reads, transposed stores
    const int r = e / S, s = e - r * S;
    ix[s * TR + r] = (unsigned short)__ldg(t.idx + (size_t)r0 * S + e);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int c = min(lane * 4, d - 4);
  const bool active = lane * 4 < d;
  int* h = hist + warp * TS_MAXROWS;
  unsigned short* pm = perm + (size_t)warp * TR * max_ns;
  const unsigned lt = (1u << lane) - 1u;
  for (int j = warp; j < tb.ntables; j += TS_WARPS) {
    const int n_i = tb.row0[j + 1] - tb.row0[j];
    const int s0 = tb.slot0[j], ns = tb.slot0[j + 1] - s0;
    for (int i = lane; i < n_i; i += 32) h[i] = 0;
    __syncwarp();
    // pass 1: bin sizes
    for (int sl = 0; sl < ns; ++sl)
      for (int rb = 0; rb < nr; rb += 32) {
        const int r = rb + lane;
        const int v = r < nr ? min((int)ix[(s0 + sl) * TR + r], n_i - 1) : -1;
        const unsigned m = __match_any_sync(0xffffffffu, v);
        if (v >= 0 && (m & lt) == 0u) h[v] += __popc(m);
        __syncwarp();
      }
    // exclusive scan of the bin sizes -> bin starts
    int carry = 0;
    for (int b0 = 0; b0 < n_i; b0 += 32) {
      const int x = b0 + lane < n_i ? h[b0 + lane] : 0;
      int inc = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
      }
      if (b0 + lane < n_i) h[b0 + lane] = carry + inc - x;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    __syncwarp();
    // pass 2: positions (slot-major, then row order inside a bin); h[i] ends as the bin's end
    for (int sl = 0; sl < ns; ++sl)
      for (int rb = 0; rb < nr; rb += 32) {
        const int r = rb + lane;
        const int v = r < nr ? min((int)ix[(s0 + sl) * TR + r], n_i - 1) : -1;
        const unsigned m = __match_any_sync(0xffffffffu, v);
        if (v >= 0) pm[h[v] + __popc(m & lt)] = (unsigned short)r;
        __syncwarp();
        if (v >= 0 && (m & lt) == 0u) h[v] += __popc(m);
        __syncwarp();
      }
    // bins -> register sums -> partial rows
    float* prow = part + ((size_t)blockIdx.x * t.table_rows + tb.row0[j]) * d + c;
    unsigned char* frow = flags + (size_t)blockIdx.x * t.table_rows + tb.row0[j];
    int b = 0;
    for (int i = 0; i < n_i; ++i) {
      const int e = h[i];
      if (e > b) {
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        int q = b;
        for (; q + 2 <= e; q += 2) {
          const float4 x0 = *reinterpret_cast<const float4*>(G + (size_t)pm[q] * d + c);
          const float4 x1 = *reinterpret_cast<const float4*>(G + (size_t)pm[q + 1] * d + c);
          a0.x += x0.x; a0.y += x0.y; a0.z += x0.z; a0.w += x0.w;
          a1.x += x1.x; a1.y += x1.y; a1.z += x1.z; a1.w += x1.w;
        }
        if (q < e) {
          const float4 x0 = *reinterpret_cast<const float4*>(G + (size_t)pm[q] * d + c);
          a0.x += x0.x; a0.y += x0.y; a0.z += x0.z; a0.w += x0.w;
        }
        a0.x += a1.x; a0.y += a1.y; a0.z += a1.z; a0.w += a1.w;
        if (active) __stcg(reinterpret_cast<float4*>(prow + (size_t)i * d), a0);
      }
      if (lane == 0) frow[i] = e > b ? 1 : 0;
      b = e;
    }
    __syncwarp();
  }
}

// dTable[t, :] = sum over the CTAs that have row t (flag), in CTA order.  One 256-thread CTA per table row: warp w
// takes the w-th eighth of the producer CTAs, 8 independent loads in flight per lane (rows without the flag are not
// read and count as zero); the 8 warp sums are then added in warp order.  A warp-per-row loop over ~120 producers
// was one dependent L2 round trip per pair of partials: 60 us.
__global__ void __launch_bounds__(256)
tsum_sorted_reduce_kernel(const float* __restrict__ part, const unsigned char* __restrict__ flags, int nctas,
                          int table_rows, int d, float* __restrict__ dTable) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x;
  const int c = min(lane * 4, d - 4);
  const int chunk = (nctas + 7) >> 3;
  const int q0 = warp * chunk, q1 = min(nctas, q0 + chunk);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int qb = q0; qb < q1; qb += 8) {
    float4 x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int q = qb + u;
      if (q < q1 && flags[(size_t)q * table_rows + row] != 0)
        x[u] = __ldcg(reinterpret_cast<const float4*>(part + ((size_t)q * table_rows + row) * d + c));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w;
    }
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && lane * 4 < d) {
    float4 s = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      s.x += red[w][lane].x; s.y += red[w][lane].y; s.z += red[w][lane].z; s.w += red[w][lane].w;
    }
    *reinterpret_cast<float4*>(dTable + (size_t)row * d + c) = s;
  }
}

struct TsSortedCfg {
  TsTables tb;
  int TR, max_ns, grid;
  size_t smem, part_bytes, flag_bytes;
};

static bool ts_sorted_config(const kp_tsum_desc& t, TsSortedCfg* c) {
  static const bool off = getenv("KP_TSUM_SORTED") && atoi(getenv("KP_TSUM_SORTED")) == 0;
  if (off || t.d % 4 || t.d < 64 || t.d > 128 || t.S < 1 || t.S > 32 || t.R < 1) return false;
  TsTables& tb = c->tb;
  tb.ntables = 0;
  int max_ns = 0;
  for (int s = 0; s < t.S; ++s) {
    if (s == 0 || t.slot_off[s] != t.slot_off[s - 1]) {
      if (s > 0 && t.slot_off[s] < t.slot_off[s - 1]) return false;
      if (tb.ntables == TS_MAXT) return false;
      tb.row0[tb.ntables] = t.slot_off[s];
      tb.slot0[tb.ntables] = s;
      ++tb.ntables;
    }
  }
  tb.row0[tb.ntables] = t.table_rows;
  tb.slot0[tb.ntables] = t.S;
  if (tb.row0[0] != 0) return false;
  for (int j = 0; j < tb.ntables; ++j) {
    const int n = tb.row0[j + 1] - tb.row0[j], ns = tb.slot0[j + 1] - tb.slot0[j];
    if (n < 1 || n > TS_MAXROWS) return false;
    if (ns > max_ns) max_ns = ns;
  }
  int TR = (int)((96 * 1024) / (sizeof(float) * t.d));
  TR -= TR % 32;
  if (TR > 256) TR = 256;
  if (TR > 192) TR = 192;
  // small inputs: spread over the SMs
  while (TR > 32 && (long long)(t.R + TR - 1) / TR < kNumSMs / 2) TR -= 32;
  c->TR = TR;
  c->max_ns = max_ns;
  c->grid = (t.R + TR - 1) / TR;
  c->smem = sizeof(float) * (size_t)TR * t.d + sizeof(int) * TS_WARPS * TS_MAXROWS +
            sizeof(unsigned short) * ((size_t)t.S * TR + (size_t)TS_WARPS * TR * max_ns);
  c->smem = (c->smem + 15) & ~(size_t)15;
  if (c->smem > 220 * 1024) return false;
  c->part_bytes = sizeof(float) * (size_t)c->grid * t.table_rows * t.d;
  c->flag_bytes = ((size_t)c->grid * t.table_rows + 255) & ~(size_t)255;
  return true;
}

// returns 0 and sets *bytes = 0 when the sorted kernel does not apply
size_t ts_sorted_workspace_bytes(const kp_tsum_desc& t) {
  TsSortedCfg c;
  if (!ts_sorted_config(t, &c)) return 0;
  return c.part_bytes + c.flag_bytes;
}

// returns -1 when not applicable, otherwise the launch status
int ts_sorted_backward(const kp_tsum_desc& t, const float* dOut, float* dTable, void* workspace, size_t workspace_bytes,
                       cudaStream_t st) {
  TsSortedCfg c;
  if (!ts_sorted_config(t, &c)) return -1;
  if (workspace_bytes < c.part_bytes + c.flag_bytes || (((uintptr_t)dOut | (uintptr_t)workspace | (uintptr_t)dTable) & 15))
    return -1;
  float* part = (float*)workspace;
  unsigned char* flags = (unsigned char*)workspace + c.part_bytes;
  KP_CUDA(cudaFuncSetAttribute(tsum_bwd_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
  KP_LAUNCH(tsum_bwd_sorted_kernel, c.grid, TS_THREADS, c.smem, st, t, c.tb, dOut, c.TR, c.max_ns, part, flags);
  KP_LAUNCH(tsum_sorted_reduce_kernel, t.table_rows, 256, 0, st, part, flags, c.grid, t.table_rows, t.d, dTable);
  return 0;
}

}  // namespace kp
