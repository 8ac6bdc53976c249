// plan.cu -- graph-plan builder: int64 hop-labelled edge list -> int32 (dst,hop)-CSR + (src,hop)-CSR.
//
// Replaces the per-layer re-derivation PyG's propagate does from edge_index/edge_attr in the reference
// (layers/KPGIN.py:100, KPGINplus.py:74, KPGCN.py:85-110, KPGraphSAGE.py:86, gine.py:52): built once per batch,
// shared by every layer's forward and backward.  HBM-bound integer work: one coalesced pass over the [E,K]
// int64 attr matrix per phase, integer atomics only (exact, order-independent), per-row ordering restored by a
// segment sort so that the float summation order downstream is reproducible.
#include <cub/device/device_scan.cuh>

#include <stdlib.h>

#include "common.cuh"

namespace kp {

__global__ void plan_count_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                  const int64_t* __restrict__ attr, int64_t attr_stride, int N, int E, int K,
                                  int* __restrict__ cnt, int* __restrict__ cntT, int* __restrict__ indeg,
                                  int* __restrict__ stats) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)E * K;
  int max0 = 0, maxk = 0, bad = 0;
  if (t < total) {
    const unsigned tu = (unsigned)t;                 // E * K < 2^31 (kp_plan_workspace_bytes): 32-bit division
    int e = (int)(tu / (unsigned)K), h = (int)(tu - (unsigned)e * (unsigned)K);
    long long s = src[e], d = dst[e];
    bool ok = (s >= 0 && s < N && d >= 0 && d < N);
    if (!ok) {
      bad = (h == 0);
    } else {
      if (h == 0) atomicAdd(&indeg[d], 1);
      long long a = attr[(long long)e * attr_stride + h];
      if (a != 0) {
        if (a < 0 || a > 65535) {
          bad = 1;
        } else {
          atomicAdd(&cnt[(long long)d * K + h], 1);
          atomicAdd(&cntT[(long long)s * K + h], 1);
          if (h == 0) max0 = (int)a; else maxk = (int)a;
        }
      }
    }
  }
  // Statistics: warp shuffle -> shared-memory atomics -> ONE guarded global update per CTA.  Every warp issuing its
  // own atomicMax on stats[1..2] serialised ~30 k same-address reductions in L2: 12 of this kernel's 17 us at the
  // bench batch, during which the step's encoder branch could not start (profiles/r1zzz_small_batch.txt).
  __shared__ int s_stat[3];
  if (threadIdx.x < 3) s_stat[threadIdx.x] = 0;
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) {
    max0 = max(max0, __shfl_xor_sync(0xffffffffu, max0, o));
    maxk = max(maxk, __shfl_xor_sync(0xffffffffu, maxk, o));
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (max0) atomicMax(&s_stat[0], max0);
    if (maxk) atomicMax(&s_stat[1], maxk);
    if (bad) atomicAdd(&s_stat[2], bad);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // the plain read only filters: a stale (smaller) value costs one redundant atomic, never a lost update
    if (s_stat[0] > __ldcg(&stats[1])) atomicMax(&stats[1], s_stat[0]);
    if (s_stat[1] > __ldcg(&stats[2])) atomicMax(&stats[2], s_stat[1]);
    if (s_stat[2]) atomicAdd(&stats[3], s_stat[2]);
  }
}

__global__ void plan_add_loops_kernel(int* __restrict__ cnt, int* __restrict__ cntT, int rows) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) {
    cnt[r] += 1;
    cntT[r] += 1;
  }
}

__global__ void plan_nnz_kernel(const int* __restrict__ rowptr, int rows, int* __restrict__ stats) {
  stats[0] = rowptr[rows];
}

// ---- small-batch head of the plan build: at a few thousand nodes the count phase was seven launches (4 memsets, the
// count kernel, 2 x {CUB scan init + scan}, the nnz read) whose launch gaps sat on the critical path of the training
// step.  One kernel zeroes all four arrays; one kernel (a CTA per array) does both exclusive scans in place and writes nnz.
__global__ void __launch_bounds__(256) plan_zero_kernel(int* __restrict__ a, int* __restrict__ b, long long nab,
                                                        int* __restrict__ c, long long nc, int* __restrict__ stats) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nab; i += stride) {
    a[i] = 0;
    b[i] = 0;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += stride) c[i] = 0;
  if (blockIdx.x == 0 && threadIdx.x < 4) stats[threadIdx.x] = 0;
}

constexpr int PLAN_SCAN_THREADS = 1024;
constexpr int PLAN_SCAN_MAX = PLAN_SCAN_THREADS * 64;        // elements one CTA scans (64 per thread)

// in-place exclusive scan of n ints by ONE CTA: thread t owns the contiguous chunk [t*per, (t+1)*per)
__global__ void __launch_bounds__(PLAN_SCAN_THREADS)
plan_scan2_kernel(int* __restrict__ a0, int* __restrict__ a1, int n, int* __restrict__ stats) {
  __shared__ int wsum[PLAN_SCAN_THREADS / 32];
  int* a = blockIdx.x == 0 ? a0 : a1;
  const int per = (n + PLAN_SCAN_THREADS - 1) / PLAN_SCAN_THREADS;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += a[i];
  // block-wide exclusive scan of the per-thread sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    wsum[lane] = wi - w;                                       // exclusive prefix of the warp sums
  }
  __syncthreads();
  int run = wsum[warp] + incl - s;
  for (int i = lo; i < hi; ++i) {
    const int v = a[i];
    a[i] = run;
    run += v;
  }
  // the last element's output is the grand total when its input is 0 (the arrays carry one slot past the rows)
  if (blockIdx.x == 0 && hi == n && lo < n) stats[0] = a[n - 1];
}

__global__ void plan_scatter_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                    const int64_t* __restrict__ attr, int64_t attr_stride, int N, int E, int K,
                                    const int* __restrict__ rowptr, const int* __restrict__ rowptrT,
                                    int* __restrict__ cur, int* __restrict__ curT, int* __restrict__ eid,
                                    int* __restrict__ eidT, int cap) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)E * K) return;
  const unsigned tu = (unsigned)t;
  int e = (int)(tu / (unsigned)K), h = (int)(tu - (unsigned)e * (unsigned)K);
  long long s = src[e], d = dst[e];
  if (s < 0 || s >= N || d < 0 || d >= N) return;
  long long a = attr[(long long)e * attr_stride + h];
  if (a <= 0 || a > 65535) return;
  long long r = d * K + h, rT = s * K + h;
  int p = rowptr[r] + atomicAdd(&cur[r], 1);
  int pT = rowptrT[rT] + atomicAdd(&curT[rT], 1);
  if (p < cap) eid[p] = e;
  if (pT < cap) eidT[pT] = e;
}

__device__ __forceinline__ void insertion_sort(int* a, int n) {
  for (int i = 1; i < n; ++i) {
    int key = a[i], j = i - 1;
    while (j >= 0 && a[j] > key) {
      a[j + 1] = a[j];
      --j;
    }
    a[j + 1] = key;
  }
}

#define KP_EMIT_LOCAL 32
// Long rows (more than KP_EMIT_LOCAL entries: every row of the far hops of an n = 1 280 regular graph) are ordered by the
// whole warp: ids staged in shared memory, every lane ranks its entries by counting the smaller ids (edge ids are
// distinct, so the ranks are a permutation) and emits straight to the rank's slot.  One thread insertion-sorting a
// 96-entry row in global memory cost 0.72 ms per 16 such graphs (profiles/r2_extract_kineto.txt).
#define KP_EMIT_WARP_CAP 1024
__device__ __forceinline__ void emit_long_row(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                              const int64_t* __restrict__ attr, int64_t attr_stride, int h,
                                              bool transposed, int* __restrict__ ids_g, int b, int n,
                                              int* __restrict__ col, uint16_t* __restrict__ attr16,
                                              int* __restrict__ colT, int* __restrict__ buf, int lane) {
  const bool staged = n <= KP_EMIT_WARP_CAP;
  const int* ids = ids_g;
  if (staged) {
    for (int i = lane; i < n; i += 32) buf[i] = ids_g[i];
    ids = buf;
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const int my = ids[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += ids[j] < my;
    if (!transposed) {
      col[b + rank] = (int)src[my];
      attr16[b + rank] = (uint16_t)attr[(long long)my * attr_stride + h];
    } else {
      colT[b + rank] = (int)dst[my];
    }
  }
  __syncwarp();
}

// two threads per row (even: the (dst,hop) row, odd: the (src,hop) row of the transpose): restore ascending edge id
// inside the row, then emit the compact arrays
__global__ void __launch_bounds__(128)
plan_emit_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                 const int64_t* __restrict__ attr, int64_t attr_stride, int N, int K,
                                 int self_loops, const int* __restrict__ rowptr, const int* __restrict__ rowptrT,
                                 int* __restrict__ eid, int* __restrict__ eidT, int* __restrict__ col,
                                 uint16_t* __restrict__ attr16, int* __restrict__ colT,
                                 float* __restrict__ dinv, int cap) {
  __shared__ int s_buf[4][KP_EMIT_WARP_CAP];
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int r = (int)(t >> 1);
  const bool transposed = t & 1u;
  const bool valid = r < N * K;
  int v = 0, h = 0, b = 0, e = 0, n = 0;
  bool cut = false;
  if (valid) {
    v = r / K;
    h = r - v * K;
    if (dinv && !transposed) dinv[r] = 1.0f / sqrtf((float)(rowptr[r + 1] - rowptr[r]));
    // Overflow (in-place refresh of a plan whose capacity the new batch exceeds; reported through stats[0] > capacity):
    // rows are cut at the capacity, so that after kp_plan_clamp every index a consumer can reach was written from THIS
    // batch (wrong results, raised by the caller at its next sync point, but never an out-of-range gather).
    b = transposed ? rowptrT[r] : rowptr[r];
    e = transposed ? rowptrT[r + 1] : rowptr[r + 1];
    cut = e > cap;
    n = b >= cap ? 0 : (cut ? cap : e) - b - ((self_loops && !cut) ? 1 : 0);
  }
  int* ids_g = (transposed ? eidT : eid) + b;
  if (valid && n > 0 && n <= KP_EMIT_LOCAL) {
    int buf[KP_EMIT_LOCAL];
#pragma unroll 4
    for (int i = 0; i < n; ++i) buf[i] = ids_g[i];
    insertion_sort(buf, n);
    if (!transposed) {
#pragma unroll 4
      for (int i = 0; i < n; ++i) {
        const int id = buf[i];
        col[b + i] = (int)src[id];
        attr16[b + i] = (uint16_t)attr[(long long)id * attr_stride + h];
      }
    } else {
#pragma unroll 4
      for (int i = 0; i < n; ++i) colT[b + i] = (int)dst[buf[i]];
    }
  }
  if (valid && self_loops && !cut && b < cap) {
    if (!transposed) {
      col[e - 1] = v;
      attr16[e - 1] = 1;
    } else {
      colT[e - 1] = v;
    }
  }
  // long rows of this warp's 32 tasks, one after the other, by the whole warp
  unsigned longmask = __ballot_sync(0xffffffffu, valid && n > KP_EMIT_LOCAL);
  while (longmask) {
    const int l = __ffs(longmask) - 1;
    longmask &= longmask - 1;
    const int lb = __shfl_sync(0xffffffffu, b, l), ln = __shfl_sync(0xffffffffu, n, l);
    const int lh = __shfl_sync(0xffffffffu, h, l);
    const bool lt = __shfl_sync(0xffffffffu, (int)transposed, l) != 0;
    emit_long_row(src, dst, attr, attr_stride, lh, lt, (lt ? eidT : eid) + lb, lb, ln, col, attr16, colT,
                  s_buf[threadIdx.x >> 5], lane);
  }
}

// In-place refresh with a fixed capacity (CUDA-graph replay loops): when a batch has more entries than were
// allocated, the rows past the capacity were not emitted; clamp the row pointers so that every consumer kernel sees
// them as EMPTY rows and stays inside col / attr16 / colT.  The overflow itself is reported through stats[0] > capacity.
__global__ void plan_clamp_kernel(int* __restrict__ rowptr, int* __restrict__ rowptrT, int n, int cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    if (rowptr[i] > cap) rowptr[i] = cap;
    if (rowptrT[i] > cap) rowptrT[i] = cap;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Closed node blocks: maximal consecutive node ranges [b, e) such that every in- and out-neighbour (over all hops)
// of a node in the range lies in the range -- for a collated batch these are the graphs (or their connected
// components), WITHOUT being told the batch vector, which the reference's layer API does not carry.  The kernels that
// keep a block's rows in shared memory (agg_block.cuh) partition their work by these ranges.
//   lo/hi per node -> prefix max of hi, suffix min of lo -> start flags -> block ids -> block_ptr
// ------------------------------------------------------------------------------------------------------------
__global__ void plan_block_range_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                        const int* __restrict__ rowptrT, const int* __restrict__ colT, int N, int K,
                                        int cap, int* __restrict__ hi, int* __restrict__ lo_rev) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  int lo = v, h = v;
  {
    const int b = rowptr[(size_t)v * K], e = min(rowptr[(size_t)(v + 1) * K], cap);
    for (int j = b; j < e; ++j) {
      const int c = __ldg(col + j);
      lo = min(lo, c);
      h = max(h, c);
    }
  }
  {
    const int b = rowptrT[(size_t)v * K], e = min(rowptrT[(size_t)(v + 1) * K], cap);
    for (int j = b; j < e; ++j) {
      const int c = __ldg(colT + j);
      lo = min(lo, c);
      h = max(h, c);
    }
  }
  hi[v] = h;
  lo_rev[N - 1 - v] = lo;
}

__global__ void plan_block_flag_kernel(const int* __restrict__ pmax, const int* __restrict__ smin_rev, int N,
                                       int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  flag[i] = (i == 0 || (pmax[i - 1] < i && smin_rev[N - 1 - i] >= i)) ? 1 : 0;
}

__global__ void plan_block_ptr_kernel(const int* __restrict__ flag, const int* __restrict__ bid, int N,
                                      int* __restrict__ block_ptr, int* __restrict__ bstats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (flag[i]) block_ptr[bid[i]] = i;
  if (i == N - 1) {
    const int nb = bid[i] + flag[i];
    block_ptr[nb] = N;
    bstats[0] = nb;
  }
}

__global__ void plan_block_stats_kernel(const int* __restrict__ block_ptr, const int* __restrict__ rowptr, int K,
                                        int N, int cap, int* __restrict__ bstats) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bstats[0]) return;
  const int v0 = block_ptr[b], v1 = block_ptr[b + 1];
  atomicMax(&bstats[1], v1 - v0);
  atomicMax(&bstats[2], min(rowptr[(size_t)v1 * K], cap) - min(rowptr[(size_t)v0 * K], cap));
}

static size_t scan_temp_bytes(int rows_plus_1) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (int*)nullptr, (int*)nullptr, rows_plus_1);
  return bytes;
}

}  // namespace kp

extern "C" {

int kp_plan_workspace_bytes(int32_t N, int32_t E, int32_t K, size_t* bytes) {
  KP_CHECK_ARG(N >= 0 && E >= 0 && K >= 1 && bytes, "kp_plan_workspace_bytes: bad arguments");
  long long rows = (long long)N * K;
  KP_CHECK_ARG(rows + 1 < (1ll << 31) && (long long)E * K + rows < (1ll << 31),
               "kp_plan: N*K or E*K exceeds int32 indexing");
  size_t scan = kp::align_up(kp::scan_temp_bytes((int)rows + 1), 256);
  // fill phase: two cursor arrays (rows) + two edge-id arrays (upper bound E*K + rows entries each)
  size_t fill = kp::align_up(sizeof(int) * (size_t)rows, 256) * 2 +
                kp::align_up(sizeof(int) * ((size_t)E * K + rows), 256) * 2;
  *bytes = scan > fill ? scan : fill;
  return 0;
}

int kp_plan_count(const kp_plan_input* in, int32_t* rowptr, int32_t* rowptrT, int32_t* indeg, int32_t* stats,
                  void* workspace, size_t workspace_bytes, void* stream) {
  KP_CHECK_ARG(in && rowptr && rowptrT && indeg && stats, "kp_plan_count: null argument");
  const int N = in->N, E = in->E, K = in->K;
  size_t need = 0;
  if (kp_plan_workspace_bytes(N, E, K, &need)) return 1;
  KP_CHECK_ARG(workspace_bytes >= need && (workspace || need == 0), "kp_plan_count: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  long long rows = (long long)N * K;
  static const bool small_off = getenv("KP_PLAN_SMALL") && atoi(getenv("KP_PLAN_SMALL")) == 0;     // A/B switch
  const bool small = !small_off && rows + 1 <= kp::PLAN_SCAN_MAX;   // one-CTA scans, fused zeroing (see plan_scan2_kernel)
  if (small) {
    KP_LAUNCH(kp::plan_zero_kernel, kp::ceil_div(rows + 1, 256 * 4), 256, 0, st, rowptr, rowptrT, rows + 1, indeg,
              (long long)(N > 0 ? N : 1), stats);
  } else {
    KP_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int) * (rows + 1), st));
    KP_CUDA(cudaMemsetAsync(rowptrT, 0, sizeof(int) * (rows + 1), st));
    KP_CUDA(cudaMemsetAsync(indeg, 0, sizeof(int) * (size_t)(N > 0 ? N : 1), st));
    KP_CUDA(cudaMemsetAsync(stats, 0, sizeof(int) * 4, st));
  }
  long long total = (long long)E * K;
  if (total > 0) {
    KP_LAUNCH(kp::plan_count_kernel, kp::ceil_div(total, 256), 256, 0, st, in->src, in->dst, in->attr,
              in->attr_stride, N, E, K, rowptr, rowptrT, indeg, stats);
  }
  if (in->self_loops && rows > 0) {
    KP_LAUNCH(kp::plan_add_loops_kernel, kp::ceil_div(rows, 256), 256, 0, st, rowptr, rowptrT, (int)rows);
  }
  if (small) {
    KP_LAUNCH(kp::plan_scan2_kernel, 2, kp::PLAN_SCAN_THREADS, 0, st, rowptr, rowptrT, (int)rows + 1, stats);
    return 0;
  }
  size_t temp = workspace_bytes;
  KP_CUDA(cub::DeviceScan::ExclusiveSum(workspace, temp, rowptr, rowptr, (int)rows + 1, st));
  temp = workspace_bytes;
  KP_CUDA(cub::DeviceScan::ExclusiveSum(workspace, temp, rowptrT, rowptrT, (int)rows + 1, st));
  kp::g_launches.fetch_add(2, std::memory_order_relaxed);
  KP_LAUNCH(kp::plan_nnz_kernel, 1, 1, 0, st, rowptr, (int)rows, stats);
  return 0;
}

int kp_plan_fill(const kp_plan_input* in, const int32_t* rowptr, const int32_t* rowptrT, int32_t* col,
                 uint16_t* attr16, int32_t* colT, float* dinv, int32_t capacity, void* workspace,
                 size_t workspace_bytes, void* stream) {
  KP_CHECK_ARG(in && rowptr && rowptrT, "kp_plan_fill: null argument");
  const int N = in->N, E = in->E, K = in->K;
  size_t need = 0;
  if (kp_plan_workspace_bytes(N, E, K, &need)) return 1;
  KP_CHECK_ARG(workspace_bytes >= need && (workspace || need == 0), "kp_plan_fill: workspace too small");
  KP_CHECK_ARG(!in->self_loops || dinv, "kp_plan_fill: self_loops requires dinv");
  cudaStream_t st = (cudaStream_t)stream;
  long long rows = (long long)N * K;
  if (rows == 0) return 0;
  char* w = (char*)workspace;
  size_t cur_b = kp::align_up(sizeof(int) * (size_t)rows, 256);
  size_t eid_b = kp::align_up(sizeof(int) * ((size_t)E * K + rows), 256);
  int* cur = (int*)w;
  int* curT = (int*)(w + cur_b);
  int* eid = (int*)(w + 2 * cur_b);
  int* eidT = (int*)(w + 2 * cur_b + eid_b);
  KP_CUDA(cudaMemsetAsync(cur, 0, 2 * cur_b, st));
  long long total = (long long)E * K;
  if (total > 0) {
    KP_CHECK_ARG(col && attr16 && colT, "kp_plan_fill: null output");
    KP_LAUNCH(kp::plan_scatter_kernel, kp::ceil_div(total, 256), 256, 0, st, in->src, in->dst, in->attr,
              in->attr_stride, N, E, K, rowptr, rowptrT, cur, curT, eid, eidT, (int)capacity);
  }
  KP_LAUNCH(kp::plan_emit_kernel, kp::ceil_div(2 * rows, 128), 128, 0, st, in->src, in->dst, in->attr,
            in->attr_stride, N, K, in->self_loops, rowptr, rowptrT, eid, eidT, col, attr16, colT, dinv, (int)capacity);
  return 0;
}

int kp_plan_clamp(int32_t* rowptr, int32_t* rowptrT, int32_t N, int32_t K, int32_t capacity, void* stream) {
  KP_CHECK_ARG(rowptr && rowptrT && N >= 0 && K >= 1 && capacity >= 0, "kp_plan_clamp: bad argument");
  const long long n = (long long)N * K + 1;
  KP_LAUNCH(kp::plan_clamp_kernel, kp::ceil_div(n, 256), 256, 0, stream, rowptr, rowptrT, (int)n, (int)capacity);
  return 0;
}

int kp_plan_blocks_workspace_bytes(int32_t N, size_t* bytes) {
  KP_CHECK_ARG(N >= 0 && bytes, "kp_plan_blocks_workspace_bytes: bad arguments");
  size_t t1 = 0, t2 = 0;
  cub::DeviceScan::InclusiveScan(nullptr, t1, (int*)nullptr, (int*)nullptr, cub::Max(), N > 0 ? N : 1);
  cub::DeviceScan::ExclusiveSum(nullptr, t2, (int*)nullptr, (int*)nullptr, N > 0 ? N : 1);
  const size_t arr = kp::align_up(sizeof(int) * (size_t)(N > 0 ? N : 1), 256);
  *bytes = 5 * arr + kp::align_up(t1 > t2 ? t1 : t2, 256);
  return 0;
}

int kp_plan_blocks(const int32_t* rowptr, const int32_t* col, const int32_t* rowptrT, const int32_t* colT, int32_t N,
                   int32_t K, int32_t capacity, int32_t* block_ptr, int32_t* block_stats, void* workspace,
                   size_t workspace_bytes, void* stream) {
  KP_CHECK_ARG(rowptr && rowptrT && block_ptr && block_stats && N >= 0 && K >= 1, "kp_plan_blocks: bad argument");
  size_t need = 0;
  if (kp_plan_blocks_workspace_bytes(N, &need)) return 1;
  KP_CHECK_ARG(workspace_bytes >= need && workspace, "kp_plan_blocks: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  KP_CUDA(cudaMemsetAsync(block_stats, 0, sizeof(int) * 4, st));
  if (N == 0) {
    KP_CUDA(cudaMemsetAsync(block_ptr, 0, sizeof(int), st));
    return 0;
  }
  KP_CHECK_ARG(col && colT, "kp_plan_blocks: null plan arrays");
  const size_t arr = kp::align_up(sizeof(int) * (size_t)N, 256);
  char* w = (char*)workspace;
  int* hi = (int*)w;
  int* lo_rev = (int*)(w + arr);
  int* pmax = (int*)(w + 2 * arr);
  int* smin_rev = (int*)(w + 3 * arr);
  int* flag = (int*)(w + 4 * arr);
  void* temp = w + 5 * arr;
  size_t temp_bytes = workspace_bytes - 5 * arr;
  const int grid = kp::ceil_div(N, 256);
  KP_LAUNCH(kp::plan_block_range_kernel, grid, 256, 0, st, rowptr, col, rowptrT, colT, N, K, (int)capacity, hi, lo_rev);
  size_t tb = temp_bytes;
  KP_CUDA(cub::DeviceScan::InclusiveScan(temp, tb, hi, pmax, cub::Max(), N, st));
  tb = temp_bytes;
  KP_CUDA(cub::DeviceScan::InclusiveScan(temp, tb, lo_rev, smin_rev, cub::Min(), N, st));
  KP_LAUNCH(kp::plan_block_flag_kernel, grid, 256, 0, st, pmax, smin_rev, N, flag);
  tb = temp_bytes;
  int* bid = hi;                               // hi is dead after the scan
  KP_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, flag, bid, N, st));
  kp::g_launches.fetch_add(3, std::memory_order_relaxed);
  KP_LAUNCH(kp::plan_block_ptr_kernel, grid, 256, 0, st, flag, bid, N, block_ptr, block_stats);
  KP_LAUNCH(kp::plan_block_stats_kernel, grid, 256, 0, st, block_ptr, rowptr, K, N, (int)capacity, block_stats);
  return 0;
}

}  // extern "C"
