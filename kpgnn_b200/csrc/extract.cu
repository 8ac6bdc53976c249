// extract.cu -- batched K-hop neighbourhood + peripheral-subgraph extraction (sm_100a).
//
// Replaces the reference's dense-matrix / networkx pipeline, data_utils.py:20-241, for a whole BATCH of graphs:
//   hops        W[g][k][s][v] = saturated walk count of length k+1 from s to v             (data_utils.py:110-125)
//               spd: kept only where dist(s,v) == k+1 (= number of shortest paths)         (:63-74)
//               gd : kept for every k (a pair can sit in several hops)                      (:57-62)
//               one warp per source node runs a frontier-BFS (spd) / sparse power iteration (gd) over the CSR
//   emit        hop-labelled edge_index / edge_attr in the reference's int64 layout and order (:76-92)
//   peripheral  per (node, hop): edge-type histogram top-k and distance configuration of the subgraph induced
//               by the hop's node set                                                       (:128-241)
// Integer work throughout; integer atomics only, so every output is bit-reproducible and equals the reference
// wherever the reference is defined (walk counts < 2^31, see oracle/extract_np.py).
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace kp {

struct GraphRef {
  int base, n, local;
  const uint16_t* W;   // this graph's [K][n][n] block
};

__device__ __forceinline__ GraphRef graph_of(const kp_extract_input& in, const uint16_t* W, int node) {
  GraphRef r;
  const int g = __ldg(in.node_graph + node);
  r.base = __ldg(in.gptr + g);
  r.n = __ldg(in.gptr + g + 1) - r.base;
  r.local = node - r.base;
  r.W = W + (size_t)in.K * (size_t)__ldg(in.pair_off + g);
  return r;
}

__device__ __forceinline__ int warp_sum(int v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum64(long long v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------------------
// hops: one warp per source node
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
extract_hops_kernel(const kp_extract_input in, uint16_t* __restrict__ W, int* __restrict__ degK,
                    unsigned int* __restrict__ tmp_all, unsigned char* __restrict__ seen_all) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned int* tmp = tmp_all + (size_t)warp * in.n_max;
  unsigned char* seen = seen_all + (size_t)warp * in.n_max;
  const unsigned int cap = (unsigned int)in.cap;
  const bool spd = (in.kernel == 0);
  for (int s = warp; s < in.N; s += nwarps) {
    const GraphRef gr = graph_of(in, W, s);
    const int n = gr.n, sl = gr.local;
    uint16_t* Wg = const_cast<uint16_t*>(gr.W);
    for (int v = lane; v < n; v += 32) {
      tmp[v] = 0;
      seen[v] = 0;
    }
    __syncwarp();
    for (int e = __ldg(in.erow + s) + lane; e < __ldg(in.erow + s + 1); e += 32)
      tmp[__ldg(in.ecol + e) - gr.base] = (unsigned int)__ldg(in.emult + e);
    __syncwarp();
    for (int k = 0; k < in.K; ++k) {
      uint16_t* row = Wg + ((size_t)k * n + sl) * n;
      if (k > 0) {
        const uint16_t* prev = Wg + ((size_t)(k - 1) * n + sl) * n;
        for (int u = lane; u < n; u += 32) {
          const unsigned int c = prev[u];
          if (c) {
            const int gu = gr.base + u;
            for (int e = __ldg(in.erow + gu); e < __ldg(in.erow + gu + 1); ++e)
              atomicAdd(&tmp[__ldg(in.ecol + e) - gr.base], c * (unsigned int)__ldg(in.emult + e));
          }
        }
        __syncwarp();
      }
      for (int v = lane; v < n; v += 32) {
        const unsigned int t = tmp[v];
        unsigned int val = t < cap ? t : cap;
        if (spd) {
          if (v == sl || seen[v]) val = 0;
          if (val) seen[v] = 1;
        }
        row[v] = (uint16_t)val;
        tmp[v] = 0;
      }
      __syncwarp();
    }
    int cnt = 0;
    for (int v = lane; v < n; v += 32) {
      if (v == sl) continue;
      unsigned int any = 0;
      for (int k = 0; k < in.K; ++k) any |= Wg[((size_t)k * n + sl) * n + v];
      cnt += any != 0;
    }
    cnt = warp_sum(cnt);
    if (lane == 0) degK[s] = cnt;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------------
// emit: reference wire layout, row-major (src asc, dst asc) inside each graph, graphs concatenated
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
extract_emit_kernel(const kp_extract_input in, const uint16_t* __restrict__ W, const int* __restrict__ eptr,
                    int64_t* __restrict__ edge_index, int64_t* __restrict__ edge_attr, long long EK) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int C = in.max_edge_attr_num;
  for (int s = warp; s < in.N; s += nwarps) {
    const GraphRef gr = graph_of(in, W, s);
    const int n = gr.n, sl = gr.local;
    long long pos = __ldg(eptr + s);
    const int eb = __ldg(in.erow + s), ee = __ldg(in.erow + s + 1);
    for (int v0 = 0; v0 < n; v0 += 32) {
      const int v = v0 + lane;
      unsigned int any = 0;
      if (v < n && v != sl)
        for (int k = 0; k < in.K; ++k) any |= gr.W[((size_t)k * n + sl) * n + v];
      const unsigned int mask = __ballot_sync(0xffffffffu, any != 0);
      if (any) {
        const long long p = pos + __popc(mask & ((1u << lane) - 1));
        edge_index[p] = s;
        edge_index[EK + p] = gr.base + v;
        long long t = 0;
        for (int e = eb; e < ee; ++e)
          if (__ldg(in.ecol + e) == gr.base + v) t = __ldg(in.etype + e);
        int64_t* arow = edge_attr + p * in.K;
        arow[0] = t;                                                   // data_utils.py:80-81
        for (int k = 1; k < in.K; ++k) {                               // :84-90
          int a = gr.W[((size_t)k * n + sl) * n + v];
          a = a < C ? a : C;
          arow[k] = a > 0 ? a + 1 : 0;
        }
      }
      pos += __popc(mask);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// peripheral: one warp per (node, hop)
// ------------------------------------------------------------------------------------------------------------
#define KP_MAX_HOPNUM 32
__global__ void __launch_bounds__(256)
extract_peripheral_kernel(const kp_extract_input in, const uint16_t* __restrict__ W, int64_t* __restrict__ pea,
                          int64_t* __restrict__ pca, uint16_t* __restrict__ mem_all,
                          unsigned char* __restrict__ dj_all, int* __restrict__ hist_all, int bins) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  uint16_t* mem = mem_all + (size_t)warp * in.n_max;
  unsigned char* dj = dj_all + (size_t)warp * in.n_max;
  int* hist = hist_all + (size_t)warp * bins;
  const int H = in.max_hop_num, MET = in.max_edge_type;
  const long long ntasks = (long long)in.N * in.K;
  for (long long task = warp; task < ntasks; task += nwarps) {
    const int i = (int)(task / in.K), k = (int)(task - (long long)i * in.K);
    const GraphRef gr = graph_of(in, W, i);
    const int n = gr.n, il = gr.local, base = gr.base;
    const uint16_t* rowk = gr.W + ((size_t)k * n + il) * n;
    // member list S (ascending), data_utils.py:185
    int m = 0;
    for (int v0 = 0; v0 < n; v0 += 32) {
      const int v = v0 + lane;
      const bool is = (v < n) && (v != il) && rowk[v] != 0;
      const unsigned int mask = __ballot_sync(0xffffffffu, is);
      if (is) mem[m + __popc(mask & ((1u << lane) - 1))] = (uint16_t)v;
      m += __popc(mask);
      if (v < n) dj[v] = 255;
    }
    __syncwarp();
    if (m < 2) continue;                                               // :188-189
    // directed edges of the induced subgraph, histogram of their type values  (:190-198)
    for (int b = lane; b < bins; b += 32) hist[b] = 0;
    __syncwarp();
    int nedges = 0;
    for (int idx = lane; idx < m; idx += 32) {
      const int ga = base + mem[idx];
      for (int e = __ldg(in.erow + ga); e < __ldg(in.erow + ga + 1); ++e) {
        const int b = __ldg(in.ecol + e) - base;
        const int t = __ldg(in.etype + e);
        if (t != 0 && b != il && rowk[b] != 0) {
          ++nedges;
          atomicAdd(&hist[t], 1);
        }
      }
    }
    nedges = warp_sum(nedges);
    __syncwarp();
    if (nedges == 0) continue;                                         // :193-194
    // stable descending top-MET over bins 2.. (:198-204): repeated arg-max, lowest index wins ties
    int64_t* pe_out = pea + task * MET * 2;
    for (int r = 0; r < MET; ++r) {
      int best = -1, bi = 0x7fffffff;
      for (int b = 2 + lane; b < bins; b += 32) {
        const int c = hist[b];
        if (c > best) {
          best = c;
          bi = b;
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      if (lane == 0) {
        pe_out[r * 2] = bi - 2;
        pe_out[r * 2 + 1] = best < in.max_edge_count ? best : in.max_edge_count;
        hist[bi] = -1;
      }
      __syncwarp();
    }
    // configuration: BFS from every member inside the induced subgraph, cutoff H  (:205-220)
    long long cf0 = 0;
    int cf[KP_MAX_HOPNUM + 1];
#pragma unroll 1
    for (int h = 0; h <= H; ++h) cf[h] = 0;
    if (H == 1) {
      // Cutoff 1 (run_simulation.py:103): the BFS from member j only reaches its neighbours inside the induced subgraph,
      // S'(j) = {b member, b != j, edge j->b of non-zero type}.  cf[1] += |S'(j)| and, when |S'(j)| >= 2, cf0 += the type
      // values of the edges a->b with a, b in S'(j) -- the same integers as the general loop below, but one LANE per
      // member instead of one warp-wide BFS per member (m sequential rounds of ~6 warp passes: 70 % of this kernel's time
      // on n = 1 280 regular graphs, where the far shells hold ~100 members).  Rows of the input CSR are sorted by
      // destination (pack_csr), so "b in S'(j)" is a binary search in row j.
      int c1 = 0;
      long long s0 = 0;
      for (int idx = lane; idx < m; idx += 32) {
        const int j = mem[idx], gj = base + j;
        const int jb = __ldg(in.erow + gj), je = __ldg(in.erow + gj + 1);
        int c = 0;
        for (int e = jb; e < je; ++e) {
          const int b = __ldg(in.ecol + e) - base;
          c += (__ldg(in.etype + e) != 0 && b != il && b != j && rowk[b] != 0);
        }
        c1 += c;
        if (c < 2) continue;
        for (int e = jb; e < je; ++e) {
          const int a = __ldg(in.ecol + e) - base;
          if (__ldg(in.etype + e) == 0 || a == il || a == j || rowk[a] == 0) continue;
          const int ga = base + a;
          for (int e2 = __ldg(in.erow + ga); e2 < __ldg(in.erow + ga + 1); ++e2) {
            const int b = __ldg(in.ecol + e2) - base;
            const int t = __ldg(in.etype + e2);
            if (t == 0 || b == il || b == j || rowk[b] == 0) continue;
            int lo = jb, hi = je;                                     // is b a type-carrying neighbour of j?
            const int key = base + b;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (__ldg(in.ecol + mid) < key) lo = mid + 1;
              else hi = mid;
            }
            if (lo < je && __ldg(in.ecol + lo) == key && __ldg(in.etype + lo) != 0) s0 += t;
          }
        }
      }
      cf[1] = warp_sum(c1);
      cf0 = warp_sum64(s0);
    } else
    for (int jx = 0; jx < m; ++jx) {
      for (int idx = lane; idx < m; idx += 32) dj[mem[idx]] = 255;
      __syncwarp();
      if (lane == 0) dj[mem[jx]] = 0;
      __syncwarp();
      for (int h = 1; h <= H; ++h) {
        int found = 0;
        for (int idx = lane; idx < m; idx += 32) {
          const int a = mem[idx];
          if (dj[a] != h - 1) continue;
          const int ga = base + a;
          for (int e = __ldg(in.erow + ga); e < __ldg(in.erow + ga + 1); ++e) {
            const int b = __ldg(in.ecol + e) - base;
            if (__ldg(in.etype + e) != 0 && b != il && rowk[b] != 0 && dj[b] == 255) {
              dj[b] = (unsigned char)h;
              found = 1;
            }
          }
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, found)) break;
        int c = 0;
        for (int idx = lane; idx < m; idx += 32) c += (dj[mem[idx]] == h);
        c = warp_sum(c);
        cf[h] += c;
        if (c >= 2) {                                                  // :209-214, sums edge-TYPE values
          long long sum = 0;
          for (int idx = lane; idx < m; idx += 32) {
            const int a = mem[idx];
            if (dj[a] != h) continue;
            const int ga = base + a;
            for (int e = __ldg(in.erow + ga); e < __ldg(in.erow + ga + 1); ++e) {
              const int b = __ldg(in.ecol + e) - base;
              const int t = __ldg(in.etype + e);
              if (t != 0 && b != il && rowk[b] != 0 && dj[b] == h) sum += t;
            }
          }
          cf0 += warp_sum64(sum);
        }
      }
      __syncwarp();
    }
    if (lane == 0) {
      int64_t* pc_out = pca + task * (H + 1);
      const long long mdc = in.max_distance_count;
      pc_out[0] = cf0 < mdc ? cf0 : mdc;                               // :218-219
      for (int h = 1; h <= H; ++h) pc_out[h] = cf[h] < mdc ? cf[h] : mdc;
    }
    __syncwarp();
  }
}

static int check_input(const kp_extract_input* in) {
  KP_CHECK_ARG(in, "kp_extract: null input");
  KP_CHECK_ARG(in->G >= 0 && in->N >= 0 && in->K >= 1, "kp_extract: bad sizes");
  KP_CHECK_ARG(in->n_max >= 0 && in->n_max <= 65535, "kp_extract: graphs are limited to 65535 nodes (got %d)",
               in->n_max);
  KP_CHECK_ARG(in->cap >= 1 && in->cap <= 65535, "kp_extract: cap must be in [1,65535]");
  KP_CHECK_ARG(in->max_hop_num <= KP_MAX_HOPNUM, "kp_extract: max_hop_num > %d not supported", KP_MAX_HOPNUM);
  KP_CHECK_ARG(in->kernel == 0 || in->kernel == 1, "kp_extract: kernel must be 0 (spd) or 1 (gd)");
  return 0;
}

static int extract_warps(long long work) {
  long long w = work < 1 ? 1 : work;
  const long long maxw = (long long)kNumSMs * 8 * 8;   // 8 CTAs of 8 warps per SM
  w = w < maxw ? w : maxw;
  return (int)((w + 7) / 8 * 8);                       // whole CTAs: the scratch is indexed by launched warp
}

}  // namespace kp

extern "C" {

int kp_extract_workspace_bytes(const kp_extract_input* in, int64_t total_pairs, size_t* hop_bytes,
                               size_t* scratch_bytes) {
  if (kp::check_input(in)) return 1;
  KP_CHECK_ARG(hop_bytes && scratch_bytes && total_pairs >= 0, "kp_extract_workspace_bytes: bad arguments");
  *hop_bytes = sizeof(uint16_t) * (size_t)total_pairs * in->K;
  const int bins = (in->max_edge_type + 2 > in->max_type_value + 1) ? in->max_edge_type + 2 : in->max_type_value + 1;
  const size_t w1 = kp::extract_warps(in->N), w2 = kp::extract_warps((long long)in->N * in->K);
  const size_t hops = kp::align_up(w1 * in->n_max * sizeof(unsigned int), 256) + kp::align_up(w1 * in->n_max, 256);
  const size_t peri = kp::align_up(w2 * in->n_max * sizeof(uint16_t), 256) + kp::align_up(w2 * in->n_max, 256) +
                      kp::align_up(w2 * bins * sizeof(int), 256);
  size_t scan = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan, (int*)nullptr, (int*)nullptr, in->N + 1);
  size_t s = hops > peri ? hops : peri;
  *scratch_bytes = (s > scan ? s : kp::align_up(scan, 256)) + 256;
  return 0;
}

int kp_extract_hops(const kp_extract_input* in, uint16_t* W, int32_t* eptr, void* scratch, size_t scratch_bytes,
                    void* stream) {
  if (kp::check_input(in)) return 1;
  KP_CHECK_ARG(eptr && (W || in->N == 0), "kp_extract_hops: null output");
  cudaStream_t st = (cudaStream_t)stream;
  KP_CUDA(cudaMemsetAsync(eptr, 0, sizeof(int) * (size_t)(in->N + 1), st));
  if (in->N == 0) return 0;
  const int warps = kp::extract_warps(in->N);
  char* s = (char*)scratch;
  unsigned int* tmp = (unsigned int*)s;
  unsigned char* seen = (unsigned char*)(s + kp::align_up((size_t)warps * in->n_max * sizeof(unsigned int), 256));
  KP_CHECK_ARG(scratch && scratch_bytes >= kp::align_up((size_t)warps * in->n_max * 4, 256) +
                                               kp::align_up((size_t)warps * in->n_max, 256),
               "kp_extract_hops: scratch too small");
  KP_LAUNCH(kp::extract_hops_kernel, kp::ceil_div(warps, 8), 256, 0, st, *in, W, eptr, tmp, seen);
  size_t temp = scratch_bytes;
  KP_CUDA(cub::DeviceScan::ExclusiveSum(scratch, temp, eptr, eptr, in->N + 1, st));
  kp::g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int kp_extract_emit(const kp_extract_input* in, const uint16_t* W, const int32_t* eptr, int64_t* edge_index,
                    int64_t* edge_attr, int64_t EK, void* stream) {
  if (kp::check_input(in)) return 1;
  if (in->N == 0 || EK == 0) return 0;
  KP_CHECK_ARG(W && eptr && edge_index && edge_attr, "kp_extract_emit: null argument");
  const int warps = kp::extract_warps(in->N);
  KP_LAUNCH(kp::extract_emit_kernel, kp::ceil_div(warps, 8), 256, 0, (cudaStream_t)stream, *in, W, eptr, edge_index,
            edge_attr, (long long)EK);
  return 0;
}

int kp_extract_peripheral(const kp_extract_input* in, const uint16_t* W, int64_t* peripheral_edge_attr,
                          int64_t* peripheral_configuration_attr, void* scratch, size_t scratch_bytes,
                          void* stream) {
  if (kp::check_input(in)) return 1;
  KP_CHECK_ARG(in->max_hop_num >= 1 && in->max_edge_type >= 1,
               "kp_extract_peripheral: needs max_hop_num >= 1 and max_edge_type >= 1 (reference returns None)");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tasks = (size_t)in->N * in->K;
  if (tasks == 0) return 0;
  KP_CHECK_ARG(W && peripheral_edge_attr && peripheral_configuration_attr, "kp_extract_peripheral: null argument");
  KP_CUDA(cudaMemsetAsync(peripheral_edge_attr, 0, sizeof(int64_t) * tasks * in->max_edge_type * 2, st));
  KP_CUDA(cudaMemsetAsync(peripheral_configuration_attr, 0, sizeof(int64_t) * tasks * (in->max_hop_num + 1), st));
  const int bins = (in->max_edge_type + 2 > in->max_type_value + 1) ? in->max_edge_type + 2 : in->max_type_value + 1;
  const int warps = kp::extract_warps((long long)tasks);
  char* s = (char*)scratch;
  const size_t b0 = kp::align_up((size_t)warps * in->n_max * sizeof(uint16_t), 256);
  const size_t b1 = kp::align_up((size_t)warps * in->n_max, 256);
  const size_t b2 = kp::align_up((size_t)warps * bins * sizeof(int), 256);
  KP_CHECK_ARG(scratch && scratch_bytes >= b0 + b1 + b2, "kp_extract_peripheral: scratch too small");
  KP_LAUNCH(kp::extract_peripheral_kernel, kp::ceil_div(warps, 8), 256, 0, st, *in, W, peripheral_edge_attr,
            peripheral_configuration_attr, (uint16_t*)s, (unsigned char*)(s + b0), (int*)(s + b0 + b1), bins);
  return 0;
}

}  // extern "C"
