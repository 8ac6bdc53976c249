// peer.cu -- data-parallel gradient exchange over NVLink peer memory (the reference is single-GPU, so the
// contract is train_ZINC.py:29-47's optimisation step with the batch split over ranks: every rank applies the MEAN of
// the ranks' gradients).  One kernel per step replaces {NCCL all-reduce launch, divide kernel, second graph launch}:
//
//   every rank owns one peer-visible block  [ flags | gradient (n floats) | result (n floats) ]  (cudaMalloc + CUDA IPC, opened by every
//   other rank once at set-up); the kernel runs PEER_CTAS CTAs, CTA b owning slice b of the vector:
//     1. publish "my gradient of epoch e is complete" into ready[b][rank] of EVERY rank's flag block (st.release.sys),
//        wait until all ranks published theirs into mine                                 -- one NVLink round trip
//     2. two-shot: rank r reduces only ITS 1/world of the vector -- (1/world) * (g_0[i] + g_1[i] + ...) read straight
//        from the peers' blocks (128-bit volatile loads, eight in flight per thread), summed in rank order -- and
//        stores the result into the result vector of EVERY rank: (world-1)/world of the vector read and written per
//        rank instead of (world-1) vectors read; one rank computes each element => replicas are bit-identical
//     3. publish "my part is delivered, I am done reading" to every rank, wait for all of theirs: the result is
//        complete here and the next step may overwrite the gradient blocks.
//   CTA-local flag protocol: no rank-wide or grid-wide barrier, no co-residency requirement beyond grid <= #SMs.
//   Every wait is bounded (PEER_TIMEOUT_NS on %globaltimer): a missing peer sets *error and lets the kernel end, it can
//   never wedge the GPU.
#include "common.cuh"

namespace kp {

constexpr int PEER_CTAS = KP_PEER_CTAS;
constexpr int PEER_THREADS = 512;
constexpr unsigned long long PEER_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_volatile1(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// flag block of one rank: ready[PEER_CTAS][KP_PEER_MAX] then done[PEER_CTAS][KP_PEER_MAX]
__device__ __forceinline__ int* flag_slot(char* block, int phase, int cta, int rank) {
  return reinterpret_cast<int*>(block) + ((phase * PEER_CTAS + cta) * KP_PEER_MAX + rank);
}

__device__ __forceinline__ void exchange(const kp_peer_desc& d, int phase, int cta, int e) {
  const int t = threadIdx.x;
  if (t < d.world) {
    st_release_sys(flag_slot(d.block[t], phase, cta, d.rank), e);        // into rank t's block
    const int* mine = flag_slot(d.block[d.rank], phase, cta, t);         // written by rank t
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(mine) < e) {
      if (global_ns() - t0 > PEER_TIMEOUT_NS) {
        atomicExch(d.error, 1 + phase);
        break;
      }
    }
  }
  __syncthreads();
}

// sum of slice [lo4, hi4) (float4 units) over the W ranks' gradients in rank order, times scale, stored into the result
// vector of EVERY rank (two-shot exchange: each rank reduces 1/W of the vector and delivers it)
template <int W>
__device__ __forceinline__ void reduce_slice(const kp_peer_desc& d, long long lo4, long long hi4) {
  const float* g[W];
  float* o[W];
  const size_t recv_off = (size_t)KP_PEER_FLAG_BYTES + KP_PEER_VECTOR_BYTES(d.n);
#pragma unroll
  for (int r = 0; r < W; ++r) {
    g[r] = reinterpret_cast<const float*>(d.block[r] + KP_PEER_FLAG_BYTES);
    o[r] = reinterpret_cast<float*>(d.block[r] + recv_off);
  }
  constexpr int U = W <= 2 ? 4 : (W <= 4 ? 2 : 1);          // W*U 128-bit loads in flight per thread
  long long i = lo4 + threadIdx.x;
  for (; i + (long long)(U - 1) * PEER_THREADS < hi4; i += (long long)U * PEER_THREADS) {
    float4 v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int r = 0; r < W; ++r) v[u][r] = ld_volatile4(g[r] + 4 * (i + (long long)u * PEER_THREADS));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4 a = v[u][0];
#pragma unroll
      for (int r = 1; r < W; ++r) { a.x += v[u][r].x; a.y += v[u][r].y; a.z += v[u][r].z; a.w += v[u][r].w; }
      a.x *= d.scale; a.y *= d.scale; a.z *= d.scale; a.w *= d.scale;
#pragma unroll
      for (int r = 0; r < W; ++r) *reinterpret_cast<float4*>(o[r] + 4 * (i + (long long)u * PEER_THREADS)) = a;
    }
  }
  for (; i < hi4; i += PEER_THREADS) {
    float4 a = ld_volatile4(g[0] + 4 * i);
#pragma unroll
    for (int r = 1; r < W; ++r) {
      const float4 b = ld_volatile4(g[r] + 4 * i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    a.x *= d.scale; a.y *= d.scale; a.z *= d.scale; a.w *= d.scale;
#pragma unroll
    for (int r = 0; r < W; ++r) *reinterpret_cast<float4*>(o[r] + 4 * i) = a;
  }
}

__global__ void __launch_bounds__(PEER_THREADS) peer_allreduce_mean_kernel(const kp_peer_desc d) {
  __shared__ int s_epoch;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) s_epoch = d.epoch[b] + 1;
  __syncthreads();
  const int e = s_epoch;
  exchange(d, 0, b, e);                                       // every rank's gradient is complete
  // this rank owns float4s [rank*per_rank, (rank+1)*per_rank); CTA b owns 1/PEER_CTAS of that
  const long long n4 = d.n >> 2;
  const long long per_rank = (n4 + d.world - 1) / d.world;
  const long long rlo = min(n4, (long long)d.rank * per_rank), rhi = min(n4, rlo + per_rank);
  const long long per = (rhi - rlo + PEER_CTAS - 1) / PEER_CTAS;
  const long long lo4 = min(rhi, rlo + (long long)b * per), hi4 = min(rhi, lo4 + per);
  switch (d.world) {
    case 1: reduce_slice<1>(d, lo4, hi4); break;
    case 2: reduce_slice<2>(d, lo4, hi4); break;
    case 3: reduce_slice<3>(d, lo4, hi4); break;
    case 4: reduce_slice<4>(d, lo4, hi4); break;
    case 5: reduce_slice<5>(d, lo4, hi4); break;
    case 6: reduce_slice<6>(d, lo4, hi4); break;
    case 7: reduce_slice<7>(d, lo4, hi4); break;
    default: reduce_slice<8>(d, lo4, hi4); break;
  }
  __threadfence_system();                                     // this thread's result stores are visible system-wide ...
  __syncthreads();                                            // ... before the CTA's release below
  exchange(d, 1, b, e);                                       // sub-slice b of every rank's slice has been delivered here
  if (threadIdx.x == 0) d.epoch[b] = e;
}

}  // namespace kp

extern "C" size_t kp_peer_block_bytes(int64_t n) {
  return (size_t)KP_PEER_FLAG_BYTES + 2 * KP_PEER_VECTOR_BYTES(n < 0 ? 0 : n);
}

extern "C" int kp_peer_alloc(size_t bytes, void** ptr) {
  KP_CHECK_ARG(ptr && bytes > 0, "kp_peer_alloc: bad argument");
  KP_CUDA(cudaMalloc(ptr, bytes));
  KP_CUDA(cudaMemset(*ptr, 0, bytes));
  KP_CUDA(cudaDeviceSynchronize());
  return 0;
}

extern "C" int kp_peer_free(void* ptr) {
  if (ptr) KP_CUDA(cudaFree(ptr));
  return 0;
}

extern "C" int kp_peer_export(const void* ptr, unsigned char handle[KP_PEER_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == KP_PEER_HANDLE_BYTES, "IPC handle size");
  KP_CHECK_ARG(ptr && handle, "kp_peer_export: bad argument");
  cudaIpcMemHandle_t h;
  KP_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle, &h, sizeof(h));
  return 0;
}

extern "C" int kp_peer_import(const unsigned char handle[KP_PEER_HANDLE_BYTES], void** ptr) {
  KP_CHECK_ARG(ptr && handle, "kp_peer_import: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  KP_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int kp_peer_release(void* ptr) {
  if (ptr) KP_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int kp_peer_allreduce_mean(const kp_peer_desc* d, void* stream) {
  KP_CHECK_ARG(d && d->world >= 1 && d->world <= KP_PEER_MAX && d->rank >= 0 && d->rank < d->world && d->n >= 0 &&
                   d->n % 4 == 0 && d->epoch && d->error,
               "kp_peer_allreduce_mean: bad descriptor (n must be a multiple of 4)");
  for (int r = 0; r < d->world; ++r) KP_CHECK_ARG(d->block[r], "kp_peer_allreduce_mean: missing peer block");
  KP_LAUNCH(kp::peer_allreduce_mean_kernel, kp::PEER_CTAS, kp::PEER_THREADS, 0, stream, *d);
  return 0;
}
