// dense.cu -- the dense block of a KP-GIN+ layer as ONE persistent kernel per direction (sm_100a):
//     y1 = X W1^T + b1;  z1 = relu(BN1(y1));  y2 = z1 W2^T + b2;  z2 = relu(BN2(y2));  out = BN3(z2) + R
// (layers/KPGINplus.py:25-30,78 and models/GNNs.py:430-438; BN3 and R optional).
//
// Why: at molecule-batch sizes (N ~ 3 000 rows, 104 channels) the step was a chain of ~45 kernels of 2-11 us each
// per layer around the aggregation (2 + 4 SIMT GEMMs, 3 + 3 BatchNorm kernels, 4 column sums, bias / residual
// adds; profiles/r1v_step_kineto.txt) -- all latency, no bandwidth.  Here each CTA keeps a slab of ~32 rows and
// BOTH weight matrices in shared memory for the whole block; the only cross-CTA dependencies are the three batch
// statistics (and, backward, the weight-gradient sums), exchanged through small per-CTA partials in L2 behind a
// grid-wide barrier.  Statistics use Chan's parallel variance (per-slab mean / M2 merged in a fixed order), the
// weight gradients are per-CTA [Cout x Cin] partials summed in a fixed order: bit-reproducible, no float atomics.
// All arithmetic is fp32 FMA (the 1e-5 parity bar rules out TF32); the GEMMs are ~64 MFLOP -- latency, not
// throughput, is what this kernel removes.
#include <stdlib.h>

#include "common.cuh"

namespace kp {

constexpr int DB_THREADS = 512;   // 16 warps: the block is a chain of short phases, warps are the only latency hiding
constexpr int DB_WARPS = DB_THREADS / 32;
constexpr int DB_ROWS = 32;          // target rows per CTA (8 row quads x 26 channel quads = 208 of 512 threads own a GEMM tile at
                                     // 104 channels); measured inside the step: backward 302 us / 8 kernels vs 316 at 36 rows
                                     // and 292 at 28 (where the forward loses 3 us), profiles/r2_dense_mma.txt
constexpr int DB_MAX_GRID = kNumSMs; // every CTA must be resident: one per SM

__device__ __forceinline__ unsigned db_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// All CTAs are co-resident (grid <= #SMs, one CTA per SM by shared-memory footprint), so spinning is safe.
__device__ __forceinline__ void db_grid_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    while (db_ld_acquire(ctr) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

// -DKP_DENSE_TIMING: CTA 0 prints the nanosecond timestamps of its phase boundaries (tuning builds only)
#ifdef KP_DENSE_TIMING
#define DB_T(i)                                                              \
  do {                                                                       \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                               \
      unsigned long long _t;                                                 \
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(_t));                  \
      tstamp[i] = _t;                                                        \
    }                                                                        \
  } while (0)
#define DB_T_DECL __shared__ unsigned long long tstamp[24];
#define DB_T_PRINT(n, name)                                                  \
  do {                                                                       \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                               \
      printf(name ":");                                                      \
      for (int _i = 1; _i < n; ++_i) printf(" %d", (int)(tstamp[_i] - tstamp[_i - 1])); \
      printf("  total %d ns\n", (int)(tstamp[n - 1] - tstamp[0]));           \
    }                                                                        \
  } while (0)
#else
#define DB_T(i)
#define DB_T_DECL
#define DB_T_PRINT(n, name)
#endif

constexpr int DB_MAXI = (DB_MAX_GRID + DB_WARPS - 1) / DB_WARPS;   // partials per warp in a merge (held in registers)

__device__ __forceinline__ void db_cp16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void db_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void db_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Forward weight layout: row o of W [Co][Kd] keeps its 16-byte chunks, chunk kc stored at position kc ^ ((o>>2)&7)
// of a row padded to a multiple of 8 chunks.  A thread that owns output channels 4ct..4ct+3 then reads float4s
// whose bank group depends on ct only through the XOR: 8 consecutive threads hit 8 different bank groups
// (conflict-free) although the rows are 4*stride apart -- and the copy from global memory is a pure 16-byte
// permutation, so cp.async does it without a transposition pass through registers.
__host__ __device__ __forceinline__ int db_wstride(int Kd) { return (((Kd >> 2) + 7) & ~7) << 2; }
// allocation stride of a forward weight row: covers both the swizzled layout and the MMA path's [Kd+4] rows
__host__ __device__ __forceinline__ int db_wstride_any(int Kd) { return db_wstride(Kd) > Kd + 4 ? db_wstride(Kd) : Kd + 4; }

__device__ __forceinline__ void db_cp_weight_swizzled(const float* __restrict__ W, int Co, int Kd, float* __restrict__ Ws) {
  const int ck = Kd >> 2, ws = db_wstride(Kd);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < Co; o += DB_WARPS)
    for (int kc = lane; kc < ck; kc += 32)
      db_cp16(Ws + o * ws + ((kc ^ ((o >> 2) & 7)) << 2), W + (size_t)o * Kd + (kc << 2));
}
__device__ __forceinline__ void db_cp_rows(const float* __restrict__ g, int nfloats, float* __restrict__ sdst) {
  for (int i = threadIdx.x * 4; i < nfloats; i += DB_THREADS * 4) db_cp16(sdst + i, g + i);
}
// global rows r0..r0+nr of a [N][C] matrix -> shared slab (rows >= nr are zero-filled by plain stores)
__device__ __forceinline__ void db_cp_slab(const float* __restrict__ g, int r0, int nr, int nrp, int C,
                                           float* __restrict__ sdst) {
  const int n = nr * C;
  const float* src = g + (size_t)r0 * C;
  for (int i = threadIdx.x * 4; i < nrp * C; i += DB_THREADS * 4) {
    if (i < n) db_cp16(sdst + i, src + i);
    else *reinterpret_cast<float4*>(sdst + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// out[r][o] = bias[o] + sum_k A[r][k] * W[o][k]   (W in the swizzled layout above), r < nrp (multiple of 4), o < Co.
// Thread tile 4 rows x 4 channels: the phase is bound by shared-memory bandwidth (every row tile re-reads W), measured
// 5.1 us with 2x4 tiles on 16 warps vs ~3 us with 4x4 tiles on 8 of them.
// thread tile: 4 rows x 4*TCQ output channels.  TCQ = 2 (-DKP_DENSE_TILE8) halves the shared-memory traffic per FMA but
// leaves 117 of 512 threads busy: measured 6.0 us per GEMM phase vs 3.9 us for TCQ = 1, which stays the default.
template <int TCQ>
__device__ __forceinline__ void db_gemm_AWt_t(const float* __restrict__ A, int lda, int Kd, const float* __restrict__ Ws,
                                              int Co, const float* __restrict__ bias, int nrp, float* __restrict__ out,
                                              int ldo) {
  const int ctn = Co / (4 * TCQ), ntiles = ctn * (nrp >> 2), ws = db_wstride(Kd);
  for (int t = threadIdx.x; t < ntiles; t += DB_THREADS) {
    const int rt = t / ctn, ct = t - rt * ctn;
    const int r = rt * 4, o = ct * 4 * TCQ;
    float4 acc[TCQ][4];
#pragma unroll
    for (int q = 0; q < TCQ; ++q) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + o + 4 * q));
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][j] = bv;
    }
    const float* a = A + r * lda;
#pragma unroll 2
    for (int kc = 0; kc < (Kd >> 2); ++kc) {
      float4 x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = *reinterpret_cast<const float4*>(a + j * lda + (kc << 2));
#pragma unroll
      for (int q = 0; q < TCQ; ++q) {
        const int sw = (ct * TCQ + q) & 7;
        const float* w0 = Ws + (o + 4 * q) * ws + ((kc ^ sw) << 2);
        float4 w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(w0 + j * ws);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4& c = acc[q][j];
          c.x = fmaf(x[j].x, w[0].x, c.x); c.x = fmaf(x[j].y, w[0].y, c.x);
          c.x = fmaf(x[j].z, w[0].z, c.x); c.x = fmaf(x[j].w, w[0].w, c.x);
          c.y = fmaf(x[j].x, w[1].x, c.y); c.y = fmaf(x[j].y, w[1].y, c.y);
          c.y = fmaf(x[j].z, w[1].z, c.y); c.y = fmaf(x[j].w, w[1].w, c.y);
          c.z = fmaf(x[j].x, w[2].x, c.z); c.z = fmaf(x[j].y, w[2].y, c.z);
          c.z = fmaf(x[j].z, w[2].z, c.z); c.z = fmaf(x[j].w, w[2].w, c.z);
          c.w = fmaf(x[j].x, w[3].x, c.w); c.w = fmaf(x[j].y, w[3].y, c.w);
          c.w = fmaf(x[j].z, w[3].z, c.w); c.w = fmaf(x[j].w, w[3].w, c.w);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < TCQ; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(out + (r + j) * ldo + o + 4 * q) = acc[q][j];
  }
}
__device__ __forceinline__ void db_gemm_AWt(const float* __restrict__ A, int lda, int Kd, const float* __restrict__ Ws,
                                            int Co, const float* __restrict__ bias, int nrp, float* __restrict__ out,
                                            int ldo) {
#ifdef KP_DENSE_TILE8
  if (Co % 8 == 0) {
    db_gemm_AWt_t<2>(A, lda, Kd, Ws, Co, bias, nrp, out, ldo);
    return;
  }
#endif
  db_gemm_AWt_t<1>(A, lda, Kd, Ws, Co, bias, nrp, out, ldo);
}

// out[r][n] = sum_k A[r][k] * B[k][n]   for r < nrp (multiple of 4), n < Nn; all in shared memory; tile 4 x 4*TCQ
template <int TCQ>
__device__ __forceinline__ void db_gemm_AB_t(const float* __restrict__ A, int lda, int Kd, const float* __restrict__ B,
                                             int Nn, int nrp, float* __restrict__ out, int ldo) {
  const int ctn = Nn / (4 * TCQ), ntiles = ctn * (nrp >> 2);
  for (int t = threadIdx.x; t < ntiles; t += DB_THREADS) {
    const int rt = t / ctn, ct = t - rt * ctn;
    const int r = rt * 4, n = ct * 4 * TCQ;
    float4 acc[TCQ][4];
#pragma unroll
    for (int q = 0; q < TCQ; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* a = A + r * lda;
    const float* b = B + n;
#pragma unroll 2
    for (int k = 0; k < Kd; k += 4) {
      float4 x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = *reinterpret_cast<const float4*>(a + j * lda + k);
#pragma unroll
      for (int q = 0; q < TCQ; ++q) {
        float4 w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(b + (k + j) * Nn + 4 * q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4& c = acc[q][j];
          c.x = fmaf(x[j].x, w[0].x, c.x); c.y = fmaf(x[j].x, w[0].y, c.y);
          c.z = fmaf(x[j].x, w[0].z, c.z); c.w = fmaf(x[j].x, w[0].w, c.w);
          c.x = fmaf(x[j].y, w[1].x, c.x); c.y = fmaf(x[j].y, w[1].y, c.y);
          c.z = fmaf(x[j].y, w[1].z, c.z); c.w = fmaf(x[j].y, w[1].w, c.w);
          c.x = fmaf(x[j].z, w[2].x, c.x); c.y = fmaf(x[j].z, w[2].y, c.y);
          c.z = fmaf(x[j].z, w[2].z, c.z); c.w = fmaf(x[j].z, w[2].w, c.w);
          c.x = fmaf(x[j].w, w[3].x, c.x); c.y = fmaf(x[j].w, w[3].y, c.y);
          c.z = fmaf(x[j].w, w[3].z, c.z); c.w = fmaf(x[j].w, w[3].w, c.w);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < TCQ; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(out + (r + j) * ldo + n + 4 * q) = acc[q][j];
  }
}
__device__ __forceinline__ void db_gemm_AB(const float* __restrict__ A, int lda, int Kd, const float* __restrict__ B,
                                           int Nn, int nrp, float* __restrict__ out, int ldo) {
#ifdef KP_DENSE_TILE8
  if (Nn % 8 == 0) {
    db_gemm_AB_t<2>(A, lda, Kd, B, Nn, nrp, out, ldo);
    return;
  }
#endif
  db_gemm_AB_t<1>(A, lda, Kd, B, Nn, nrp, out, ldo);
}

// outg[m][n] = sum_{r<nr} A[r][m] * B[r][n]  (A, B in shared memory; outg in global memory, [M][Nn])
__device__ __forceinline__ void db_gemm_AtB(const float* __restrict__ A, int lda, int M, const float* __restrict__ B,
                                            int ldb, int Nn, int nr, float* __restrict__ outg) {
  const int ctn = Nn >> 2, ntiles = ctn * (M >> 2);
  for (int t = threadIdx.x; t < ntiles; t += DB_THREADS) {
    const int mt = t / ctn, ct = t - mt * ctn;
    const int m = mt * 4, n = ct * 4;
    float4 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int r = 0; r < nr; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(A + r * lda + m);
      const float4 b = *reinterpret_cast<const float4*>(B + r * ldb + n);
      acc[0].x = fmaf(a.x, b.x, acc[0].x); acc[0].y = fmaf(a.x, b.y, acc[0].y);
      acc[0].z = fmaf(a.x, b.z, acc[0].z); acc[0].w = fmaf(a.x, b.w, acc[0].w);
      acc[1].x = fmaf(a.y, b.x, acc[1].x); acc[1].y = fmaf(a.y, b.y, acc[1].y);
      acc[1].z = fmaf(a.y, b.z, acc[1].z); acc[1].w = fmaf(a.y, b.w, acc[1].w);
      acc[2].x = fmaf(a.z, b.x, acc[2].x); acc[2].y = fmaf(a.z, b.y, acc[2].y);
      acc[2].z = fmaf(a.z, b.z, acc[2].z); acc[2].w = fmaf(a.z, b.w, acc[2].w);
      acc[3].x = fmaf(a.w, b.x, acc[3].x); acc[3].y = fmaf(a.w, b.y, acc[3].y);
      acc[3].z = fmaf(a.w, b.z, acc[3].z); acc[3].w = fmaf(a.w, b.w, acc[3].w);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) __stcg(reinterpret_cast<float4*>(outg + (size_t)(m + j) * Nn + n), acc[j]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Tensor-core variants of the three slab GEMMs: mma.sync m16n8k8 TF32 with error compensation ("3xTF32":
// x = hi + lo with hi = tf32(x), lo = tf32(x - hi);  a*b ~ hi*hi + hi*lo + lo*hi, the two small terms in their own
// accumulator) -- the dropped lo*lo term is 2^-22 relative, the same order as the rounding of an fp32 FMA chain of
// length 104, so the 1e-5 parity bar holds (tests/test_dense_gpu.py runs both paths).  Why: with 4x4 register tiles the
// SIMT phases are bound by shared-memory bandwidth (8 LDS.128 per 64 FMA: 3.4 us per 36x104x104 product); a warp-level
// MMA needs 6 LDS.32 per 1024 MACs.  Same shared-memory layouts as the SIMT path except the forward weights, which
// are kept [Co][Kd+4] (un-swizzled; +4 makes the B-fragment loads conflict-free).  Channels must be multiples of 8.
// Rows / columns of a 16-wide tile that fall outside the operand read whatever follows it in shared memory (always
// inside the kernel's allocation) and are never stored; a product row depends on its own operand row only.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void db_tf32_split(float x, unsigned& hi, unsigned& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void db_mma_tf32(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
constexpr int DB_MMA_NT = 3;     // 8-column tiles per warp unit: the A fragment (and its split) is reused across them

// C[m][n] = init(n) + sum_k a(m,k) b(k,n) over Mt x Nt8 tiles of 16 x 8, Ksteps steps of 8.  fa(m,k), fb(k,n) fetch one
// operand element; fstore(m, n, v0, v1) receives C[m][n], C[m][n+1].  A unit = one 16-row tile x DB_MMA_NT column tiles.
template <typename FA, typename FB, typename FI, typename FS>
__device__ __forceinline__ void db_mma_gemm(int Mt, int Nt8, int Ksteps, FA fa, FB fb, FI finit, FS fstore) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ngroups = (Nt8 + DB_MMA_NT - 1) / DB_MMA_NT;
  for (int u = warp; u < Mt * ngroups; u += DB_WARPS) {
    const int mt = u % Mt, ng = u / Mt;
    const int m0 = mt * 16, nt0 = ng * DB_MMA_NT;
    float accM[DB_MMA_NT][4], accS[DB_MMA_NT][4];
#pragma unroll
    for (int j = 0; j < DB_MMA_NT; ++j) {
      const int n = (nt0 + j) * 8 + 2 * t;
      const bool on = nt0 + j < Nt8;
      accM[j][0] = accM[j][2] = on ? finit(n) : 0.f;
      accM[j][1] = accM[j][3] = on ? finit(n + 1) : 0.f;
      accS[j][0] = accS[j][1] = accS[j][2] = accS[j][3] = 0.f;
    }
    for (int ks = 0; ks < Ksteps; ++ks) {
      const int k0 = ks * 8;
      unsigned ah[4], al[4];
      db_tf32_split(fa(m0 + g, k0 + t), ah[0], al[0]);
      db_tf32_split(fa(m0 + g + 8, k0 + t), ah[1], al[1]);
      db_tf32_split(fa(m0 + g, k0 + t + 4), ah[2], al[2]);
      db_tf32_split(fa(m0 + g + 8, k0 + t + 4), ah[3], al[3]);
#pragma unroll
      for (int j = 0; j < DB_MMA_NT; ++j) {
        if (nt0 + j < Nt8) {                          // warp-uniform
          const int n = (nt0 + j) * 8 + g;
          unsigned bh0, bl0, bh1, bl1;
          db_tf32_split(fb(k0 + t, n), bh0, bl0);
          db_tf32_split(fb(k0 + t + 4, n), bh1, bl1);
          db_mma_tf32(accS[j], al, bh0, bh1);
          db_mma_tf32(accS[j], ah, bl0, bl1);
          db_mma_tf32(accM[j], ah, bh0, bh1);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < DB_MMA_NT; ++j)
      if (nt0 + j < Nt8) {
        const int n = (nt0 + j) * 8 + 2 * t;
        fstore(m0 + g, n, accM[j][0] + accS[j][0], accM[j][1] + accS[j][1]);
        fstore(m0 + g + 8, n, accM[j][2] + accS[j][2], accM[j][3] + accS[j][3]);
      }
  }
}

__host__ __device__ __forceinline__ int db_wstride_mma(int Kd) { return Kd + 4; }
// forward weights for the MMA path: row o of W [Co][Kd] at stride Kd+4 (16-byte chunks in order)
__device__ __forceinline__ void db_cp_weight_padded(const float* __restrict__ W, int Co, int Kd, float* __restrict__ Ws) {
  const int ck = Kd >> 2, ws = db_wstride_mma(Kd);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < Co; o += DB_WARPS)
    for (int kc = lane; kc < ck; kc += 32) db_cp16(Ws + o * ws + (kc << 2), W + (size_t)o * Kd + (kc << 2));
}
// out[r][o] = bias[o] + sum_k A[r][k] W[o][k], W at stride Kd+4
__device__ __forceinline__ void db_mma_AWt(const float* __restrict__ A, int lda, int Kd, const float* __restrict__ Ws,
                                           int Co, const float* __restrict__ bias, int nrp, float* __restrict__ out,
                                           int ldo) {
  const int ws = db_wstride_mma(Kd);
  db_mma_gemm((nrp + 15) >> 4, Co >> 3, Kd >> 3,
              [&](int r, int k) { return A[r * lda + k]; }, [&](int k, int n) { return Ws[n * ws + k]; },
              [&](int n) { return __ldg(bias + n); },
              [&](int r, int n, float v0, float v1) {
                if (r < nrp) *reinterpret_cast<float2*>(out + r * ldo + n) = make_float2(v0, v1);
              });
}
// out[r][n] = sum_k A[r][k] B[k][n]
__device__ __forceinline__ void db_mma_AB(const float* __restrict__ A, int lda, int Kd, const float* __restrict__ B,
                                          int Nn, int nrp, float* __restrict__ out, int ldo) {
  db_mma_gemm((nrp + 15) >> 4, Nn >> 3, Kd >> 3,
              [&](int r, int k) { return A[r * lda + k]; }, [&](int k, int n) { return B[k * Nn + n]; },
              [&](int) { return 0.f; },
              [&](int r, int n, float v0, float v1) {
                if (r < nrp) *reinterpret_cast<float2*>(out + r * ldo + n) = make_float2(v0, v1);
              });
}
// outg[m][n] = sum_{r<nr} A[r][m] B[r][n]  (global partial); operand rows >= nr are masked to zero
__device__ __forceinline__ void db_mma_AtB(const float* __restrict__ A, int lda, int M, const float* __restrict__ B,
                                           int ldb, int Nn, int nr, float* __restrict__ outg) {
  db_mma_gemm((M + 15) >> 4, Nn >> 3, (nr + 7) >> 3,
              [&](int m, int k) { return k < nr ? A[k * lda + m] : 0.f; },
              [&](int k, int n) { return k < nr ? B[k * ldb + n] : 0.f; }, [&](int) { return 0.f; },
              [&](int m, int n, float v0, float v1) {
                if (m < M) __stcg(reinterpret_cast<float2*>(outg + (size_t)m * Nn + n), make_float2(v0, v1));
              });
}

__device__ __forceinline__ void db_store_slab(const float* __restrict__ s, int r0, int nr, int C, float* __restrict__ g) {
  float* dst = g + (size_t)r0 * C;                      // the slab's rows are contiguous in the [N][C] matrix
  for (int i = threadIdx.x * 4; i < nr * C; i += DB_THREADS * 4)
    *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(s + i);
}

// Two column sums over the slab's rows, row-parallel: warp w takes rows w, w+16, ...; lane l the channels 4l..4l+3.
// F(r, c) -> (float4 u, float4 v) contributions; per-warp partials go through `red` [DB_WARPS][2][C]; thread c < C
// then adds the DB_WARPS partials in order and hands (sum_u, sum_v) to FIN(c, su, sv).  One __syncthreads inside.
template <typename F, typename FIN>
__device__ __forceinline__ void db_col_reduce2(int C, int nrows, float* __restrict__ red, F f, FIN fin) {
  const int warp = threadIdx.x >> 5, c = (threadIdx.x & 31) * 4;
  if (c < C) {
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f), v = u;
    for (int r = warp; r < nrows; r += DB_WARPS) f(r, c, u, v);
    *reinterpret_cast<float4*>(red + (warp * 2) * C + c) = u;
    *reinterpret_cast<float4*>(red + (warp * 2 + 1) * C + c) = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += DB_THREADS) {
    float su = 0.f, sv = 0.f;
#pragma unroll
    for (int w = 0; w < DB_WARPS; ++w) {
      su += red[(w * 2) * C + i];
      sv += red[(w * 2 + 1) * C + i];
    }
    fin(i, su, sv);
  }
}

// per-slab column statistics: psum[c] = sum_r S[r][c],  pm2[c] = sum_r (S[r][c] - slab mean)^2   (global partials).
// Single pass over data shifted by the slab's first row (a sample of the column, so no catastrophic cancellation):
// with d = x - K:  sum = sum(d) + n K,  M2 = sum(d^2) - sum(d)^2 / n.
__device__ __forceinline__ void db_slab_stats(const float* __restrict__ S, int C, int nr, float* __restrict__ red,
                                              float* __restrict__ psum, float* __restrict__ pm2) {
  db_col_reduce2(
      C, nr, red,
      [&](int r, int c, float4& u, float4& v) {
        const float4 k = *reinterpret_cast<const float4*>(S + c);
        const float4 x = *reinterpret_cast<const float4*>(S + r * C + c);
        const float dx = x.x - k.x, dy = x.y - k.y, dz = x.z - k.z, dw = x.w - k.w;
        u.x += dx; u.y += dy; u.z += dz; u.w += dw;
        v.x = fmaf(dx, dx, v.x); v.y = fmaf(dy, dy, v.y); v.z = fmaf(dz, dz, v.z); v.w = fmaf(dw, dw, v.w);
      },
      [&](int c, float su, float sv) {
        const float fn = (float)nr;
        __stcg(psum + c, nr > 0 ? fmaf(fn, S[c], su) : 0.f);
        __stcg(pm2 + c, nr > 0 ? fmaxf(sv - su * su / fn, 0.f) : 0.f);
      });
}

// Merge the per-CTA partials [grid][2][C] into batch mean / inverse std (Chan et al.), identically in every CTA:
//   mean = K + sum_b n_b d_b / N,   M2 = sum_b M2_b + sum_b n_b d_b^2 - (sum_b n_b d_b)^2 / N,   d_b = mean_b - K
// with the pivot K = mean of slab 0 (shifted data: single pass, no catastrophic cancellation).  Every partial this
// thread needs is requested before the first one is used: one L2 round trip.  red: shared [DB_WARPS][3][C].
__device__ __forceinline__ void db_merge_stats(const float* __restrict__ part, int grid, int Rc, int N, int C,
                                               float eps, float* __restrict__ red, float* __restrict__ mean_s,
                                               float* __restrict__ istd_s, float* __restrict__ var_s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = lane * 4;
  const bool on = c < C;
  float4 ps[DB_MAXI], pm[DB_MAXI];
  float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (on) p0 = __ldcg(reinterpret_cast<const float4*>(part + c));
#pragma unroll
  for (int i = 0; i < DB_MAXI; ++i) {
    const int b = warp + i * DB_WARPS;
    ps[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    pm[i] = ps[i];
    if (on && b < grid) {
      ps[i] = __ldcg(reinterpret_cast<const float4*>(part + (size_t)b * 2 * C + c));
      pm[i] = __ldcg(reinterpret_cast<const float4*>(part + (size_t)b * 2 * C + C + c));
    }
  }
  // rows of slab b: Rc for the full ones, fewer for the one that holds row N-1, none behind it (the grid is sized for
  // the row CAPACITY when the row count comes from device memory, kp_dense_desc.n_dev)
  const float fn0 = (float)min(Rc, N);
  const float inv0 = 1.f / fn0;
  const float4 K = make_float4(p0.x * inv0, p0.y * inv0, p0.z * inv0, p0.w * inv0);
  float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa, sm = sa;
#pragma unroll
  for (int i = 0; i < DB_MAXI; ++i) {
    const int b = warp + i * DB_WARPS;
    const int nb = min(Rc, N - b * Rc);
    if (b < grid && nb > 0) {
      const float fn = (float)nb, inv = 1.f / fn;
      const float dx = fmaf(ps[i].x, inv, -K.x), dy = fmaf(ps[i].y, inv, -K.y), dz = fmaf(ps[i].z, inv, -K.z),
                  dw = fmaf(ps[i].w, inv, -K.w);
      sa.x = fmaf(fn, dx, sa.x); sa.y = fmaf(fn, dy, sa.y); sa.z = fmaf(fn, dz, sa.z); sa.w = fmaf(fn, dw, sa.w);
      sb.x = fmaf(fn * dx, dx, sb.x); sb.y = fmaf(fn * dy, dy, sb.y);
      sb.z = fmaf(fn * dz, dz, sb.z); sb.w = fmaf(fn * dw, dw, sb.w);
      sm.x += pm[i].x; sm.y += pm[i].y; sm.z += pm[i].z; sm.w += pm[i].w;
    }
  }
  if (on) {
    *reinterpret_cast<float4*>(red + (warp * 3) * C + c) = sa;
    *reinterpret_cast<float4*>(red + (warp * 3 + 1) * C + c) = sb;
    *reinterpret_cast<float4*>(red + (warp * 3 + 2) * C + c) = sm;
    if (warp == 0) *reinterpret_cast<float4*>(mean_s + c) = K;          // pivot, replaced by the mean below
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += DB_THREADS) {
    float a = 0.f, b2 = 0.f, m2 = 0.f;
#pragma unroll
    for (int w = 0; w < DB_WARPS; ++w) {
      a += red[(w * 3) * C + i];
      b2 += red[(w * 3 + 1) * C + i];
      m2 += red[(w * 3 + 2) * C + i];
    }
    const float invn = 1.f / (float)N;
    const float var = fmaxf((m2 + (b2 - a * a * invn)) * invn, 0.f);
    mean_s[i] = fmaf(a, invn, mean_s[i]);
    var_s[i] = var;
    istd_s[i] = rsqrtf(var + eps);
  }
  __syncthreads();
}

// running statistics + saved statistics, by CTA 0 only (torch.nn.BatchNorm1d: unbiased variance into running_var)
__device__ __forceinline__ void db_publish_stats(const float* mean_s, const float* istd_s, const float* var_s, int N,
                                                 int C, float mom, float* rm, float* rv, long long* nbt,
                                                 float* save_mean, float* save_istd) {
  const float unb = N > 1 ? (float)N / (float)(N - 1) : 1.f;
  for (int i = threadIdx.x; i < C; i += DB_THREADS) {
    save_mean[i] = mean_s[i];
    save_istd[i] = istd_s[i];
    if (rm) rm[i] = fmaf(mom, mean_s[i] - rm[i], rm[i]);
    if (rv) rv[i] = fmaf(mom, var_s[i] * unb - rv[i], rv[i]);
  }
  if (nbt && threadIdx.x == 0) *nbt += 1;
}

// S[r][c] <- relu?( g[c] * (S[r][c] - mean[c]) * istd[c] + be[c] ) for r < nr (rows >= nr stay zero)
template <bool RELU>
__device__ __forceinline__ void db_bn_apply(const float* __restrict__ S, int C, int nr, const float* __restrict__ g,
                                            const float* __restrict__ be, const float* __restrict__ mean_s,
                                            const float* __restrict__ istd_s, float* __restrict__ D) {
  const int warp = threadIdx.x >> 5, c = (threadIdx.x & 31) * 4;
  if (c >= C) return;
  const float4 mu = *reinterpret_cast<const float4*>(mean_s + c);
  const float4 is = *reinterpret_cast<const float4*>(istd_s + c);
  const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
  const float4 bb = __ldg(reinterpret_cast<const float4*>(be + c));
  for (int r = warp; r < nr; r += DB_WARPS) {
    const float4 v = *reinterpret_cast<const float4*>(S + r * C + c);
    // ((v - mean) * invstd) * gamma + beta, in this order: the backward recomputes x-hat = (v - mean) * invstd
    float4 o = make_float4(fmaf((v.x - mu.x) * is.x, gg.x, bb.x), fmaf((v.y - mu.y) * is.y, gg.y, bb.y),
                           fmaf((v.z - mu.z) * is.z, gg.z, bb.z), fmaf((v.w - mu.w) * is.w, gg.w, bb.w));
    if (RELU) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    *reinterpret_cast<float4*>(D + r * C + c) = o;
  }
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DB_THREADS, 1)
dense_block_fwd_kernel(const kp_dense_desc m, float* __restrict__ out, float* __restrict__ part, unsigned* bar, int Rc,
                       int mma) {
  extern __shared__ __align__(16) float smem[];
  DB_T_DECL
  DB_T(0);
  const int Ci = m.Cin, Co = m.Cout;
  // rows: m.N is the CAPACITY of the row buffers (and sizes the grid); the batch's row count may come from device
  // memory (m.n_dev), so that one captured launch serves batches of different sizes.  Rows in [N, m.N) are padding:
  // excluded from every statistic, written as zeros.
  const int N = m.n_dev ? min(__ldg(m.n_dev), m.N) : m.N;
  const int grid = gridDim.x;
  const int r0 = blockIdx.x * Rc;
  const int nr = max(0, min(Rc, N - r0));
  const int nr_cap = max(0, min(Rc, m.N - r0));
  const int Rp = (Rc + 3) & ~3;
  float* W1s = smem;                         // [Co][wstride(Ci)] swizzled  (MMA path: [Co][Ci+4])
  float* W2s = W1s + Co * db_wstride_any(Ci);    // [Co][wstride(Co)] swizzled  (MMA path: [Co][Co+4])
  float* A = W2s + Co * db_wstride_any(Co);  // [Rp][Ci]
  float* Y = A + Rp * Ci;                    // [Rp][Co]
  float* Z = Y + Rp * Co;                    // [Rp][Co]
  float* red = Z + Rp * Co;                  // [DB_WARPS][3][Co]
  float* st = red + DB_WARPS * 3 * Co;       // [3][3][Co]: (mean, istd, var) of BN1, BN2, BN3
  // everything this CTA will read from global memory except the partials is requested now, asynchronously
  // The first weight matrix does not depend on the preceding kernel: under programmatic dependent launch
  // (KP_DENSE_PDL=1) its copy is in flight while that kernel drains; the X slab (the predecessor's output) is requested
  // behind kp_pdl_wait().  The row count was written at the head of the step.
  if (mma) db_cp_weight_padded(m.W1, Co, Ci, W1s);
  else db_cp_weight_swizzled(m.W1, Co, Ci, W1s);
  db_cp_commit();
  kp_pdl_wait();
  kp_pdl_trigger();      // dependents (the next aggregation kernel) may run their prologue from here on
  db_cp_slab(m.X, r0, nr, Rp, Ci, A);
  db_cp_commit();
  if (mma) db_cp_weight_padded(m.W2, Co, Co, W2s);       // third group: lands behind the first GEMM
  else db_cp_weight_swizzled(m.W2, Co, Co, W2s);
  db_cp_commit();
  for (int i = threadIdx.x * 4; i < Rp * Co; i += DB_THREADS * 4)
    *reinterpret_cast<float4*>(Z + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("cp.async.wait_group 1;" ::: "memory");
  __syncthreads();
  DB_T(1);
  float* part0 = part;
  float* part1 = part + (size_t)grid * 2 * Co;
  float* part2 = part1 + (size_t)grid * 2 * Co;

  // ---- Linear1 + BN1 + ReLU ----
  if (mma) db_mma_AWt(A, Ci, Ci, W1s, Co, m.b1, Rp, Y, Co);
  else db_gemm_AWt(A, Ci, Ci, W1s, Co, m.b1, Rp, Y, Co);
  __syncthreads();
  DB_T(2);
  db_slab_stats(Y, Co, nr, red, part0 + (size_t)blockIdx.x * 2 * Co, part0 + (size_t)blockIdx.x * 2 * Co + Co);
  db_store_slab(Y, r0, nr, Co, m.Y1);
  DB_T(3);
  db_grid_barrier(bar, 1u * grid);
  DB_T(4);
  db_merge_stats(part0, grid, Rc, N, Co, m.eps1, red, st, st + Co, st + 2 * Co);
  DB_T(5);
  db_bn_apply<true>(Y, Co, nr, m.g1, m.be1, st, st + Co, Z);
  db_cp_wait_all();
  __syncthreads();
  DB_T(6);

  // ---- Linear2 + BN2 + ReLU ----
  if (mma) db_mma_AWt(Z, Co, Co, W2s, Co, m.b2, Rp, Y, Co);
  else db_gemm_AWt(Z, Co, Co, W2s, Co, m.b2, Rp, Y, Co);
  __syncthreads();
  DB_T(7);
  db_slab_stats(Y, Co, nr, red, part1 + (size_t)blockIdx.x * 2 * Co, part1 + (size_t)blockIdx.x * 2 * Co + Co);
  db_store_slab(Y, r0, nr, Co, m.Y2);
  DB_T(8);
  db_grid_barrier(bar, 2u * grid);
  DB_T(9);
  db_merge_stats(part1, grid, Rc, N, Co, m.eps2, red, st + 3 * Co, st + 4 * Co, st + 5 * Co);
  DB_T(10);
  db_bn_apply<true>(Y, Co, nr, m.g2, m.be2, st + 3 * Co, st + 4 * Co, Z);
  __syncthreads();
  DB_T(11);

  // ---- outer BatchNorm + residual ----
  if (m.g3) {
    db_slab_stats(Z, Co, nr, red, part2 + (size_t)blockIdx.x * 2 * Co, part2 + (size_t)blockIdx.x * 2 * Co + Co);
    db_store_slab(Z, r0, nr, Co, m.Z2);
    DB_T(12);
    db_grid_barrier(bar, 3u * grid);
    DB_T(13);
    db_merge_stats(part2, grid, Rc, N, Co, m.eps3, red, st + 6 * Co, st + 7 * Co, st + 8 * Co);
    DB_T(14);
    db_bn_apply<false>(Z, Co, nr, m.g3, m.be3, st + 6 * Co, st + 7 * Co, Y);
    __syncthreads();
    DB_T(15);
  }
  const float* res = m.g3 ? Y : Z;
  {
    const size_t so = m.out_stride ? (size_t)m.out_stride : (size_t)Co, sr = m.r_stride ? (size_t)m.r_stride : (size_t)Co;
    const int warp = threadIdx.x >> 5, c = (threadIdx.x & 31) * 4;
    if (c < Co)
      for (int r = warp; r < nr; r += DB_WARPS) {
        float4 v = *reinterpret_cast<const float4*>(res + r * Co + c);
        if (m.R) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(m.R + (size_t)(r0 + r) * sr + c));
          v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
        }
        *reinterpret_cast<float4*>(out + (size_t)(r0 + r) * so + c) = v;
      }
    if (c < Co)
      for (int r = nr + warp; r < nr_cap; r += DB_WARPS)       // padding rows of the capacity: zeros
        *reinterpret_cast<float4*>(out + (size_t)(r0 + r) * so + c) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  DB_T(16);
  DB_T_PRINT(17, "fwd load gemm1 stats1 bar1 merge1 apply1 gemm2 stats2 bar2 merge2 apply2 stats3 bar3 merge3 apply3 out");
  if (threadIdx.x == 0 && atomicAdd(bar + 1, 1u) == gridDim.x - 1) {   // every CTA is past its last barrier: leave zeros
    bar[0] = 0u;
    bar[1] = 0u;
  }
  // saved + running statistics: by the last CTA (the shortest slab), off everybody's critical path
  if (blockIdx.x == grid - 1) {
    db_publish_stats(st, st + Co, st + 2 * Co, N, Co, m.mom1, m.rm1, m.rv1, (long long*)m.nbt1, m.stats, m.stats + Co);
    db_publish_stats(st + 3 * Co, st + 4 * Co, st + 5 * Co, N, Co, m.mom2, m.rm2, m.rv2, (long long*)m.nbt2,
                     m.stats + 2 * Co, m.stats + 3 * Co);
    if (m.g3)
      db_publish_stats(st + 6 * Co, st + 7 * Co, st + 8 * Co, N, Co, m.mom3, m.rm3, m.rv3, (long long*)m.nbt3,
                       m.stats + 4 * Co, m.stats + 5 * Co);
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
// slab partial sums p1[c] = sum_r D[r][c], p2[c] = sum_r D[r][c] * XH[r][c]  -> global partial [2][C]
__device__ __forceinline__ void db_slab_dots(const float* __restrict__ D, const float* __restrict__ XH, int C, int nr,
                                             float* __restrict__ red, float* __restrict__ p) {
  db_col_reduce2(
      C, nr, red,
      [&](int r, int c, float4& u, float4& v) {
        const float4 d = *reinterpret_cast<const float4*>(D + r * C + c);
        const float4 x = *reinterpret_cast<const float4*>(XH + r * C + c);
        u.x += d.x; u.y += d.y; u.z += d.z; u.w += d.w;
        v.x = fmaf(d.x, x.x, v.x); v.y = fmaf(d.y, x.y, v.y); v.z = fmaf(d.z, x.z, v.z); v.w = fmaf(d.w, x.w, v.w);
      },
      [&](int c, float su, float sv) {
        __stcg(p + c, su);
        __stcg(p + C + c, sv);
      });
}
// plain fixed-order sums of the [grid][2][C] partials into shared s1[C], s2[C]
__device__ __forceinline__ void db_merge_sums(const float* __restrict__ part, int grid, int C, float* __restrict__ red,
                                              float* __restrict__ s1, float* __restrict__ s2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = lane * 4;
  const bool on = c < C;
  float4 pa[DB_MAXI], pb[DB_MAXI];                       // all requested up front: one L2 round trip
#pragma unroll
  for (int i = 0; i < DB_MAXI; ++i) {
    const int b = warp + i * DB_WARPS;
    pa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    pb[i] = pa[i];
    if (on && b < grid) {
      pa[i] = __ldcg(reinterpret_cast<const float4*>(part + (size_t)b * 2 * C + c));
      pb[i] = __ldcg(reinterpret_cast<const float4*>(part + (size_t)b * 2 * C + C + c));
    }
  }
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a;
#pragma unroll
  for (int i = 0; i < DB_MAXI; ++i) {
    a.x += pa[i].x; a.y += pa[i].y; a.z += pa[i].z; a.w += pa[i].w;
    b4.x += pb[i].x; b4.y += pb[i].y; b4.z += pb[i].z; b4.w += pb[i].w;
  }
  if (on) {
    *reinterpret_cast<float4*>(red + warp * 2 * C + c) = a;
    *reinterpret_cast<float4*>(red + warp * 2 * C + C + c) = b4;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += DB_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < DB_WARPS; ++w) t += red[w * 2 * C + i];
    if (i < C) s1[i] = t;
    else s2[i - C] = t;
  }
  __syncthreads();
}
// XH <- (S - mean) * istd   (normalised activations of the slab), all nrp rows
__device__ __forceinline__ void db_xhat(const float* __restrict__ S, int C, int nrp, const float* __restrict__ mean,
                                        const float* __restrict__ istd, float* __restrict__ XH) {
  const int warp = threadIdx.x >> 5, c = (threadIdx.x & 31) * 4;
  if (c >= C) return;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
  const float4 is = __ldg(reinterpret_cast<const float4*>(istd + c));
  for (int r = warp; r < nrp; r += DB_WARPS) {
    const float4 v = *reinterpret_cast<const float4*>(S + r * C + c);
    *reinterpret_cast<float4*>(XH + r * C + c) =
        make_float4((v.x - mu.x) * is.x, (v.y - mu.y) * is.y, (v.z - mu.z) * is.z, (v.w - mu.w) * is.w);
  }
}
// D <- g * istd * (D - s1/N - XH * s2/N)   for r < nr
__device__ __forceinline__ void db_bn_bwd_apply(float* __restrict__ D, const float* __restrict__ XH, int C, int nr, int N,
                                                const float* __restrict__ g, const float* __restrict__ istd,
                                                const float* __restrict__ s1, const float* __restrict__ s2) {
  const int warp = threadIdx.x >> 5, c = (threadIdx.x & 31) * 4;
  if (c >= C) return;
  const float invn = 1.f / (float)N;
  const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
  const float4 is = __ldg(reinterpret_cast<const float4*>(istd + c));
  const float4 a = *reinterpret_cast<const float4*>(s1 + c);
  const float4 b = *reinterpret_cast<const float4*>(s2 + c);
  const float4 k = make_float4(gg.x * is.x, gg.y * is.y, gg.z * is.z, gg.w * is.w);
  const float4 m1 = make_float4(a.x * invn, a.y * invn, a.z * invn, a.w * invn);
  const float4 m2 = make_float4(b.x * invn, b.y * invn, b.z * invn, b.w * invn);
  for (int r = warp; r < nr; r += DB_WARPS) {
    const float4 dv = *reinterpret_cast<const float4*>(D + r * C + c);
    const float4 xh = *reinterpret_cast<const float4*>(XH + r * C + c);
    *reinterpret_cast<float4*>(D + r * C + c) =
        make_float4(k.x * (dv.x - m1.x - xh.x * m2.x), k.y * (dv.y - m1.y - xh.y * m2.y),
                    k.z * (dv.z - m1.z - xh.z * m2.z), k.w * (dv.w - m1.w - xh.w * m2.w));
  }
}
// D[r][c] <- 0 where g*XH + be <= 0 (the ReLU after the BatchNorm was inactive)
__device__ __forceinline__ void db_relu_mask(float* __restrict__ D, const float* __restrict__ XH, int C, int nr,
                                             const float* __restrict__ g, const float* __restrict__ be) {
  const int warp = threadIdx.x >> 5, c = (threadIdx.x & 31) * 4;
  if (c >= C) return;
  const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
  const float4 bb = __ldg(reinterpret_cast<const float4*>(be + c));
  for (int r = warp; r < nr; r += DB_WARPS) {
    float4 dv = *reinterpret_cast<const float4*>(D + r * C + c);
    const float4 xh = *reinterpret_cast<const float4*>(XH + r * C + c);
    if (fmaf(gg.x, xh.x, bb.x) <= 0.f) dv.x = 0.f;
    if (fmaf(gg.y, xh.y, bb.y) <= 0.f) dv.y = 0.f;
    if (fmaf(gg.z, xh.z, bb.z) <= 0.f) dv.z = 0.f;
    if (fmaf(gg.w, xh.w, bb.w) <= 0.f) dv.w = 0.f;
    *reinterpret_cast<float4*>(D + r * C + c) = dv;
  }
}
// column sums of a slab -> global partial [C]
__device__ __forceinline__ void db_slab_colsum(const float* __restrict__ A, int C, int nr, float* __restrict__ red,
                                               float* __restrict__ pA) {
  db_col_reduce2(
      C, nr, red,
      [&](int r, int c, float4& u, float4& v) {
        const float4 a = *reinterpret_cast<const float4*>(A + r * C + c);
        u.x += a.x; u.y += a.y; u.z += a.z; u.w += a.w;
      },
      [&](int c, float su, float sv) { __stcg(pA + c, su); });
}

// workspace layout (floats): pa, pb, pc [grid][2][Co] | pW1 [grid][Co*Ci] | pW2 [grid][Co*Co] | pb1, pb2 [grid][Co]
__global__ void __launch_bounds__(DB_THREADS, 1)
dense_block_bwd_kernel(const kp_dense_desc m, const float* __restrict__ dOut, float* __restrict__ dX,
                       float* __restrict__ dW1, float* __restrict__ db1, float* __restrict__ dW2,
                       float* __restrict__ db2, float* __restrict__ dbn, float* __restrict__ ws, unsigned* bar, int Rc,
                       int mma) {
  extern __shared__ __align__(16) float smem[];
  DB_T_DECL
  DB_T(0);
  const int Ci = m.Cin, Co = m.Cout;
  const int N = m.n_dev ? min(__ldg(m.n_dev), m.N) : m.N;      // see dense_block_fwd_kernel
  const int grid = gridDim.x;
  const int r0 = blockIdx.x * Rc;
  const int nr = max(0, min(Rc, N - r0));
  const int nr_cap = max(0, min(Rc, m.N - r0));
  const int Rp = (Rc + 3) & ~3;
  float* W1 = smem;                   // [Co][Ci] row-major (dX = dy1 W1)
  float* W2 = W1 + Co * Ci;           // [Co][Co] row-major (dz1 = dy2 W2)
  float* D = W2 + Co * Co;            // [Rp][Co]  dOut -> dz2 -> dy2, finally dX ([Rp][Ci], Ci <= Co)
  float* X3 = D + Rp * Co;            // [Rp][Co]  z2 -> x-hat3
  float* X2 = X3 + Rp * Co;           // [Rp][Co]  y2 -> x-hat2
  float* X1 = X2 + Rp * Co;           // [Rp][Co]  y1 -> x-hat1
  float* Z1 = X1 + Rp * Co;           // [Rp][Co]  z1 = relu(BN1(y1))
  float* E = Z1 + Rp * Co;            // [Rp][Co]  dz1 -> dy1
  float* XS = E + Rp * Co;            // [Rp][Ci]  X slab
  float* red = XS + Rp * Ci;          // [DB_WARPS][2*Co]
  float* s1 = red + DB_WARPS * 2 * Co;
  float* s2 = s1 + Co;
  float* pa = ws;
  float* pb = pa + (size_t)grid * 2 * Co;
  float* pc = pb + (size_t)grid * 2 * Co;
  float* pW1 = pc + (size_t)grid * 2 * Co;
  float* pW2 = pW1 + (size_t)grid * Co * Ci;
  float* pb1 = pW2 + (size_t)grid * Co * Co;
  float* pb2 = pb1 + (size_t)grid * Co;
  const float* mean1 = m.stats, *istd1 = m.stats + Co, *mean2 = m.stats + 2 * Co, *istd2 = m.stats + 3 * Co;
  const float* mean3 = m.stats + 4 * Co, *istd3 = m.stats + 5 * Co;

  // every slab and both weight matrices are requested now, asynchronously: one round trip to L2 / HBM.  Saved
  // activations and weights first -- they do not depend on the preceding kernel, so under programmatic dependent launch
  // (KP_DENSE_PDL=1) they are in flight while it drains -- then, behind kp_pdl_wait(), the incoming gradient.
  if (m.g3) db_cp_slab(m.Z2, r0, nr, Rp, Co, X3);
  db_cp_slab(m.Y2, r0, nr, Rp, Co, X2);
  db_cp_slab(m.Y1, r0, nr, Rp, Co, X1);
  db_cp_slab(m.X, r0, nr, Rp, Ci, XS);
  db_cp_rows(m.W2, Co * Co, W2);
  db_cp_rows(m.W1, Co * Ci, W1);
  db_cp_commit();
  kp_pdl_wait();
  kp_pdl_trigger();      // dependents (the next aggregation kernel) may run their prologue from here on
  {
    const size_t sd = m.dout_stride ? (size_t)m.dout_stride : (size_t)Co;
    const int warp = threadIdx.x >> 5, c = (threadIdx.x & 31) * 4;
    if (c < Co)
      for (int r = warp; r < Rp; r += DB_WARPS) {
        if (r < nr) db_cp16(D + r * Co + c, dOut + (size_t)(r0 + r) * sd + c);
        else *reinterpret_cast<float4*>(D + r * Co + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
  }
  db_cp_commit();
  db_cp_wait_all();
  __syncthreads();
  DB_T(1);
  if (m.dR) {                                   // residual: dR[row] += dOut[row] (rows are owned by exactly one CTA)
    const size_t sr = m.dr_stride ? (size_t)m.dr_stride : (size_t)Co;
    const int warp = threadIdx.x >> 5, c = (threadIdx.x & 31) * 4;
    if (c < Co)
      for (int r = warp; r < nr; r += DB_WARPS) {
        float4* p = reinterpret_cast<float4*>(m.dR + (size_t)(r0 + r) * sr + c);
        float4 a = *p;
        const float4 g = *reinterpret_cast<const float4*>(D + r * Co + c);
        a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
        *p = a;
      }
  }
  unsigned phase = 0;
  // x-hats of the three BatchNorms and z1, while nothing else can proceed anyway
  if (m.g3) db_xhat(X3, Co, Rp, mean3, istd3, X3);
  db_xhat(X2, Co, Rp, mean2, istd2, X2);
  db_xhat(X1, Co, Rp, mean1, istd1, X1);
  {
    const int warp = threadIdx.x >> 5, c = (threadIdx.x & 31) * 4;
    if (c < Co) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(m.g1 + c));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(m.be1 + c));
      for (int r = warp; r < Rp; r += DB_WARPS) {             // same (warp, lane) -> element mapping as db_xhat
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nr) {
          const float4 xh = *reinterpret_cast<const float4*>(X1 + r * Co + c);
          z = make_float4(fmaxf(fmaf(gg.x, xh.x, bb.x), 0.f), fmaxf(fmaf(gg.y, xh.y, bb.y), 0.f),
                          fmaxf(fmaf(gg.z, xh.z, bb.z), 0.f), fmaxf(fmaf(gg.w, xh.w, bb.w), 0.f));
        }
        *reinterpret_cast<float4*>(Z1 + r * Co + c) = z;
      }
    }
  }
  __syncthreads();
  DB_T(2);
  if (m.g3) {
    // ---- outer BatchNorm ----
    db_slab_dots(D, X3, Co, nr, red, pa + (size_t)blockIdx.x * 2 * Co);
    db_grid_barrier(bar, ++phase * grid);
    db_merge_sums(pa, grid, Co, red, s1, s2);
    if (blockIdx.x == grid - 1)
      for (int i = threadIdx.x; i < Co; i += DB_THREADS) {
        dbn[4 * Co + i] = s2[i];
        dbn[5 * Co + i] = s1[i];
      }
    db_bn_bwd_apply(D, X3, Co, nr, N, m.g3, istd3, s1, s2);
    __syncthreads();
  }
  DB_T(3);
  // ---- ReLU2 + BN2 ----
  db_relu_mask(D, X2, Co, nr, m.g2, m.be2);
  __syncthreads();
  db_slab_dots(D, X2, Co, nr, red, pb + (size_t)blockIdx.x * 2 * Co);
  db_grid_barrier(bar, ++phase * grid);
  DB_T(4);
  db_merge_sums(pb, grid, Co, red, s1, s2);
  if (blockIdx.x == grid - 1)
    for (int i = threadIdx.x; i < Co; i += DB_THREADS) {
      dbn[2 * Co + i] = s2[i];
      dbn[3 * Co + i] = s1[i];
    }
  db_bn_bwd_apply(D, X2, Co, nr, N, m.g2, istd2, s1, s2);      // D = dy2
  __syncthreads();
  DB_T(5);
  // ---- Linear2: dz1 = dy2 W2 first (critical path), then the dW2 / db2 partials ----
  if (mma) db_mma_AB(D, Co, Co, W2, Co, Rp, E, Co);
  else db_gemm_AB(D, Co, Co, W2, Co, Rp, E, Co);
  __syncthreads();
  DB_T(6);
  db_relu_mask(E, X1, Co, nr, m.g1, m.be1);
  __syncthreads();
  db_slab_dots(E, X1, Co, nr, red, pc + (size_t)blockIdx.x * 2 * Co);
  __syncthreads();
  if (threadIdx.x == 0) {                                       // arrive early, wait after the off-path work
    __threadfence();
    atomicAdd(bar, 1u);
  }
  ++phase;
  DB_T(7);
  if (mma) db_mma_AtB(D, Co, Co, Z1, Co, Co, nr, pW2 + (size_t)blockIdx.x * Co * Co);
  else db_gemm_AtB(D, Co, Co, Z1, Co, Co, nr, pW2 + (size_t)blockIdx.x * Co * Co);
  db_slab_colsum(D, Co, nr, red, pb2 + (size_t)blockIdx.x * Co);
  DB_T(8);
  if (threadIdx.x == 0) {
    while (db_ld_acquire(bar) < phase * grid) {
    }
    __threadfence();
  }
  __syncthreads();
  DB_T(9);
  db_merge_sums(pc, grid, Co, red, s1, s2);
  if (blockIdx.x == grid - 1)
    for (int i = threadIdx.x; i < Co; i += DB_THREADS) {
      dbn[0 * Co + i] = s2[i];
      dbn[1 * Co + i] = s1[i];
    }
  db_bn_bwd_apply(E, X1, Co, nr, N, m.g1, istd1, s1, s2);      // E = dy1
  __syncthreads();
  DB_T(10);
  // ---- Linear1: dX = dy1 W1, dW1 / db1 partials ----
  if (mma) db_mma_AB(E, Co, Co, W1, Ci, Rp, D, Ci);
  else db_gemm_AB(E, Co, Co, W1, Ci, Rp, D, Ci);
  DB_T(11);
  if (mma) db_mma_AtB(E, Co, Co, XS, Ci, Ci, nr, pW1 + (size_t)blockIdx.x * Co * Ci);
  else db_gemm_AtB(E, Co, Co, XS, Ci, Ci, nr, pW1 + (size_t)blockIdx.x * Co * Ci);
  db_slab_colsum(E, Co, nr, red, pb1 + (size_t)blockIdx.x * Co);
  __syncthreads();
  DB_T(12);
  db_store_slab(D, r0, nr_cap, Ci, dX);      // rows [nr, nr_cap) are padding: their dy1 rows are zero, so is dX
  DB_T(13);
  if (threadIdx.x == 0 && atomicAdd(bar + 1, 1u) == gridDim.x - 1) {
    bar[0] = 0u;
    bar[1] = 0u;
  }
  DB_T(14);
  DB_T_PRINT(15, "bwd load xhat bn3 mask2+dots2+bar merge2+apply2 gemm_dz1 mask1+dots1 dW2+colsum wait merge1+apply1 gemm_dX dW1+colsum store+bar... reduce");
}


// dW1, db1, dW2, db2 = fixed-order sums of the per-CTA partials of dense_block_bwd_kernel: 16 lanes per output float4,
// every partial requested before the first is used.  A separate launch: inside the persistent kernel this phase sat
// behind a fourth grid barrier and ran on the kernel's 82 CTAs only (7.6 us, profiles/r1z_dense_block_phases.txt);
// as its own grid it covers every output at once.
__global__ void __launch_bounds__(256)
dense_block_wgrad_reduce_kernel(const float* __restrict__ ws, int grid, int Ci, int Co, float* __restrict__ dW1,
                                float* __restrict__ db1, float* __restrict__ dW2, float* __restrict__ db2) {
  const float* pW1 = ws + (size_t)grid * 6 * Co;
  const float* pW2 = pW1 + (size_t)grid * Co * Ci;
  const float* pb1 = pW2 + (size_t)grid * Co * Co;
  const float* pb2 = pb1 + (size_t)grid * Co;
  constexpr int DB_THREADS_R = 256;
  {
    const int nW1 = (Co * Ci) >> 2, nW2 = (Co * Co) >> 2, nb = Co >> 2;
    const int total = nW1 + nW2 + 2 * nb;
    const int sub = threadIdx.x & 15;
    const int wg = (blockIdx.x * DB_THREADS_R + threadIdx.x) >> 5, nwg = ((int)gridDim.x * DB_THREADS_R) >> 5;
    for (int e0 = wg * 2; e0 < total; e0 += nwg * 2) {          // warp-uniform trip count (full-mask shuffles)
      const int e = e0 + ((threadIdx.x & 31) >> 4);
      const bool ok = e < total;
      const float* src = pW1;
      float* dst = dW1;
      size_t stride = 0;
      if (e < nW1) {
        src = pW1 + (size_t)e * 4; dst = dW1 + (size_t)e * 4; stride = (size_t)Co * Ci;
      } else if (e < nW1 + nW2) {
        src = pW2 + (size_t)(e - nW1) * 4; dst = dW2 + (size_t)(e - nW1) * 4; stride = (size_t)Co * Co;
      } else if (e < nW1 + nW2 + nb) {
        src = pb1 + (size_t)(e - nW1 - nW2) * 4; dst = db1 + (size_t)(e - nW1 - nW2) * 4; stride = Co;
      } else if (ok) {
        src = pb2 + (size_t)(e - nW1 - nW2 - nb) * 4; dst = db2 + (size_t)(e - nW1 - nW2 - nb) * 4; stride = Co;
      }
      float4 pv[DB_MAXI];                                        // DB_MAXI * 16 >= the largest grid
#pragma unroll
      for (int i = 0; i < DB_MAXI; ++i) {
        const int q = sub + i * 16;
        pv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && q < grid) pv[i] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)q * stride));
      }
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < DB_MAXI; ++i) {
        s.x += pv[i].x; s.y += pv[i].y; s.z += pv[i].z; s.w += pv[i].w;
      }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, o);
        s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
      }
      if (ok && sub == 0) *reinterpret_cast<float4*>(dst) = s;
    }
  }
}

struct DbCfg {
  int grid, Rc;
  size_t smem_fwd, smem_bwd, ws_fwd, ws_bwd;
};

static size_t db_smem_fwd(int Ci, int Co, int Rc) {
  const int Rp = (Rc + 3) & ~3;
  return sizeof(float) * ((size_t)Co * db_wstride_any(Ci) + (size_t)Co * db_wstride_any(Co) + (size_t)Rp * Ci +
                          2 * (size_t)Rp * Co + (size_t)DB_WARPS * 3 * Co + 9 * (size_t)Co);
}
static size_t db_smem_bwd(int Ci, int Co, int Rc) {
  const int Rp = (Rc + 3) & ~3;
  return sizeof(float) * ((size_t)Ci * Co + (size_t)Co * Co + 6 * (size_t)Rp * Co + (size_t)Rp * Ci +
                          (size_t)DB_WARPS * 2 * Co + 2 * (size_t)Co);
}
static const size_t kDbSmemBudget = 220 * 1024;

// tensor-core GEMM phases (3xTF32, channel counts multiples of 8): OPT-IN (KP_DENSE_MMA=1 / kp_dense_block_set_mma(1)).
// Measured inside the training step (profiles/r2_dense_mma.txt): as accurate as the fp32-FMA tiles (3.3e-7 vs 4.3e-7
// against float64 at the bench shape) but SLOWER -- forward 295-307 us vs 280 us per step, backward 353-387 vs 302-334:
// legacy mma.sync TF32 issues at a fraction of the tcgen05 rate, and three MMAs plus the register-side hi/lo split per
// 16x8x8 tile cost more issue slots than the 4x4 FMA tiles they replace.  A win would need tcgen05 (operands from
// shared-memory descriptors, hi/lo planes pre-split in shared memory, TMEM accumulators, M = 64 row tiles).
static int g_db_mma = -1;            // -1: environment / default; 0 / 1: kp_dense_block_set_mma
static int db_use_mma(const kp_dense_desc& m) {
  static const int env = getenv("KP_DENSE_MMA") ? atoi(getenv("KP_DENSE_MMA")) : 0;
  const int on = g_db_mma >= 0 ? g_db_mma : env;
  return (on && m.Cin % 8 == 0 && m.Cout % 8 == 0) ? 1 : 0;
}

static int db_max_rc(int Ci, int Co) {
  int rc = 0;
  for (int r = 4; r <= 256; r += 4)
    if (db_smem_fwd(Ci, Co, r) <= kDbSmemBudget && db_smem_bwd(Ci, Co, r) <= kDbSmemBudget) rc = r;
  return rc;
}

static int db_config(const kp_dense_desc& m, DbCfg* c) {
  KP_CHECK_ARG(m.N >= 2 && m.Cin >= 4 && m.Cout >= 4 && m.Cin % 4 == 0 && m.Cout % 4 == 0 && m.Cin <= 128 &&
                   m.Cout <= 128 && m.Cin <= m.Cout,
               "kp_dense_block: need N >= 2, channels multiples of 4, Cin <= Cout <= 128 (got N=%d Cin=%d Cout=%d)",
               m.N, m.Cin, m.Cout);
  KP_CHECK_ARG(m.out_stride % 4 == 0 && m.r_stride % 4 == 0 && m.dout_stride % 4 == 0 && m.dr_stride % 4 == 0 &&
                   m.out_stride >= 0 && m.r_stride >= 0 && m.dout_stride >= 0 && m.dr_stride >= 0 &&
                   (((uintptr_t)m.dR) & 15) == 0,
               "kp_dense_block: row strides must be non-negative multiples of 4 elements");
  const int rcmax = db_max_rc(m.Cin, m.Cout);
  KP_CHECK_ARG(rcmax > 0 && (long long)m.N <= (long long)rcmax * DB_MAX_GRID,
               "kp_dense_block: N=%d exceeds kp_dense_block_max_rows", m.N);
  static const int env_rows = getenv("KP_DENSE_ROWS") ? atoi(getenv("KP_DENSE_ROWS")) : 0;
  int Rc = env_rows > 0 ? env_rows : DB_ROWS;
  if ((long long)Rc * DB_MAX_GRID < m.N) Rc = (m.N + DB_MAX_GRID - 1) / DB_MAX_GRID;
  Rc = (Rc + 3) & ~3;
  if (Rc > rcmax) Rc = rcmax;
  c->Rc = Rc;
  c->grid = (m.N + Rc - 1) / Rc;
  c->smem_fwd = db_smem_fwd(m.Cin, m.Cout, Rc);
  c->smem_bwd = db_smem_bwd(m.Cin, m.Cout, Rc);
  c->ws_fwd = 256 + sizeof(float) * (size_t)c->grid * 2 * m.Cout * 3;
  c->ws_bwd = 256 + sizeof(float) * (size_t)c->grid *
                        (6 * (size_t)m.Cout + (size_t)m.Cout * m.Cin + (size_t)m.Cout * m.Cout + 2 * (size_t)m.Cout);
  return 0;
}

}  // namespace kp

extern "C" {

int kp_dense_block_set_mma(int mode) {
  kp::g_db_mma = mode < 0 ? -1 : (mode ? 1 : 0);
  return 0;
}

int kp_dense_block_max_rows(int32_t Cin, int32_t Cout) {
  if (Cin < 4 || Cout < 4 || Cin % 4 || Cout % 4 || Cin > 128 || Cout > 128 || Cin > Cout) return 0;
  return kp::db_max_rc(Cin, Cout) * kp::DB_MAX_GRID;
}

int kp_dense_block_workspace_bytes(const kp_dense_desc* desc, size_t* fwd_bytes, size_t* bwd_bytes) {
  KP_CHECK_ARG(desc && fwd_bytes && bwd_bytes, "kp_dense_block_workspace_bytes: null argument");
  kp::DbCfg c;
  if (kp::db_config(*desc, &c)) return 1;
  *fwd_bytes = c.ws_fwd;
  *bwd_bytes = c.ws_bwd;
  return 0;
}

static int kp_dense_check(const kp_dense_desc& m) {
  KP_CHECK_ARG(m.X && m.W1 && m.b1 && m.g1 && m.be1 && m.W2 && m.b2 && m.g2 && m.be2 && m.Y1 && m.Y2 && m.stats,
               "kp_dense_block: null argument");
  KP_CHECK_ARG(!m.g3 || (m.be3 && m.Z2), "kp_dense_block: BN3 needs be3 and Z2");
  KP_CHECK_ARG((((uintptr_t)m.X | (uintptr_t)m.W1 | (uintptr_t)m.b1 | (uintptr_t)m.g1 | (uintptr_t)m.be1 |
                 (uintptr_t)m.W2 | (uintptr_t)m.b2 | (uintptr_t)m.g2 | (uintptr_t)m.be2 | (uintptr_t)m.g3 |
                 (uintptr_t)m.be3 | (uintptr_t)m.R | (uintptr_t)m.Y1 | (uintptr_t)m.Y2 | (uintptr_t)m.Z2 |
                 (uintptr_t)m.stats) & 15) == 0,
               "kp_dense_block: pointers must be 16-byte aligned");
  return 0;
}

int kp_dense_block_forward(const kp_dense_desc* desc, float* out, void* workspace, size_t workspace_bytes,
                           void* stream) {
  KP_CHECK_ARG(desc && out && workspace, "kp_dense_block_forward: null argument");
  const kp_dense_desc& m = *desc;
  kp::DbCfg c;
  if (kp::db_config(m, &c)) return 1;
  if (kp_dense_check(m)) return 1;
  KP_CHECK_ARG(workspace_bytes >= c.ws_fwd && (((uintptr_t)workspace | (uintptr_t)out) & 15) == 0,
               "kp_dense_block_forward: workspace too small or misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* bar = m.barrier ? (unsigned*)m.barrier : (unsigned*)workspace;
  if (!m.barrier) KP_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
  KP_CUDA(cudaFuncSetAttribute(kp::dense_block_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)c.smem_fwd));
  {
    kp_dense_desc marg = m;
    float* part = (float*)((char*)workspace + 256);
    int rc_rows = c.Rc;
    int mma = kp::db_use_mma(m);
    void* args[] = {&marg, &out, &part, &bar, &rc_rows, &mma};
    KP_LAUNCH_COOP(kp::dense_block_fwd_kernel, c.grid, kp::DB_THREADS, c.smem_fwd, st, args);
  }
  return 0;
}

int kp_dense_block_backward(const kp_dense_desc* desc, const float* dOut, float* dX, float* dW1, float* db1,
                            float* dW2, float* db2, float* dbn, void* workspace, size_t workspace_bytes,
                            void* stream) {
  KP_CHECK_ARG(desc && dOut && dX && dW1 && db1 && dW2 && db2 && dbn && workspace,
               "kp_dense_block_backward: null argument");
  const kp_dense_desc& m = *desc;
  kp::DbCfg c;
  if (kp::db_config(m, &c)) return 1;
  if (kp_dense_check(m)) return 1;
  KP_CHECK_ARG(workspace_bytes >= c.ws_bwd &&
                   (((uintptr_t)workspace | (uintptr_t)dOut | (uintptr_t)dX | (uintptr_t)dW1 | (uintptr_t)db1 |
                     (uintptr_t)dW2 | (uintptr_t)db2 | (uintptr_t)dbn) & 15) == 0,
               "kp_dense_block_backward: workspace too small or pointers misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* bar = m.barrier ? (unsigned*)m.barrier : (unsigned*)workspace;
  if (!m.barrier) KP_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
  KP_CUDA(cudaFuncSetAttribute(kp::dense_block_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)c.smem_bwd));
  {
    kp_dense_desc marg = m;
    float* part = (float*)((char*)workspace + 256);
    int rc_rows = c.Rc;
    int mma = kp::db_use_mma(m);
    void* args[] = {&marg, (void*)&dOut, &dX, &dW1, &db1, &dW2, &db2, &dbn, &part, &bar, &rc_rows, &mma};
    KP_LAUNCH_COOP(kp::dense_block_bwd_kernel, c.grid, kp::DB_THREADS, c.smem_bwd, st, args);
  }
  {
    cudaStream_t lst = st;
    if (m.leaf_stream && m.leaf_stream != stream) {          // leaf gradients: fork (see kp_dense_desc.leaf_stream)
      lst = (cudaStream_t)m.leaf_stream;
      KP_CUDA(kp::fork_stream(st, lst));
    }
    const int total4 = (m.Cout * m.Cin + m.Cout * m.Cout + 2 * m.Cout) >> 2;   // output float4s, 16 lanes each
    KP_LAUNCH(kp::dense_block_wgrad_reduce_kernel, kp::ceil_div((long long)total4 * 16, 256), 256, 0, lst,
              (const float*)((char*)workspace + 256), c.grid, m.Cin, m.Cout, dW1, db1, dW2, db2);
  }
  return 0;
}

}  // extern "C"
