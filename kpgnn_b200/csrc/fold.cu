// fold.cu -- folding the peripheral-attribute encoders into one lookup table, forward and backward, one kernel each.
//
// The reference embeds every integer peripheral attribute, concatenates the embeddings and applies a Linear
// (layers/feature_encoder.py:37-67, called from models/GNNs.py:172-179 / :393-400).  Lookup -> concat -> Linear is
// linear in the table rows,  cat_i(E_i[x_i]) W^T + b = sum_i (E_i W_i^T)[x_i] + b,  so the whole stage is a gather-sum over
// the FOLDED tables  M_i = gate * E_i W_i^T  (kp_table_sum_*).  The folding itself is nine 51 x 104 x 104 products; as
// PyTorch ops it was ~14 launches forward and ~40 backward (78 us of 2-us kernels at the tail of the training step,
// profiles/r1zzz_step_kineto.txt 1102-1180 us).  Here: one CTA per table, forward and backward.
//   forward   M[row_off[i] + r, o] = g_i * sum_c E_i[r,c] W_i[o,c];   last row = sum_g mult_g * g_g * bias_g
//   backward  dE_i = g_i dM_i W_i;  dW_i = g_i dM_i^T E_i;  dbias_g = mult_g g_g dM_last;
//             d(raw gate_g) = act'(raw_g) * ( sum_{i in g} <W_i, dM_i^T E_i> + mult_g <dM_last, bias_g> )
// with g = tanh(raw) (GNNPlus, GNNs.py:396) or sigmoid(raw) (GNN / GNNPrime, GNNs.py:175).  Fixed-order sums: the gate
// gradient is finished by the last CTA to arrive, which adds the per-table partials in table order.
#include <math.h>

#include "common.cuh"

namespace kp {

constexpr int FOLD_THREADS = 256;
constexpr int FOLD_SLICES = 13;    // CTAs per table: row slices (forward, dE) / output-channel slices (dW)

__device__ __forceinline__ float fold_gate(float raw, int act) { return act == 0 ? tanhf(raw) : 1.f / (1.f + expf(-raw)); }
__device__ __forceinline__ float fold_gate_grad(float raw, int act) {
  if (act == 0) {
    const float t = tanhf(raw);
    return 1.f - t * t;
  }
  const float s = 1.f / (1.f + expf(-raw));
  return s * (1.f - s);
}

// shared: E [rows][Hi] | W [Ho][Hi+1]
__global__ void __launch_bounds__(FOLD_THREADS) fold_fwd_kernel(const kp_fold_desc f, float* __restrict__ table) {
  extern __shared__ __align__(16) float sm[];
  const int Hi = f.H_in, Ho = f.H_out;
  const int i = blockIdx.x / FOLD_SLICES, sl = blockIdx.x - i * FOLD_SLICES;
  if (i == f.T) {                                         // the bias row
    if (sl) return;
    const float g0 = fold_gate(__ldg(f.gate_raw[0]), f.gate_act), g1 = fold_gate(__ldg(f.gate_raw[1]), f.gate_act);
    for (int o = threadIdx.x; o < Ho; o += FOLD_THREADS)
      table[(size_t)f.row_off[f.T] * Ho + o] =
          f.bias_mult[0] * g0 * __ldg(f.bias[0] + o) + f.bias_mult[1] * g1 * __ldg(f.bias[1] + o);
    return;
  }
  const int rows = f.rows[i];
  const int rlo = (int)((long long)rows * sl / FOLD_SLICES), rhi = (int)((long long)rows * (sl + 1) / FOLD_SLICES);
  if (rlo >= rhi) return;
  float* Es = sm;
  float* Ws = sm + rows * Hi;
  const int ws = Hi + 1;
  for (int t = rlo * Hi + threadIdx.x; t < rhi * Hi; t += FOLD_THREADS) Es[t] = __ldg(f.E[i] + t);
  for (int o = threadIdx.x >> 5; o < Ho; o += FOLD_THREADS / 32) {
    const float* wrow = f.W[i] + (size_t)o * f.w_stride[i];
    for (int c = threadIdx.x & 31; c < Hi; c += 32) Ws[o * ws + c] = __ldg(wrow + c);
  }
  __syncthreads();
  const float g = fold_gate(__ldg(f.gate_raw[f.gate[i]]), f.gate_act);
  float* out = table + (size_t)f.row_off[i] * Ho;
  const int groups = FOLD_THREADS / Ho > 0 ? FOLD_THREADS / Ho : 1;
  const int og = threadIdx.x / Ho, o = threadIdx.x - og * Ho;
  if (og < groups)
    for (int r0 = rlo + og * 4; r0 < rhi; r0 += groups * 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float* w = Ws + o * ws;
      for (int c = 0; c < Hi; ++c) {
        const float wv = w[c];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (r0 + j < rhi) acc[j] = fmaf(Es[(r0 + j) * Hi + c], wv, acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r0 + j < rhi) out[(size_t)(r0 + j) * Ho + o] = g * acc[j];
    }
}

// shared: E [rows][Hi] | W [Ho][Hi] | dM [rows][Ho] | red [FOLD_THREADS]
__global__ void __launch_bounds__(FOLD_THREADS)
fold_bwd_kernel(const kp_fold_desc f, const float* __restrict__ dTable, kp_fold_grads out, float* __restrict__ part,
                unsigned* __restrict__ counter) {
  extern __shared__ __align__(16) float sm[];
  const int Hi = f.H_in, Ho = f.H_out;
  // grid: FOLD_SLICES CTAs per table (slice sl: a row range of dE and an output-channel range of dW), then the bias CTA
  const int i = blockIdx.x / FOLD_SLICES, sl = blockIdx.x - i * FOLD_SLICES;
  // The gate gradients are sums of ~rows*H*H signed products that cancel to a few per cent of their running partial
  // sums: they are accumulated in double (a few dozen DFMA per thread), so that neither the slicing nor the order of
  // the partials shows in the result.
  __shared__ double red[FOLD_THREADS];
  __shared__ bool last;
  double partial = 0.0;
  double* part_gate = reinterpret_cast<double*>(part);      // [T*FOLD_SLICES] table shares, then [2] bias shares
  if (i == f.T) {                                         // bias row: dbias_g, and its share of the gate gradients
    const float* dMb = dTable + (size_t)f.row_off[f.T] * Ho;
    for (int gsel = 0; gsel < 2; ++gsel) {
      const float g = fold_gate(__ldg(f.gate_raw[gsel]), f.gate_act);
      double s = 0.0;
      for (int o = threadIdx.x; o < Ho; o += FOLD_THREADS) {
        const float dm = __ldg(dMb + o);
        if (out.dbias[gsel]) out.dbias[gsel][o] = f.bias_mult[gsel] * g * dm;
        s = fma((double)dm, (double)__ldg(f.bias[gsel] + o), s);
      }
      red[threadIdx.x] = s;
      __syncthreads();
      for (int w = FOLD_THREADS / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
      }
      if (threadIdx.x == 0) part_gate[f.T * FOLD_SLICES + gsel] = f.bias_mult[gsel] * red[0];
      __syncthreads();
    }
  } else {
    const int rows = f.rows[i];
    const int rlo = (int)((long long)rows * sl / FOLD_SLICES), rhi = (int)((long long)rows * (sl + 1) / FOLD_SLICES);
    const int olo = (int)((long long)Ho * sl / FOLD_SLICES), ohi = (int)((long long)Ho * (sl + 1) / FOLD_SLICES);
    float* Es = sm;
    float* Ws = Es + rows * Hi;
    float* Ms = Ws + Ho * Hi;
    const float* dM = dTable + (size_t)f.row_off[i] * Ho;
    for (int t = threadIdx.x; t < rows * Hi; t += FOLD_THREADS) Es[t] = __ldg(f.E[i] + t);
    for (int o = threadIdx.x >> 5; o < Ho; o += FOLD_THREADS / 32) {          // a warp per weight row: no index division
      const float* wrow = f.W[i] + (size_t)o * f.w_stride[i];
      for (int c = threadIdx.x & 31; c < Hi; c += 32) Ws[o * Hi + c] = __ldg(wrow + c);
    }
    for (int t = threadIdx.x; t < rows * Ho; t += FOLD_THREADS) Ms[t] = __ldg(dM + t);
    __syncthreads();
    const float g = fold_gate(__ldg(f.gate_raw[f.gate[i]]), f.gate_act);
    // (1) dE[r][c] = g * sum_o dM[r][o] W[o][c]: thread owns column c for a group of rows
    {
      const int groups = FOLD_THREADS / Hi > 0 ? FOLD_THREADS / Hi : 1;
      const int rg = threadIdx.x / Hi, c = threadIdx.x - rg * Hi;
      if (rg < groups && out.dE[i])
        for (int r0 = rlo + rg * 4; r0 < rhi; r0 += groups * 4) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          for (int o = 0; o < Ho; ++o) {
            const float wv = Ws[o * Hi + c];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (r0 + j < rhi) acc[j] = fmaf(Ms[(r0 + j) * Ho + o], wv, acc[j]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (r0 + j < rhi) out.dE[i][(size_t)(r0 + j) * Hi + c] = g * acc[j];
        }
    }
    // (2) dW[o][c] = g * sum_r dM[r][o] E[r][c]; the gate gradient is <W, dM^T E> (before the gate factor)
    {
      const int groups = FOLD_THREADS / Hi > 0 ? FOLD_THREADS / Hi : 1;
      const int og = threadIdx.x / Hi, c = threadIdx.x - og * Hi;
      if (og < groups)
        for (int o0 = olo + og * 4; o0 < ohi; o0 += groups * 4) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          for (int r = 0; r < rows; ++r) {
            const float ev = Es[r * Hi + c];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (o0 + j < ohi) acc[j] = fmaf(Ms[r * Ho + o0 + j], ev, acc[j]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (o0 + j < ohi) {
              partial = fma((double)Ws[(o0 + j) * Hi + c], (double)acc[j], partial);
              if (out.dW[i]) out.dW[i][(size_t)(o0 + j) * f.w_stride[i] + c] = g * acc[j];
            }
        }
    }
    red[threadIdx.x] = partial;
    __syncthreads();
    for (int w = FOLD_THREADS / 2; w > 0; w >>= 1) {
      if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) part_gate[i * FOLD_SLICES + sl] = red[0];
  }
  // ---- the last CTA to arrive finishes the two gate gradients, adding the partials in table order
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x < 2) {
    __threadfence();
    const int gsel = threadIdx.x;
    double s = 0.0;
    for (int t = 0; t < f.T; ++t)
      if (f.gate[t] == gsel)
        for (int q = 0; q < FOLD_SLICES; ++q) s += __ldcg(part_gate + t * FOLD_SLICES + q);
    s += __ldcg(part_gate + f.T * FOLD_SLICES + gsel);
    if (out.dgate_raw[gsel]) out.dgate_raw[gsel][0] = (float)s * fold_gate_grad(__ldg(f.gate_raw[gsel]), f.gate_act);
    if (gsel == 0) *counter = 0u;                            // self-resetting: no memset before the next launch
  }
}

static int fold_check(const kp_fold_desc& f) {
  KP_CHECK_ARG(f.T >= 1 && f.T <= 16 && f.H_in >= 1 && f.H_out >= 1 && f.H_in <= FOLD_THREADS && f.H_out <= FOLD_THREADS,
               "kp_fold: need 1 <= T <= 16 tables and widths <= %d", FOLD_THREADS);
  KP_CHECK_ARG(f.gate_raw[0] && f.gate_raw[1] && f.bias[0] && f.bias[1] && (f.gate_act == 0 || f.gate_act == 1),
               "kp_fold: gates / biases missing");
  for (int i = 0; i < f.T; ++i)
    KP_CHECK_ARG(f.E[i] && f.W[i] && f.rows[i] >= 1 && (f.gate[i] == 0 || f.gate[i] == 1) && f.w_stride[i] >= f.H_in,
                 "kp_fold: table %d incomplete", i);
  return 0;
}

}  // namespace kp

extern "C" {

int kp_fold_forward(const kp_fold_desc* desc, float* table, void* stream) {
  KP_CHECK_ARG(desc && table, "kp_fold_forward: null argument");
  const kp_fold_desc& f = *desc;
  if (kp::fold_check(f)) return 1;
  int maxrows = 0;
  for (int i = 0; i < f.T; ++i) maxrows = f.rows[i] > maxrows ? f.rows[i] : maxrows;
  const size_t smem = sizeof(float) * ((size_t)maxrows * f.H_in + (size_t)f.H_out * (f.H_in + 1));
  KP_CHECK_ARG(smem <= 200 * 1024, "kp_fold_forward: a table needs %zu bytes of shared memory", smem);
  if (smem > 32 * 1024)
    KP_CUDA(cudaFuncSetAttribute(kp::fold_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KP_LAUNCH(kp::fold_fwd_kernel, f.T * kp::FOLD_SLICES + 1, kp::FOLD_THREADS, smem, stream, f, table);
  return 0;
}

int kp_fold_backward(const kp_fold_desc* desc, const float* dTable, const kp_fold_grads* grads, void* workspace,
                     size_t workspace_bytes, void* stream) {
  KP_CHECK_ARG(desc && dTable && grads && workspace, "kp_fold_backward: null argument");
  const kp_fold_desc& f = *desc;
  if (kp::fold_check(f)) return 1;
  KP_CHECK_ARG(workspace_bytes >= KP_FOLD_WORKSPACE_BYTES && (((uintptr_t)workspace) & 15) == 0,
               "kp_fold_backward: workspace needs KP_FOLD_WORKSPACE_BYTES bytes");
  int maxrows = 0;
  for (int i = 0; i < f.T; ++i) maxrows = f.rows[i] > maxrows ? f.rows[i] : maxrows;
  const size_t smem = sizeof(float) * ((size_t)maxrows * f.H_in + (size_t)f.H_out * f.H_in + (size_t)maxrows * f.H_out);
  KP_CHECK_ARG(smem <= 200 * 1024, "kp_fold_backward: a table needs %zu bytes of shared memory", smem);
  if (smem > 32 * 1024)
    KP_CUDA(cudaFuncSetAttribute(kp::fold_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // workspace: the arrival counter (zeroed ONCE by the caller; the kernel leaves it zero), then at byte 256 the DOUBLE
  // partials of the gate gradients (T * FOLD_SLICES table shares + 2 bias shares)
  unsigned* counter = (unsigned*)workspace;
  float* part = (float*)((char*)workspace + 256);
  static_assert(256 + 8 * (16 * kp::FOLD_SLICES + 2) <= KP_FOLD_WORKSPACE_BYTES, "fold workspace");
  KP_LAUNCH(kp::fold_bwd_kernel, f.T * kp::FOLD_SLICES + 1, kp::FOLD_THREADS, smem, stream, f, dTable, *grads, part, counter);
  return 0;
}

}  // extern "C"
