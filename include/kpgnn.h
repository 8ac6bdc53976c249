/* kpgnn.h -- C ABI of libkpgnn_b200.so: the B200 (sm_100a) kernels behind the K-hop aggregation path of KP-GNN.
 *
 * The reference (JiaruiFeng/KP-GNN) is pure Python and has no FFI; the interface these entry points replace is
 * the body of its layer modules and of its extraction function.  Each entry cites the reference lines it
 * replaces.  Conventions:
 *   - every function returns 0 on success, non-zero on failure; kp_last_error() gives the message (thread-local);
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns and sizes every buffer;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and capturable in CUDA graphs
 *     (no allocation, no synchronisation inside), except where a function is documented as synchronising;
 *   - no global mutable state besides the launch counter, so calls on different streams are thread-safe.
 *   - index arrays are int32 on the device; the int64 tensors of the reference layout are converted by
 *     kp_plan_* (aggregation) and produced by kp_extract_export (extraction).
 */
#ifndef KPGNN_B200_H
#define KPGNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KPGNN_ABI_VERSION 17

const char* kp_last_error(void);
int kp_abi_version(void);
/* kernels launched by this library since load (monotonic; bench.py reports the per-step difference) */
uint64_t kp_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Graph plan: (dst,hop)-sorted CSR + (src,hop)-sorted transposed CSR of the hop-labelled edge list.
 * Replaces what PyG's propagate re-derives every layer from `edge_index [2,E]` / `edge_attr [E,K]`
 * (layers/KPGIN.py:100, KPGINplus.py:74, KPGCN.py:85-110, KPGraphSAGE.py:86, gine.py:52).
 *
 * Row r = v*K + h holds the in-edges e=(u->v) with edge_attr[e,h] != 0, in ascending edge id.
 * self_loops != 0 appends, to every row (v,h), the entry (u=v, attr=1) -- KPGCN.py:85-89 -- and fills
 * dinv[v*K+h] = deg^-1/2 with deg counted as in KPGCN.py:11-25 (loop included).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  const int64_t* src;      /* edge_index[0], E entries */
  const int64_t* dst;      /* edge_index[1], E entries */
  const int64_t* attr;     /* edge_attr, element (e,h) at attr[e*attr_stride + h] */
  int64_t attr_stride;
  int32_t N, E, K;
  int32_t self_loops;
} kp_plan_input;

/* bytes of scratch kp_plan_count / kp_plan_fill need (same buffer may be reused for both) */
int kp_plan_workspace_bytes(int32_t N, int32_t E, int32_t K, size_t* bytes);

/* Pass 1.  Writes rowptr/rowptrT (N*K+1 ints each, exclusive scan; nnz = rowptr[N*K] = rowptrT[N*K]),
 * indeg[N] (in-edges per node irrespective of hop masks; KPGraphSAGE aggr="mean"), and
 * stats[0..3] = {nnz, max attr in hop column 0, max attr in hop columns >= 1, number of out-of-range src/dst}. */
int kp_plan_count(const kp_plan_input* in, int32_t* rowptr, int32_t* rowptrT, int32_t* indeg, int32_t* stats,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Pass 2.  Fills col/attr16 (by dst rows), colT (by src rows; holds dst ids) -- nnz entries each -- and, when
 * self_loops, dinv[N*K].  Deterministic: entries of a row are in ascending edge id, loop entry last.
 * `capacity` = entries allocated in col/attr16/colT; writes beyond it are dropped, so a caller that refreshes
 * a plan in place without first reading nnz back (CUDA-graph replay loops) cannot overrun its buffers and
 * detects the overflow later from stats[0] > capacity. */
int kp_plan_fill(const kp_plan_input* in, const int32_t* rowptr, const int32_t* rowptrT, int32_t* col,
                 uint16_t* attr16, int32_t* colT, float* dinv, int32_t capacity, void* workspace,
                 size_t workspace_bytes, void* stream);

/* In-place refreshes with a fixed capacity: rows past `capacity` were not emitted by kp_plan_fill; clamp both row
 * pointer arrays (N*K+1 ints) to it so that consumers see those rows as empty and never index past col/attr16/colT. */
int kp_plan_clamp(int32_t* rowptr, int32_t* rowptrT, int32_t N, int32_t K, int32_t capacity, void* stream);

/* Closed node blocks of a plan: maximal runs of consecutive nodes whose in- and out-neighbours (all hops) stay inside
 * the run -- the graphs of a collated batch (PyG Batch.from_data_list numbers nodes graph by graph), found without
 * the `batch` vector, which the reference's layer signature (KPGINplus.py:61) does not carry.  block_ptr [N+1] receives
 * the block boundaries (num_blocks+1 used), block_stats[0..2] = {num_blocks, max nodes per block, max entries per
 * block}.  The block-resident kernels (kp_agg_desc.block_ptr) partition their work by these ranges. */
int kp_plan_blocks_workspace_bytes(int32_t N, size_t* bytes);
int kp_plan_blocks(const int32_t* rowptr, const int32_t* col, const int32_t* rowptrT, const int32_t* colT, int32_t N,
                   int32_t K, int32_t capacity, int32_t* block_ptr, int32_t* block_stats, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Per-hop masked aggregation with fused epilogue (forward) -- the message/aggregate/update of
 * KPGIN.py:100-105,115-121; KPGINplus.py:74-78,82-88; KPGCN.py:107-126; KPGraphSAGE.py:86-89,100-106;
 * gine.py:52-59; run_simulation.py:73-75,87-93; and GeometricCombine.forward, combine.py:43-58, when fused.
 *
 *   acc[v,h,:] = sum_{j in row(v,h)} wsrc[col_j,h] * ( X[col_j,h,:] + T_h[attr_j,:] )      T_0=T0, T_{h>=1}=Tk
 *   z[v,h,:]   = act( acc * wdst[v,h] * mean_scale[v] ) + P[v,h,:] + (1+eps) * X[v,h,:]
 *   out        = fuse ? sum_h theta[h,:] * z[v,h,:]   ([N,d])   :   z   ([N,k,d] contiguous)
 * wsrc = wdst = dinv when dinv != NULL (else 1); mean_scale = 1/max(indeg,1) when indeg != NULL (else 1);
 * the P term is skipped when P == NULL, the self term when eps == NULL, the tables when T0 == NULL.
 * ---------------------------------------------------------------------------------------------------------- */
enum { KP_ACT_NONE = 0, KP_ACT_GELU = 1, KP_ACT_RELU = 2 };

typedef struct {
  int32_t N, Kplan, k, d;          /* nodes; hop count the plan was built with; hops used (k<=Kplan); width */
  const int32_t* rowptr;           /* N*Kplan+1 */
  const int32_t* col;              /* nnz */
  const uint16_t* attr16;          /* nnz */
  const int32_t* rowptrT;          /* backward only */
  const int32_t* colT;             /* backward only */
  const float* dinv;               /* N*Kplan or NULL */
  const int32_t* indeg;            /* N or NULL */
  const float* X;  int64_t x_node_stride, x_hop_stride;      /* element strides */
  const float* P;  int64_t p_node_stride, p_hop_stride;      /* NULL = no peripheral term */
  const float* T0; const float* Tk; int32_t rows0, rowsk;    /* embedding tables [rows, d]; NULL = none */
  const float* theta;              /* [k,d], required when fuse */
  const float* eps;                /* device scalar or NULL */
  int32_t act, fuse;
  int32_t amax0, amaxk;            /* largest attr16 value present in hop 0 / in hops >= 1 of the plan
                                      (kp_plan_count stats[1], stats[2]); -1 = not supplied.  Backward only: selects
                                      the register-accumulator table-gradient kernel when both are <= 31 */
  /* Backward only: dX row (v,h) lives at dX + v*dx_node_stride + h*dx_hop_stride (elements, multiples of 4; 0 = the
   * contiguous [N,k,d] layout) and, with dx_accumulate, is ADDED to instead of written -- the layer-history gradient
   * buffer of kpgnn_b200/stack.py.  Served by the lean gather kernel only: kp_agg_backward returns 3 (and touches
   * nothing) when another kernel family would have to run, and the caller falls back to a temporary. */
  int64_t dx_node_stride, dx_hop_stride;
  int32_t dx_accumulate;
  /* Backward only: node-range (chunked) backward for batches whose hand-over tensor Gs [N,k,d] exceeds the L2.  Graphs
   * are closed node sets, so the backward of nodes [n0, n1) (whole graphs) is an independent call: N = n1-n0, rowptr /
   * rowptrT advanced by n0*Kplan entries, P / dOut / dX / dP advanced by n0 rows, X and col / colT UNCHANGED (sources
   * are gathered by their batch-wide ids), node_base = n0.  The chunk's Gs then lives in a workspace of the CHUNK's
   * size that every chunk reuses -- it stays in L2 between the three kernels and never goes to HBM.  Table / theta
   * gradients of the chunks are partial sums the caller adds.  Only where kp_agg_backward_chunkable() says so. */
  int32_t node_base;
  /* Backward only, with fuse: when both are non-NULL the reduction of the per-CTA dtheta partials is fused with
   * GeometricCombine's backward (combine.py:51-58): geo_dalphas [d] receives d(loss)/d(alphas) for theta =
   * kp_geometric_theta_forward(geo_alphas); the dtheta argument of kp_agg_backward may then be NULL. */
  const float* geo_alphas;
  float* geo_dalphas;
  /* Backward only, optional: a second CUDA stream for the LEAF gradients (dT0/dTk, dtheta/dalphas, deps), which feed
   * nothing but the optimizer.  They are forked onto it right after the first backward kernel, so they overlap the
   * dX gather and whatever the caller enqueues next on `stream`.  The caller must make its stream wait for
   * leaf_stream before reading those outputs and keep outputs + workspace alive until then.  NULL: one stream. */
  void* leaf_stream;
  /* Optional closed node blocks of the plan (kp_plan_blocks: block_ptr [num_blocks+1], the largest block's node
   * count).  When supplied, unfused calls whose largest block fits in shared memory run block-resident: a CTA stages one
   * (block, hop) slice of X and takes every gather of the block's rows from shared memory -- the long-row regime of
   * the regular-graph workload (run_simulation.py, ~177 entries per node).  NULL: row-streaming kernels only. */
  const int32_t* block_ptr;
  const int32_t* block_stats;          /* device: [0] = current number of blocks (read by the kernel, so that a captured
                                          launch follows in-place plan refreshes); NULL: num_blocks below is exact */
  int32_t num_blocks, max_block_nodes; /* host-side sizing: grid ~ num_blocks * k units, shared memory ~ max_block_nodes */
} kp_agg_desc;

int kp_agg_forward(const kp_agg_desc* desc, float* out, void* stream);
/* Test hook, process-wide; a bit set of kernel-family selectors so every implementation is parity-tested on the
 * same inputs.  bit 0: generic (any width / alignment / stride) kernels instead of the float4 fast path;
 * bit 1: register-prefetch instead of cp.async-ring forward kernel; bit 2: no packed-math ("lean") kernels;
 * bit 4: TMA-staged forward kernel for every eligible call; bit 5: TMA-staged forward kernel for large batches;
 * bit 6: ignore desc.block_ptr (no block-resident kernels).
 * 0 = production default (lean kernels, no TMA staging). */
int kp_agg_set_force_generic(int flag);
/* Test hook, process-wide: max_ctas > 0 caps the grid of every persistent aggregation kernel, lean_threads > 0 (a
 * multiple of 32 in [256,1024]) forces the CTA size of the packed-math kernels -- so a batch of a few thousand nodes
 * runs the same multi-node-per-lane-group loops, software pipeline and partial reductions as the 8 192-graph launch
 * the roofline is quoted on.  (0, 0) = production geometry. */
int kp_agg_set_launch_geometry(int max_ctas, int lean_threads);

/* Backward of kp_agg_forward (autograd of the same reference lines).  Deterministic, no float atomics:
 * transposed-CSR gather for dX, owner-computes partial tables for dT0/dTk, per-CTA partials for dtheta/deps.
 *   dOut        [N,d] if fuse else [N,k,d]
 *   dX          [N,k,d] contiguous (written, not accumulated), may be NULL
 *   dP          [N,k,d] contiguous or NULL (when !fuse, dP == dOut and the caller should alias instead)
 *   dT0,dTk     [rows0,d],[rowsk,d] or NULL;  dtheta [k,d] or NULL;  deps [1] or NULL
 *   workspace   kp_agg_backward_workspace_bytes(desc) bytes
 */
int kp_agg_backward_workspace_bytes(const kp_agg_desc* desc, size_t* bytes);
/* *ok = 1 when this call's backward runs on the kernels that honour kp_agg_desc.node_base (the packed-math gather
 * family); kp_agg_backward refuses a non-zero node_base otherwise. */
int kp_agg_backward_chunkable(const kp_agg_desc* desc, int32_t* ok);
int kp_agg_backward(const kp_agg_desc* desc, const float* dOut, float* dX, float* dP, float* dT0, float* dTk,
                    float* dtheta, float* deps, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-table gather-sum: the peripheral-attribute embedding stage of the reference's backbones
 * (models/GNNs.py:172-179 / :393-400 / :637-644 with layers/feature_encoder.py:37-67), after the caller folded
 * each embedding table through its slice of the encoder's Linear:
 *     out[r,:]    = sum_{s<S} table[slot_off[s] + idx[r,s], :]
 *     dTable[t,:] = sum_{(r,s): slot_off[s]+idx[r,s] == t} dOut[r,:]          (deterministic, no float atomics)
 * idx is [R,S] int64 row-major (table-local indices, the reference's integer peripheral attributes);
 * slots must be ordered so that their tables are non-decreasing; range_slot/range_row partition the slots and
 * the table rows into num_ranges (<= 16) consecutive pieces, each small enough for shared memory (<= 200 KB).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t R, S, d, table_rows;
  const int64_t* idx;
  int32_t slot_off[32];
  int32_t num_ranges;
  int32_t range_slot[17];
  int32_t range_row[17];
} kp_tsum_desc;

int kp_table_sum_forward(const kp_tsum_desc* desc, const float* table, float* out, void* stream);
/* Profiling hook: caps the shared memory the gradient kernel asks for (ncu cannot re-launch CUDA-graph kernel
 * nodes that opted into > 48 KB of dynamic shared memory).  Clamped to [16 KB, 200 KB]; default 200 KB. */
int kp_table_sum_set_smem_cap(size_t bytes);
int kp_table_sum_backward_workspace_bytes(const kp_tsum_desc* desc, size_t* bytes);
int kp_table_sum_backward(const kp_tsum_desc* desc, const float* dOut, float* dTable, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Folding the peripheral-attribute encoders (layers/feature_encoder.py:37-67 as used by models/GNNs.py:172-179 /
 * :393-400: embedding lookups -> concat -> Linear, gated by tanh(pew) / sigmoid(pew)) into ONE lookup table for
 * kp_table_sum_*:  table[row_off[i] + r, :] = g_{gate[i]} * E_i[r, :] W_i^T,  last row (row_off[T]) = sum_g
 * bias_mult[g] * g_g * bias[g], with g = tanh(raw) (gate_act 0) or sigmoid(raw) (gate_act 1).  One kernel forward, one
 * backward (dE_i, the dW_i slices, the bias and raw-gate gradients; fixed-order sums).  W[i] points at the first element
 * of the table's [H_out, H_in] slice of its Linear weight, w_stride[i] = that weight's row length. */
typedef struct {
  int32_t T, H_in, H_out, gate_act;
  const float* E[16];
  const float* W[16];
  int64_t w_stride[16];
  int32_t rows[16];
  int32_t gate[16];
  int32_t row_off[17];
  int32_t pad;
  const float* gate_raw[2];
  const float* bias[2];
  float bias_mult[2];
} kp_fold_desc;
typedef struct {
  float* dE[16];          /* [rows_i, H_in] or NULL */
  float* dW[16];          /* slice pointers with the forward's w_stride, or NULL */
  float* dbias[2];        /* [H_out] or NULL */
  float* dgate_raw[2];    /* [1] or NULL */
} kp_fold_grads;
int kp_fold_forward(const kp_fold_desc* desc, float* table, void* stream);
/* workspace: KP_FOLD_WORKSPACE_BYTES, 16-byte aligned, ZEROED ONCE by the caller and private to one stream (the kernel
 * leaves its arrival counter at zero again). */
#define KP_FOLD_WORKSPACE_BYTES 2048
int kp_fold_backward(const kp_fold_desc* desc, const float* dTable, const kp_fold_grads* grads, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Training-mode BatchNorm1d (+ optional fused ReLU) over [N,C] fp32 rows: the BN / ReLU of the KP-GIN+ MLP
 * (layers/KPGINplus.py:25-30) and the backbone's norm (models/GNNs.py:430, norm_type "Batch").  Biased variance
 * for the normalisation, unbiased for running_var, running = (1-momentum)*running + momentum*batch, as
 * torch.nn.BatchNorm1d.  One kernel each way; N <= kp_bn_max_rows(), C % 4 == 0; callers fall back to the
 * framework's BatchNorm otherwise (eval mode, larger N). */
int kp_bn_max_rows(void);
int kp_bn_forward(const float* x, int32_t N, int32_t C, const float* gamma, const float* beta, float eps,
                  float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked, int32_t relu,
                  float* y, float* save_mean, float* save_invstd, void* stream);
int kp_bn_backward(const float* x, const float* dy, int32_t N, int32_t C, const float* gamma, const float* beta,
                   const float* save_mean, const float* save_invstd, int32_t relu, float* dx, float* dgamma,
                   float* dbeta, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused dense block of a KP-GIN+ layer, training mode: everything between the aggregation and the next layer,
 *     y1 = X W1^T + b1;  z1 = relu(BN1(y1));  y2 = z1 W2^T + b2;  z2 = relu(BN2(y2))     (layers/KPGINplus.py:25-30,78)
 *     out = BN3(z2) + R                                                                    (models/GNNs.py:430-438)
 * as ONE persistent kernel per direction (BN3 and R optional).  Each CTA owns a slab of rows and both weight
 * matrices in shared memory; the three batch statistics are merged across CTAs (Chan's parallel variance, fixed
 * order -> bit-reproducible) behind a grid-wide barrier.  fp32 FMA throughout (no TF32).  The backward is the
 * autograd of the same lines: dX, dW1, db1, dW2, db2 and the BN affine gradients, per-CTA partial weight
 * gradients summed in a fixed order.  Replaces 2 GEMMs + 3 BatchNorm kernels + bias/residual elementwise kernels
 * forward and 4 GEMMs + 3 BatchNorm backward kernels + 4 column sums backward.
 * N <= kp_dense_block_max_rows(Cin, Cout); Cin, Cout multiples of 4, <= 128; running statistics required.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t N, Cin, Cout;
  const float* X;                                   /* [N,Cin] */
  const float *W1, *b1, *g1, *be1;                  /* Linear1 [Cout,Cin],[Cout]; BN1 weight/bias */
  const float *W2, *b2, *g2, *be2;                  /* Linear2 [Cout,Cout],[Cout]; BN2 weight/bias */
  const float *g3, *be3;                            /* outer BatchNorm weight/bias, NULL = no BN3 */
  const float* R;                                   /* residual [N,Cout] added to the result, or NULL */
  float eps1, eps2, eps3, mom1, mom2, mom3;
  float *rm1, *rv1, *rm2, *rv2, *rm3, *rv3;         /* running mean / var (updated in place) or NULL */
  int64_t *nbt1, *nbt2, *nbt3;                      /* num_batches_tracked (incremented) or NULL */
  float *Y1, *Y2, *Z2;                              /* saved for backward: y1, y2 [N,Cout]; z2 (only with BN3) */
  float* stats;                                     /* [6,Cout]: mean1, invstd1, mean2, invstd2, mean3, invstd3 */
  /* Row strides in elements (0 = Cout): `out` and R may be column blocks of a wider matrix -- the layer-history
   * buffer [N, L+1, H] of kpgnn_b200/stack.py -- and so may dOut in the backward.  dR (backward only, may be NULL):
   * the residual's gradient is ACCUMULATED there, dR[row] += dOut[row], instead of being returned to the caller. */
  int64_t out_stride, r_stride, dout_stride, dr_stride;
  float* dR;
  /* Optional persistent grid-barrier state: 256 bytes of device memory, zeroed ONCE by the caller and used by one
   * stream at a time; every launch leaves it zero again, so no memset node precedes the kernel.  NULL: the first
   * 256 bytes of the workspace are used and cleared by a cudaMemsetAsync before each launch. */
  uint32_t* barrier;
  /* Backward only, optional: stream for the final weight-gradient reduction (dW1, db1, dW2, db2), same contract as
   * kp_agg_desc.leaf_stream.  dX and the BatchNorm affine gradients are always produced on `stream`. */
  void* leaf_stream;
  /* Optional: the batch's row count in DEVICE memory.  N above is then the CAPACITY of every [N,*] buffer (it sizes the
   * grid and the workspace) and *n_dev <= N the number of rows that exist: rows in [*n_dev, N) are padding -- excluded
   * from the batch statistics, the running statistics and every gradient sum; `out` and dX are written as zeros there.
   * One captured launch (CUDA graph) then serves batches of different sizes (train_ZINC.py:29-47: a new batch, with a
   * new node count, every step).  NULL: N rows exist. */
  const int32_t* n_dev;
} kp_dense_desc;

/* Test / A-B hook, process-wide: 1 = GEMM phases of the dense block on the tensor cores (mma.sync TF32 with error
 * compensation, "3xTF32"; channel counts must be multiples of 8, otherwise the call keeps the fp32-FMA tiles), 0 = fp32
 * FMA tiles everywhere, -1 = default (off: measured slower than the FMA tiles, profiles/r2_dense_mma.txt; environment
 * KP_DENSE_MMA=1 turns it on). */
int kp_dense_block_set_mma(int mode);
int kp_dense_block_max_rows(int32_t Cin, int32_t Cout);
int kp_dense_block_workspace_bytes(const kp_dense_desc* desc, size_t* fwd_bytes, size_t* bwd_bytes);
int kp_dense_block_forward(const kp_dense_desc* desc, float* out, void* workspace, size_t workspace_bytes,
                           void* stream);
/* dOut [N,Cout]; dX [N,Cin]; dbn = [6,Cout]: dg1, dbe1, dg2, dbe2, dg3, dbe3 (last two untouched without BN3).
 * The residual's gradient is dOut itself (the caller aliases it). */
int kp_dense_block_backward(const kp_dense_desc* desc, const float* dOut, float* dX, float* dW1, float* db1,
                            float* dW2, float* db2, float* dbn, void* workspace, size_t workspace_bytes,
                            void* stream);

/* Gradient of the peripheral term shared by the L layers of a KP-GIN+ stack (models/GNNs.py:393-400,429): layer l
 * adds P[:, :k_l] after its activation and combines hops with theta_l (combine.py:43-47), so
 *     dP[v,h,:] = sum over layers l with k_l > h of theta_l[h,:] * dAgg_l[v,:]        (theta_l == NULL: weight 1)
 * where dAgg_l [N,d] is the gradient w.r.t. layer l's aggregation output.  One kernel instead of L zero-padded
 * slices and L-1 accumulations of [N,K,d] tensors. */
typedef struct {
  int32_t N, K, d, L;
  const float* dagg[32];
  const float* theta[32];           /* [k_l, d] or NULL */
  int32_t k[32];
} kp_pgrad_desc;
int kp_peripheral_grad(const kp_pgrad_desc* desc, float* dP, void* stream);

/* Adam over a list of tensors in ONE launch (torch.optim.Adam semantics, no amsgrad / weight decay; the reference
 * trains with torch.optim.Adam(lr), train_ZINC.py:244).  `tensors_dev` is a device array describing every tensor,
 * `chunks_dev` a device array of (tensor index, element offset) pairs, one per 1024-element chunk; `state_dev` two
 * device ints {steps taken, scratch}, zero-initialised by the caller, advanced by the kernel (graph-replay safe). */
typedef struct {
  float* p;
  const float* g;
  float* m;
  float* v;
  int32_t n, pad;
} kp_adam_tensor;
int kp_adam_step(const kp_adam_tensor* tensors_dev, const int32_t* chunks_dev, int32_t nchunks, float lr, double beta1,
                 double beta2, float eps, int32_t* state_dev, void* stream);   /* betas in double: 1-beta is formed exactly */

/* ------------------------------------------------------------------------------------------------------------
 * AttentionCombine, layers/combine.py:8-27: a bidirectional single-layer LSTM (input size d, hidden size K,
 * batch_first) over the hop axis of x [N,K,d], its [N,K,2K] output summed over the last axis, a softmax over hops,
 * and the weighted sum of x over hops -> out [N,d].  One kernel forward (a warp per node; x is read once, nothing
 * of size [N,K,*] is written), one recompute-and-backpropagate kernel backward plus fixed-order reductions of the
 * parameter gradients (no float atomics).  Weights use torch.nn.LSTM's layout: w_ih [4K,d], w_hh [4K,K], b_ih [4K],
 * b_hh [4K], gate order i,f,g,o; index 0 = forward direction (`*_l0`), 1 = reverse (`*_l0_reverse`).
 * 1 <= K <= 16, 1 <= d <= 128.  x rows have a dense last dimension; node / hop strides in elements.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t N, K, d, pad;
  const float* x;
  int64_t x_node_stride, x_hop_stride;
  const float* w_ih[2];
  const float* w_hh[2];
  const float* b_ih[2];
  const float* b_hh[2];
} kp_attn_desc;
/* out [N,d]; weights [N,K] (the softmax weights, optional, NULL to skip) */
int kp_attn_combine_forward(const kp_attn_desc* desc, float* out, float* weights, void* stream);
int kp_attn_combine_backward_workspace_bytes(const kp_attn_desc* desc, size_t* bytes);
/* dOut [N,d] contiguous; dX [N,K,d] contiguous (written); dw_ih_* [4K,d], dw_hh_* [4K,K], db_* [4K] (the gradient of
 * b_ih and of b_hh alike), _f = forward direction, _r = reverse. */
int kp_attn_combine_backward(const kp_attn_desc* desc, const float* dOut, float* dX, float* dw_ih_f, float* dw_ih_r,
                             float* dw_hh_f, float* dw_hh_r, float* db_f, float* db_r, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Compact batch wire format -> the reference's int64 wire tensors at a fixed capacity (csrc/wire.cu).  Replaces the
 * host-side collation + int64 upload of the reference's loader (train_ZINC.py:29-47, PyG Batch.from_data_list + `.to`):
 * int32 node ids, 1/2-byte attributes and per-graph node offsets cross PCIe; one kernel widens them into the static
 * tensors the layers read and pads the tail (nodes >= N: x = 0, peripheral attrs = 0, batch = g; edges >= E: src = dst
 * = 0 and every hop attr = 0, i.e. masked in every hop), so that the whole training step can be captured once and
 * replayed for batches of different sizes.  hdr (device) = {N, E}; the node count is also written to *o_n.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_cap, e_cap, g, K, met, hp1;      /* capacities; graphs per batch; hops; max_edge_type; max_hop_num + 1 */
  int32_t x_bytes, attr_bytes, p_bytes, pad; /* element widths of x (1|2|4), edge attrs (1|2), peripheral attrs (1|2) */
  const int32_t* hdr;                        /* device [2]: N, E of this batch */
  const int32_t* gptr;                       /* [g+1] node offsets of the graphs */
  const void* x;                             /* [N] node types, or NULL */
  const int32_t* src; const int32_t* dst;    /* [E] K-hop edge list (global node ids) */
  const void* attr;                          /* [E,K] hop attributes */
  const void* pea; const void* pca;          /* [N,K,met,2], [N,K,hp1] or NULL */
  int64_t* o_x; int64_t* o_batch;            /* [n_cap] or NULL */
  int64_t* o_ei;                             /* [2,e_cap] */
  int64_t* o_ea;                             /* [e_cap,K] */
  int64_t* o_pea; int64_t* o_pca;            /* [n_cap,K,met,2], [n_cap,K,hp1] or NULL */
  int32_t* o_n;                              /* device scalar or NULL */
} kp_wire_desc;
int kp_wire_unpack(const kp_wire_desc* desc, void* stream);

/* ---- Data-parallel gradient exchange over NVLink peer memory --------------------------------------------
 * The reference trains on one GPU (train_ZINC.py:29-47: loss.backward(); optimizer.step()); split over ranks, every rank
 * must apply the mean of the ranks' gradients.  Each rank owns one peer-visible block
 *     [ KP_PEER_FLAG_BYTES of flags | gradient: n floats | result: n floats ]   (vectors padded to 256 bytes:
 *     KP_PEER_VECTOR_BYTES(n); kp_peer_block_bytes(n) bytes in all, zero-initialised)
 * allocated by kp_peer_alloc (cudaMalloc), exported as a CUDA-IPC handle, and imported once by every other rank.
 * kp_peer_allreduce_mean (n % 4 == 0): result[i] = (g_0[i] + ... + g_{world-1}[i]) / world in EVERY rank's block, summed
 * in rank order by the rank that owns element i (two-shot: each rank reduces 1/world of the vector from the peers'
 * blocks and stores it into all of them; bit-identical everywhere); flag exchanges before (gradients complete) and
 * after (results delivered, blocks may be overwritten).  Every rank must call it the same number of times.  A wait that exceeds 20 s sets *error (1: waiting for
 * gradients, 2: waiting for readers) and ends the kernel; the caller checks it at its next synchronisation point. */
#define KP_PEER_MAX 8
#define KP_PEER_CTAS 64
#define KP_PEER_HANDLE_BYTES 64
#define KP_PEER_FLAG_BYTES (2 * KP_PEER_CTAS * KP_PEER_MAX * 4)
#define KP_PEER_VECTOR_BYTES(n) ((((size_t)(n)) * 4 + 255) & ~(size_t)255)
typedef struct kp_peer_desc {
  int32_t world, rank;
  int64_t n;                    /* gradient length in floats */
  char* block[KP_PEER_MAX];     /* block[r]: rank r's block as mapped in THIS process (own entry: the local allocation) */
  float* out;                   /* unused (the result lives in the block); keep NULL */
  int32_t* epoch;               /* [KP_PEER_CTAS] local, zero-initialised, owned by the kernel */
  int32_t* error;               /* local device int */
  float scale;                  /* 1 / world */
} kp_peer_desc;
size_t kp_peer_block_bytes(int64_t n);
int kp_peer_alloc(size_t bytes, void** ptr);
int kp_peer_free(void* ptr);
int kp_peer_export(const void* ptr, unsigned char handle[KP_PEER_HANDLE_BYTES]);
int kp_peer_import(const unsigned char handle[KP_PEER_HANDLE_BYTES], void** ptr);
int kp_peer_release(void* ptr);
int kp_peer_allreduce_mean(const kp_peer_desc* desc, void* stream);

/* Graph-regression head as one kernel each way: h = relu(rep); pooled[g] = sum (mean != 0: mean) of h over the rows of
 * graph g; score[g] = <w, pooled[g]> + b; loss = mean_g |score - y| (loss_kind 0, train_ZINC.py:42) or mean_g (score - y)^2
 * (loss_kind 1) -- models/GNNs.py:276-277 (ReLU of the output projection, dropout 0), models/GraphRegression.py:26 and its
 * Linear(H,1) regressor.  batch is int64, non-decreasing, graph ids in [0,G) (rows with id >= G, the padding of a capacity
 * batch, contribute nothing and receive zero gradient); n_dev as in kp_dense_desc.  Fixed-order sums, no float atomics.
 * workspace: kp_head_workspace_bytes(G), 16-byte aligned, ZEROED ONCE by the caller and private to one stream. */
typedef struct kp_head_desc {
  int32_t N, H, G, mean, loss_kind, pad;
  const float* rep; int64_t rep_stride;      /* [N,H] pre-activation of the output projection */
  int64_t rep_stride_out;                    /* backward: row stride of drep */
  const int64_t* batch;
  const int32_t* n_dev;                      /* live row count (device) or NULL */
  const float* w; const float* b;            /* regressor weight [H] and bias [1] */
  const float* y;                            /* targets [G] */
} kp_head_desc;
size_t kp_head_workspace_bytes(int32_t G);
int kp_head_forward(const kp_head_desc* desc, float* pooled, float* score, float* loss, void* workspace,
                    size_t workspace_bytes, void* stream);
int kp_head_backward(const kp_head_desc* desc, const float* pooled, const float* score, const float* dloss, float* drep,
                     float* dw, float* db, void* workspace, size_t workspace_bytes, void* stream);

/* Graph readout over a sorted segment vector: PyG global_add_pool / global_mean_pool on `data.batch`
 * (models/GraphRegression.py:26, models/GraphClassification.py:30).  out[g,:] = sum (mean != 0: mean) of the rows i of
 * x [N,C] (row stride x_stride elements) with seg[i] == g; seg is int64, non-decreasing, values in [0,G).  Rows are
 * added in ascending order by one thread per column: bit-reproducible, no float atomics. */
int kp_segment_sum(const float* x, int64_t x_stride, const int64_t* seg, int32_t N, int32_t C, int32_t G, int32_t mean,
                   float* out, void* stream);

/* GeometricCombine weights, layers/combine.py:51-58: theta[h,c] = softmax over h of a_c (1-a_c)^h with
 * a = sigmoid(alphas); theta is [K,d].  Backward returns d(loss)/d(alphas) from d(loss)/d(theta). */
int kp_geometric_theta_forward(const float* alphas, int32_t K, int32_t d, float* theta, void* stream);
int kp_geometric_theta_backward(const float* alphas, const float* theta, const float* dtheta, int32_t K, int32_t d,
                                float* dalphas, void* stream);
/* The combine weights of up to 32 layers in one launch (layer l: alphas[l] [d] -> theta[l] [k[l], d]). */
typedef struct {
  int32_t L, d;
  const float* alphas[32];
  float* theta[32];
  int32_t k[32];
} kp_theta_batch;
int kp_geometric_theta_forward_batched(const kp_theta_batch* batch, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Batched K-hop neighbourhood + peripheral-subgraph extraction.  Replaces, for a whole batch of graphs,
 * data_utils.py:20-107 (extract_multi_hop_neighbors), :110-125 (adj_K_order), :128-162 (get_peripheral_attr),
 * :165-221 (extract_peripheral_attr_v2) and :224-241 (nx_compute_shortest_path_length).
 *
 * Input: the batch's ORIGINAL edges as a CSR by source over global node ids, duplicates merged
 * (emult = multiplicity, etype = summed edge-type value, as the reference's COO->dense conversions do,
 * data_utils.py:52-53); graphs are the contiguous node ranges gptr[g]..gptr[g+1]; pair_off[g] = sum of n^2 over
 * the graphs before g.  The dense hop tensor W (uint16, K * total_pairs entries) is the only large scratch:
 * W[K*pair_off[g] + (k*n + s)*n + v] = min(walk count, cap), masked to dist(s,v)==k+1 for kernel spd.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t G, N, K;
  int32_t n_max;                  /* largest graph in the batch (<= 65535) */
  const int32_t* gptr;            /* [G+1] */
  const int32_t* node_graph;      /* [N] */
  const int64_t* pair_off;        /* [G+1] */
  const int32_t* erow;            /* [N+1] */
  const int32_t* ecol;            /* [E1] global destination ids, ascending inside a row */
  const int32_t* emult;           /* [E1] */
  const int32_t* etype;           /* [E1] */
  int32_t kernel;                 /* 0 = spd, 1 = gd */
  int32_t cap;                    /* saturation of the walk counts: max(max_edge_attr_num, 1), <= 65535 */
  int32_t max_edge_attr_num, max_hop_num, max_edge_type, max_edge_count, max_distance_count;
  int32_t max_type_value;         /* largest etype in the batch (sizes the histogram) */
} kp_extract_input;

int kp_extract_workspace_bytes(const kp_extract_input* in, int64_t total_pairs, size_t* hop_bytes,
                               size_t* scratch_bytes);
/* Fills W and eptr[N+1] = exclusive scan of the K-hop out-degree of every node; E_K = eptr[N]. */
int kp_extract_hops(const kp_extract_input* in, uint16_t* W, int32_t* eptr, void* scratch, size_t scratch_bytes,
                    void* stream);
/* edge_index [2,E_K] (row 0 = src, row 1 = dst) and edge_attr [E_K,K], int64, in the reference's order. */
int kp_extract_emit(const kp_extract_input* in, const uint16_t* W, const int32_t* eptr, int64_t* edge_index,
                    int64_t* edge_attr, int64_t EK, void* stream);
/* peripheral_edge_attr [N,K,max_edge_type,2] and peripheral_configuration_attr [N,K,max_hop_num+1], int64. */
int kp_extract_peripheral(const kp_extract_input* in, const uint16_t* W, int64_t* peripheral_edge_attr,
                          int64_t* peripheral_configuration_attr, void* scratch, size_t scratch_bytes,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KPGNN_B200_H */
