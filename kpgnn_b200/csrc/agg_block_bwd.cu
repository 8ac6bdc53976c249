// agg_block_bwd.cu -- the whole backward of the fused K-hop aggregation as ONE block-resident kernel (sm_100a).
//
// Why: the three-kernel backward (B1 by destination -> Gs [N,k,d] in HBM -> B2 by source, B3 table gradients) moves
// 7 streams of N*k*d floats where the algorithm needs 4 (X in, dX / dP out, P in for dtheta): the hand-over tensor Gs
// is written once and read back twice -- 1.10 ms = 37 % of the roofline at 8 192 molecules (profiles/r1 numbers).
// Molecule batches are made of CLOSED node blocks (the graphs: kp_plan_blocks), and a block's share of Gs is tiny
// (23 nodes x 4 hops x 104 floats = 38 KB), so it never has to leave the SM:
//   unit = (block, group of HG hops); a 512-thread CTA walks its units (round-robin, persistent):
//   phase 1 (B1)  a warp per destination node: recompute the pre-activation (gathers + table rows), dy = theta_h dOut
//                 (or dOut[v,h]), dP -> global, Gs = dy act'(pre) -> SHARED memory, dtheta partial sums in registers
//   phase 2 (B2)  a warp per source node: dX[u,h] = sum over its out-entries of Gs[v,h] -- every gather a shared-memory
//                 read -- written (or accumulated, strided) to global
//   phase 3 (B3)  table gradients: a thread per row builds the row's attr counts (no atomics); thread (attr slot,
//                 channel quad) then adds count * Gs[row] over the unit's rows, in row order, into the CTA's private
//                 gradient tables in shared memory
// and at the end every CTA writes its dtheta / table partials once; the existing fixed-order reductions finish.
// No float atomics, fixed assignment of units to CTAs, fixed order inside a unit: bit-reproducible.
#include "agg_fast_host.h"
#include "agg_lean.cuh"

namespace kp {

constexpr int FB_THREADS = 512;
constexpr int FB_WARPS = FB_THREADS / 32;
constexpr int FB_MAXHG = 4;            // hops per unit (register accumulators for dtheta: HG float4 per lane)
constexpr int FB_TROWS = 64;           // embedding rows per table class the count matrix covers (attr < 64)

struct FbArgs {
  const int32_t* block_ptr;
  const int32_t* block_stats;
  int num_blocks, HG, groups;          // hops per unit, units per block = ceil(k / HG)
  int maxn;                            // node capacity of the shared Gs tile
  int rows0, rowsk;                    // table rows (0 = no tables)
  const float* dOut;
  float* dP;                           // [N,k,d] or NULL
  float* dX;                           // base of dX (strides in fa.os / fa.oh, accumulate in fa.oacc) or NULL
  float* dth_part;                     // [grid][k][d] or NULL
  float* tab_part;                     // [grid][rows0 + rowsk][d] or NULL
};

// shared layout (floats): theta [k*d] | Gs [maxn][HG][d] | dT [rows0+rowsk][d] | cnt (bytes) [maxn*HG][FB_TROWS] | masks
template <int ACT, bool FUSE, bool TAB>
__global__ void __launch_bounds__(FB_THREADS, 2)
agg_block_bwd_kernel(const FastArgs fa, const FbArgs fb) {
  extern __shared__ __align__(16) float sm[];
  const kp_agg_desc& a = fa.d;
  const int d = a.d, k = a.k, Kp = a.Kplan, dq = d >> 2;
  const int HG = fb.HG;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool active = lane * 4 < d;
  const unsigned c = (unsigned)min(lane * 4, d - 4);
  float* theta_s = sm;
  float* Gs_s = theta_s + (FUSE ? k * d : 0);
  float* dT_s = Gs_s + (size_t)fb.maxn * HG * d;
  const int trows = TAB ? fb.rows0 + fb.rowsk : 0;
  unsigned char* cnt_s = reinterpret_cast<unsigned char*>(dT_s + (size_t)trows * d);
  unsigned long long* mask_s = reinterpret_cast<unsigned long long*>(cnt_s + (size_t)fb.maxn * HG * FB_TROWS);
  if (FUSE)
    for (int i = threadIdx.x * 4; i < k * d; i += FB_THREADS * 4) st4(theta_s + i, ld4(a.theta + i));
  for (int i = threadIdx.x * 4; i < trows * d; i += FB_THREADS * 4) st4(dT_s + i, make_float4(0.f, 0.f, 0.f, 0.f));
  const bool need_z = FUSE && fb.dth_part != nullptr;
  const bool hasP = need_z && a.P != nullptr;
  // A CTA serves ONE hop group for its whole life (group = blockIdx.x % groups; the grid is a multiple of `groups`):
  // the dtheta sums of the group's hops stay in registers across all of the CTA's units and are reduced over the warps
  // once, at the end, through the (then free) Gs tile.
  float4 dth[FB_MAXHG];
#pragma unroll
  for (int i = 0; i < FB_MAXHG; ++i) dth[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int nblocks = fb.block_stats ? __ldg(fb.block_stats) : fb.num_blocks;
  const int g = blockIdx.x % fb.groups;
  const int h0 = g * HG, h1 = min(k, h0 + HG), nh = h1 - h0;
  const float* Xg = a.X + (size_t)h0 * fa.xh + c;
  for (int b = blockIdx.x / fb.groups; b < nblocks; b += gridDim.x / fb.groups) {
    const int v0 = __ldg(fb.block_ptr + b), v1 = __ldg(fb.block_ptr + b + 1);
    const int nb = v1 - v0;
    if (nb > fb.maxn) continue;        // cannot happen when the caller sized maxn from the plan statistics
    // ---------------------------------------------------------------- phase 1: B1, a warp per destination node
    for (int vl = warp; vl < nb; vl += FB_WARPS) {
      const int v = v0 + vl;
      const int rp = (lane <= nh) ? __ldg(a.rowptr + (size_t)v * Kp + h0 + lane) : 0;
      float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
      if (FUSE) go = ld4s(fb.dOut + ((size_t)v * d + c));
      const int eb = __shfl_sync(0xffffffffu, rp, 0), ee = __shfl_sync(0xffffffffu, rp, nh);
      // the node's entries of this hop group, one per lane (molecules: ~10; longer lists re-load per 32)
      int mycol = 0, myattr = 0;
      if (eb + lane < ee) {
        mycol = __ldg(a.col + eb + lane);
        if (TAB) myattr = (int)__ldg(a.attr16 + eb + lane);
      }
      int wb = eb;                                                    // first entry held in the lanes
      for (int hl = 0; hl < nh; ++hl) {
        const int h = h0 + hl;
        const int hb = __shfl_sync(0xffffffffu, rp, hl), he = __shfl_sync(0xffffffffu, rp, hl + 1);
        const size_t row = ((size_t)v * k + h) * d + c;
        // independent loads of the hop first: dOut / P rows, then the gathers four at a time
        float4 dy = make_float4(0.f, 0.f, 0.f, 0.f), p = dy;
        if (!FUSE) dy = ld4s(fb.dOut + row);
        if (hasP) p = ld4s(a.P + ((size_t)v * fa.ps + (size_t)h * fa.ph + c));
        const float* Xh = Xg + (size_t)hl * fa.xh;
        const float* Th = TAB ? ((h == 0 ? a.T0 : a.Tk) + c) : nullptr;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = hb; j < he; j += 4) {
          if (j + 4 > wb + 32) {                                     // (rare) window moves: reload 32 entries from j
            wb = j;
            mycol = 0; myattr = 0;
            if (wb + lane < ee) {
              mycol = __ldg(a.col + wb + lane);
              if (TAB) myattr = (int)__ldg(a.attr16 + wb + lane);
            }
          }
          float4 x[4], t[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const unsigned cj = (unsigned)__shfl_sync(0xffffffffu, mycol, (j - wb + i) & 31);
            const int aj = TAB ? __shfl_sync(0xffffffffu, myattr, (j - wb + i) & 31) : 0;
            x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            t[i] = x[i];
            if (j + i < he) {
              x[i] = ld4(Xh + cj * fa.xs);
              if (TAB) t[i] = ld4(Th + aj * d);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            add4(x[i], t[i]);
            add4(acc, x[i]);
          }
        }
        if (FUSE) {
          const float4 th = *reinterpret_cast<const float4*>(theta_s + h * d + c);
          dy = make_float4(th.x * go.x, th.y * go.y, th.z * go.z, th.w * go.w);
        }
        if (fb.dP && active) st4s(fb.dP + row, dy);
        const float4 gsv = make_float4(dy.x * act_bwd<ACT>(acc.x), dy.y * act_bwd<ACT>(acc.y),
                                       dy.z * act_bwd<ACT>(acc.z), dy.w * act_bwd<ACT>(acc.w));
        if (active) *reinterpret_cast<float4*>(Gs_s + ((size_t)vl * HG + hl) * d + c) = gsv;
        if (need_z) {
          const float4 z = make_float4(act_fwd<ACT>(acc.x) + p.x, act_fwd<ACT>(acc.y) + p.y, act_fwd<ACT>(acc.z) + p.z,
                                       act_fwd<ACT>(acc.w) + p.w);
#pragma unroll
          for (int i = 0; i < FB_MAXHG; ++i)
            if (i == hl) {
              dth[i].x = fmaf(go.x, z.x, dth[i].x); dth[i].y = fmaf(go.y, z.y, dth[i].y);
              dth[i].z = fmaf(go.z, z.z, dth[i].z); dth[i].w = fmaf(go.w, z.w, dth[i].w);
            }
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase 2: B2, a warp per source node
    if (fb.dX) {
      const unsigned os = fa.os ? fa.os : (unsigned)(k * d), oh = fa.oh ? fa.oh : (unsigned)d;
      for (int ul = warp; ul < nb; ul += FB_WARPS) {
        const int uu = v0 + ul;
        const int rp = (lane <= nh) ? __ldg(a.rowptrT + (size_t)uu * Kp + h0 + lane) : 0;
        const int eb = __shfl_sync(0xffffffffu, rp, 0), ee = __shfl_sync(0xffffffffu, rp, nh);
        int myv = (eb + lane < ee) ? __ldg(a.colT + eb + lane) - v0 : 0;
        int wb = eb;
        for (int hl = 0; hl < nh; ++hl) {
          const int rb = __shfl_sync(0xffffffffu, rp, hl), re = __shfl_sync(0xffffffffu, rp, hl + 1);
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          float* po = fb.dX + ((size_t)uu * os + (size_t)(h0 + hl) * oh + c);
          float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
          if (fa.oacc && active) old = *reinterpret_cast<const float4*>(po);
          for (int j = rb; j < re; j += 4) {
            if (j + 4 > wb + 32) {
              wb = j;
              myv = (wb + lane < ee) ? __ldg(a.colT + wb + lane) - v0 : 0;
            }
            float4 gq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int vl = __shfl_sync(0xffffffffu, myv, (j - wb + i) & 31);
              gq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (j + i < re) gq[i] = *reinterpret_cast<const float4*>(Gs_s + ((size_t)vl * HG + hl) * d + c);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) add4(acc, gq[i]);
          }
          if (active) {
            add4(acc, old);
            *reinterpret_cast<float4*>(po) = acc;
          }
        }
      }
    }
    // ---------------------------------------------------------------- phase 3: B3, table gradients of this unit
    if (TAB && fb.tab_part) {
      const int nrows = nb * nh;
      // (i) count matrix: thread r owns row r = (vl, hl); counts of each attr value among the row's entries
      for (int i = threadIdx.x; i < nb * HG * (FB_TROWS / 4); i += FB_THREADS) reinterpret_cast<unsigned*>(cnt_s)[i] = 0u;
      if (threadIdx.x < 2) mask_s[threadIdx.x] = 0ull;
      __syncthreads();
      for (int r = threadIdx.x; r < nrows; r += FB_THREADS) {
        const int vl = r / nh, hl = r - vl * nh;
        const int rr = (v0 + vl) * Kp + h0 + hl;
        const int rb = __ldg(a.rowptr + rr), re = __ldg(a.rowptr + rr + 1);
        unsigned char* crow = cnt_s + (size_t)(vl * HG + hl) * FB_TROWS;
        unsigned long long m = 0ull;
        for (int j = rb; j < re; ++j) {
          const int aj = (int)__ldg(a.attr16 + j);
          if (aj < FB_TROWS) {
            ++crow[aj];
            m |= 1ull << aj;
          }
        }
        if (m) atomicOr(&mask_s[(h0 + hl) == 0 ? 0 : 1], m);         // integer OR: order-independent
      }
      __syncthreads();
      // (ii) thread (slot, quad): slot-th attr value present in the unit (class 0 = hop-1 table T0, class 1 = Tk)
      const int nslots = FB_THREADS / dq;
      const int slot = threadIdx.x / dq, q = threadIdx.x - slot * dq;
      if (slot < nslots) {
#pragma unroll 1
        for (int cls = 0; cls < 2; ++cls) {
          if (cls == 0 && h0 != 0) continue;                         // hop 0 lives in the first group only
          const int hl_lo = (cls == 0) ? 0 : (h0 == 0 ? 1 : 0);      // hops of this class inside the group
          const int hl_hi = (cls == 0) ? 1 : nh;
          const unsigned long long m = mask_s[cls];
          const int present = __popcll(m);
          for (int s0 = slot; s0 < present; s0 += nslots) {
            unsigned long long mm = m;                               // s0-th set bit
            for (int i = 0; i < s0; ++i) mm &= mm - 1;
            const int t = __ffsll((long long)mm) - 1;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int vl = 0; vl < nb; ++vl)
              for (int hl = hl_lo; hl < hl_hi; ++hl) {
                const unsigned cn = cnt_s[(size_t)(vl * HG + hl) * FB_TROWS + t];
                if (cn) fma4(acc, (float)cn, *reinterpret_cast<const float4*>(Gs_s + ((size_t)vl * HG + hl) * d + 4 * q));
              }
            const int trow = cls == 0 ? t : fb.rows0 + t;
            if (t < (cls == 0 ? fb.rows0 : fb.rowsk)) {
              float4* p = reinterpret_cast<float4*>(dT_s + (size_t)trow * d + 4 * q);
              float4 sv = *p;
              sv.x += acc.x; sv.y += acc.y; sv.z += acc.z; sv.w += acc.w;
              *p = sv;
            }
          }
        }
      }
    }
    __syncthreads();                    // Gs / count tile free for the next unit
  }
  // ------------------------------------------------------------------ per-CTA partials, written once
  if (need_z) {
    // fixed-order sum of the warps' register sums through the Gs tile (free now), in rounds of `per_round` warps
    float* stage = Gs_s;
    float* mine = fb.dth_part + (size_t)blockIdx.x * k * d;
    for (int i = threadIdx.x; i < k * d; i += FB_THREADS) mine[i] = 0.f;
    __syncthreads();
    const int per_round = max(1, min(FB_WARPS, fb.maxn));
    for (int w0 = 0; w0 < FB_WARPS; w0 += per_round) {
      if (warp >= w0 && warp < w0 + per_round && active) {
#pragma unroll
        for (int i = 0; i < FB_MAXHG; ++i)
          if (i < nh) *reinterpret_cast<float4*>(stage + ((size_t)(warp - w0) * HG + i) * d + c) = dth[i];
      }
      __syncthreads();
      const int nw = min(per_round, FB_WARPS - w0);
      for (int i = threadIdx.x; i < nh * d; i += FB_THREADS) {
        const int hl = i / d, cc = i - hl * d;
        float sv = mine[(size_t)(h0 + hl) * d + cc];
        for (int w = 0; w < nw; ++w) sv += stage[((size_t)w * HG + hl) * d + cc];
        mine[(size_t)(h0 + hl) * d + cc] = sv;
      }
      __syncthreads();
    }
  }
  if (TAB && fb.tab_part) {
    float* out = fb.tab_part + (size_t)blockIdx.x * trows * d;
    for (int i = threadIdx.x * 4; i < trows * d; i += FB_THREADS * 4) st4(out + i, *reinterpret_cast<const float4*>(dT_s + i));
  }
}

static int g_fb_mode = 1;               // 0 = never, 1 = when the caller supplies closed blocks (default)
void block_bwd_set_mode(int mode) { g_fb_mode = mode; }

static size_t fb_smem(const kp_agg_desc& a, int HG, bool need_z, int maxn) {
  const int trows = a.T0 ? a.rows0 + a.rowsk : 0;
  size_t b = sizeof(float) * ((a.fuse ? (size_t)a.k * a.d : 0) + (size_t)maxn * HG * a.d + (size_t)trows * a.d);
  b += (size_t)maxn * HG * FB_TROWS + 16;
  (void)need_z;
  return (b + 15) & ~(size_t)15;
}

// hops per unit: the largest group (<= FB_MAXHG) whose tile leaves room for two CTAs per SM
static int fb_pick_hg(const kp_agg_desc& a, bool need_z) {
  for (int hg = a.k < FB_MAXHG ? a.k : FB_MAXHG; hg >= 1; --hg)
    if (fb_smem(a, hg, need_z, a.max_block_nodes) <= 110 * 1024) return hg;
  return 0;
}

bool block_bwd_eligible(const kp_agg_desc& a, int G, int tab, bool want_dtheta) {
  if (g_fb_mode == 0 || !a.block_ptr || a.num_blocks <= 0 || a.max_block_nodes <= 0) return false;
  if (G != 32 || a.dinv || a.indeg || a.eps) return false;
  if (a.act != KP_ACT_GELU) return false;
  if (tab == TAB_GLOBAL) return false;
  if (a.T0 && (a.rows0 > FB_TROWS || a.rowsk > FB_TROWS || a.amax0 < 0 || a.amax0 >= FB_TROWS ||
               (a.k > 1 && (a.amaxk < 0 || a.amaxk >= FB_TROWS))))
    return false;
  if (a.max_block_nodes > 255) return false;      // count-matrix bytes: a row holds at most one entry per source node
  return fb_pick_hg(a, a.fuse && want_dtheta) > 0;
}

int block_bwd_grid() { return kNumSMs * 2; }

template <int ACT, bool FUSE>
static int fb_launch(const FastArgs& fa, const FbArgs& fb, bool tab, int grid, size_t smem, cudaStream_t st) {
  if (tab) {
    if (smem > 32 * 1024)
      KP_CUDA(cudaFuncSetAttribute(agg_block_bwd_kernel<ACT, FUSE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    KP_LAUNCH((agg_block_bwd_kernel<ACT, FUSE, true>), grid, FB_THREADS, smem, st, fa, fb);
  } else {
    if (smem > 32 * 1024)
      KP_CUDA(cudaFuncSetAttribute(agg_block_bwd_kernel<ACT, FUSE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    KP_LAUNCH((agg_block_bwd_kernel<ACT, FUSE, false>), grid, FB_THREADS, smem, st, fa, fb);
  }
  return 0;
}

// dth_part: [block_bwd_grid()][k][d] or NULL; tab_part: [block_bwd_grid()][rows0+rowsk][d] or NULL
int block_bwd(const FastArgs& fa, const float* dOut, float* dX, float* dP, float* dth_part, float* tab_part, int* grid_out,
              cudaStream_t st) {
  const kp_agg_desc& a = fa.d;
  FbArgs fb;
  const bool need_z = a.fuse && dth_part;
  fb.block_ptr = a.block_ptr;
  fb.block_stats = a.block_stats;
  fb.num_blocks = a.num_blocks;
  fb.HG = fb_pick_hg(a, need_z);
  fb.groups = (a.k + fb.HG - 1) / fb.HG;
  fb.maxn = a.max_block_nodes;
  fb.rows0 = a.T0 ? a.rows0 : 0;
  fb.rowsk = a.T0 ? a.rowsk : 0;
  fb.dOut = dOut;
  fb.dP = dP;
  fb.dX = dX;
  fb.dth_part = dth_part;
  fb.tab_part = a.T0 ? tab_part : nullptr;
  FastArgs f2 = fa;
  f2.os = (unsigned)a.dx_node_stride;
  f2.oh = (unsigned)a.dx_hop_stride;
  f2.oacc = a.dx_accumulate;
  const size_t smem = fb_smem(a, fb.HG, need_z, fb.maxn);
  const long long units = (long long)a.num_blocks * fb.groups;
  int grid = geom_cap(units < block_bwd_grid() ? units : block_bwd_grid());
  grid = grid / fb.groups * fb.groups;           // a CTA serves one hop group: group = blockIdx.x % groups
  if (grid < fb.groups) grid = fb.groups;
  *grid_out = grid;
  return a.fuse ? fb_launch<KP_ACT_GELU, true>(f2, fb, a.T0 != nullptr, grid, smem, st)
                : fb_launch<KP_ACT_GELU, false>(f2, fb, a.T0 != nullptr, grid, smem, st);
}

}  // namespace kp
