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
  float4 dth[FB_MAXHG];               // dtheta partial sums of this lane's 4 channels, hops of the CURRENT group
  // (a CTA's units alternate between hop groups; one accumulator set per group id lives in registers only for
  //  groups == 1; otherwise partial sums are flushed to the per-CTA shared staging at the end of every unit)
  float* dth_cta = reinterpret_cast<float*>(mask_s + 2);               // [FB_WARPS][k][d] staging, only if need_z
  if (need_z)
    for (int i = threadIdx.x; i < FB_WARPS * k * d; i += FB_THREADS) dth_cta[i] = 0.f;
  __syncthreads();
  const int nblocks = fb.block_stats ? __ldg(fb.block_stats) : fb.num_blocks;
  const int units = nblocks * fb.groups;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int b = u / fb.groups, g = u - b * fb.groups;
    const int h0 = g * HG, h1 = min(k, h0 + HG), nh = h1 - h0;
    const int v0 = __ldg(fb.block_ptr + b), v1 = __ldg(fb.block_ptr + b + 1);
    const int nb = v1 - v0;
    if (nb > fb.maxn) continue;        // cannot happen when the caller sized maxn from the plan statistics
    // ---------------------------------------------------------------- phase 1: B1, a warp per destination node
#pragma unroll
    for (int i = 0; i < FB_MAXHG; ++i) dth[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int vl = warp; vl < nb; vl += FB_WARPS) {
      const int v = v0 + vl;
      const int rp = (lane <= nh) ? __ldg(a.rowptr + (size_t)v * Kp + h0 + lane) : 0;
      const int eb = __shfl_sync(0xffffffffu, rp, 0), ee = __shfl_sync(0xffffffffu, rp, nh);
      float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
      if (FUSE) go = ld4s(fb.dOut + ((size_t)v * d + c));
      int hcur = 0, hend = __shfl_sync(0xffffffffu, rp, 1);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      // closes hop `hcur` of node v: activation derivative, Gs -> shared, dP -> global, dtheta partial
      auto finish_hop = [&](int hl) {
        const int h = h0 + hl;
        const size_t row = ((size_t)v * k + h) * d + c;
        float4 dy;
        if (FUSE) {
          const float4 t = *reinterpret_cast<const float4*>(theta_s + h * d + c);
          dy = make_float4(t.x * go.x, t.y * go.y, t.z * go.z, t.w * go.w);
        } else {
          dy = ld4s(fb.dOut + row);
        }
        if (fb.dP && active) st4s(fb.dP + row, dy);
        float4 gsv = make_float4(dy.x * act_bwd<ACT>(acc.x), dy.y * act_bwd<ACT>(acc.y), dy.z * act_bwd<ACT>(acc.z),
                                 dy.w * act_bwd<ACT>(acc.w));
        if (active) *reinterpret_cast<float4*>(Gs_s + ((size_t)vl * HG + hl) * d + c) = gsv;
        if (need_z) {
          float4 z = make_float4(act_fwd<ACT>(acc.x), act_fwd<ACT>(acc.y), act_fwd<ACT>(acc.z), act_fwd<ACT>(acc.w));
          if (hasP) {
            const float4 p = ld4s(a.P + ((size_t)v * fa.ps + (size_t)h * fa.ph + c));
            z.x += p.x; z.y += p.y; z.z += p.z; z.w += p.w;
          }
#pragma unroll
          for (int i = 0; i < FB_MAXHG; ++i)
            if (i == hl) {
              dth[i].x = fmaf(go.x, z.x, dth[i].x); dth[i].y = fmaf(go.y, z.y, dth[i].y);
              dth[i].z = fmaf(go.z, z.z, dth[i].z); dth[i].w = fmaf(go.w, z.w, dth[i].w);
            }
        }
      };
      for (int j0 = eb; j0 < ee; j0 += 32) {
        int mycol = 0, myattr = 0;
        if (j0 + lane < ee) {
          mycol = __ldg(a.col + j0 + lane);
          if (TAB) myattr = (int)__ldg(a.attr16 + j0 + lane);
        }
        const int cnt = min(32, ee - j0);
        for (int i = 0; i < cnt; ++i) {
          const int j = j0 + i;
          while (j >= hend) {                                        // hop boundary (also skips empty hops)
            finish_hop(hcur);
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
            ++hcur;
            hend = __shfl_sync(0xffffffffu, rp, hcur + 1);
          }
          const unsigned cj = (unsigned)__shfl_sync(0xffffffffu, mycol, i);
          float4 x = ld4(a.X + (cj * fa.xs + (unsigned)(h0 + hcur) * fa.xh + c));
          if (TAB) {
            const int aj = __shfl_sync(0xffffffffu, myattr, i);
            add4(x, ld4(((h0 + hcur) == 0 ? a.T0 : a.Tk) + aj * d + c));
          }
          add4(acc, x);
        }
      }
      while (hcur < nh) {                                            // the last hop with entries and the empty ones behind it
        finish_hop(hcur);
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
        ++hcur;
      }
    }
    if (need_z) {                       // flush this unit's dtheta sums into the warp's staging rows (fixed order: unit by unit)
#pragma unroll
      for (int i = 0; i < FB_MAXHG; ++i)
        if (i < nh && active) {
          float4* p = reinterpret_cast<float4*>(dth_cta + ((size_t)warp * k + h0 + i) * d + c);
          float4 s = *p;
          s.x += dth[i].x; s.y += dth[i].y; s.z += dth[i].z; s.w += dth[i].w;
          *p = s;
        }
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase 2: B2, a warp per source node
    if (fb.dX) {
      const unsigned os = fa.os ? fa.os : (unsigned)(k * d), oh = fa.oh ? fa.oh : (unsigned)d;
      for (int ul = warp; ul < nb; ul += FB_WARPS) {
        const int uu = v0 + ul;
        const int rp = (lane <= nh) ? __ldg(a.rowptrT + (size_t)uu * Kp + h0 + lane) : 0;
        for (int hl = 0; hl < nh; ++hl) {
          const int rb = __shfl_sync(0xffffffffu, rp, hl), re = __shfl_sync(0xffffffffu, rp, hl + 1);
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int j0 = rb; j0 < re; j0 += 32) {
            const int myv = (j0 + lane < re) ? __ldg(a.colT + j0 + lane) - v0 : 0;
            const int cnt = min(32, re - j0);
            for (int i = 0; i < cnt; ++i) {
              const int vl = __shfl_sync(0xffffffffu, myv, i);
              add4(acc, *reinterpret_cast<const float4*>(Gs_s + ((size_t)vl * HG + hl) * d + c));
            }
          }
          if (active) {
            float* po = fb.dX + ((size_t)uu * os + (size_t)(h0 + hl) * oh + c);
            if (fa.oacc) add4(acc, *reinterpret_cast<const float4*>(po));
            *reinterpret_cast<float4*>(po) = acc;
          }
        }
      }
    }
    // ---------------------------------------------------------------- phase 3: B3, table gradients of this unit
    if (TAB && fb.tab_part) {
      const int nrows = nb * nh;
      // (i) count matrix: thread r owns row r = (vl, hl); counts of each attr value among the row's entries
      for (int i = threadIdx.x; i < nrows * (FB_TROWS / 4); i += FB_THREADS) reinterpret_cast<unsigned*>(cnt_s)[i] = 0u;
      if (threadIdx.x < 2) mask_s[threadIdx.x] = 0ull;
      __syncthreads();
      for (int r = threadIdx.x; r < nrows; r += FB_THREADS) {
        const int vl = r / nh, hl = r - vl * nh;
        const int rr = (v0 + vl) * Kp + h0 + hl;
        const int rb = __ldg(a.rowptr + rr), re = __ldg(a.rowptr + rr + 1);
        unsigned long long m = 0ull;
        for (int j = rb; j < re; ++j) {
          const int aj = (int)__ldg(a.attr16 + j);
          if (aj < FB_TROWS) {
            ++cnt_s[r * FB_TROWS + aj];
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
          const unsigned long long m = mask_s[cls];
          const int present = __popcll(m);
          for (int s0 = slot; s0 < present; s0 += nslots) {
            unsigned long long mm = m;                               // s0-th set bit
            for (int i = 0; i < s0; ++i) mm &= mm - 1;
            const int t = __ffsll((long long)mm) - 1;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < nrows; ++r) {
              const int hl = r % nh;
              if (((h0 + hl) == 0) != (cls == 0)) continue;
              const unsigned cn = cnt_s[r * FB_TROWS + t];
              if (cn) fma4(acc, (float)cn, *reinterpret_cast<const float4*>(Gs_s + (size_t)r * d + 4 * q));
            }
            const int trow = cls == 0 ? t : fb.rows0 + t;
            if (t < (cls == 0 ? fb.rows0 : fb.rowsk)) {
              float4* p = reinterpret_cast<float4*>(dT_s + (size_t)trow * d + 4 * q);
              float4 s = *p;
              s.x += acc.x; s.y += acc.y; s.z += acc.z; s.w += acc.w;
              *p = s;
            }
          }
        }
      }
    }
    __syncthreads();                    // Gs / count tile free for the next unit
  }
  // ------------------------------------------------------------------ per-CTA partials, written once
  if (need_z) {
    for (int i = threadIdx.x; i < k * d; i += FB_THREADS) {           // fixed-order sum over the CTA's warps
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < FB_WARPS; ++w) s += dth_cta[(size_t)w * k * d + i];
      fb.dth_part[(size_t)blockIdx.x * k * d + i] = s;
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
  if (need_z) b += sizeof(float) * (size_t)FB_WARPS * a.k * a.d;
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
  const int grid = geom_cap(units < block_bwd_grid() ? units : block_bwd_grid());
  *grid_out = grid;
  return a.fuse ? fb_launch<KP_ACT_GELU, true>(f2, fb, a.T0 != nullptr, grid, smem, st)
                : fb_launch<KP_ACT_GELU, false>(f2, fb, a.T0 != nullptr, grid, smem, st);
}

}  // namespace kp
