// common.cuh -- shared host/device helpers for libkpgnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>

#include "../../include/kpgnn.h"

namespace kp {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define KP_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      kp::set_error(__VA_ARGS__);          \
      return 1;                            \
    }                                      \
  } while (0)

#define KP_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      kp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                            \
    }                                                                                      \
  } while (0)

// every kernel launch goes through this so that kp_launch_count() is the library's own evidence
#define KP_LAUNCH(kernel, grid, block, smem, stream, ...)                                  \
  do {                                                                                     \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);              \
    kp::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
    KP_CUDA(cudaGetLastError());                                                           \
  } while (0)

// Launch with the programmatic-dependent-launch attribute -- OPT-IN (KP_AGG_PDL=1 for the aggregation kernels,
// KP_DENSE_PDL=1 for the dense block): parity-tested with both on (full GPU suite), but measured with CUDA events
// inside the captured step it gains nothing (dense: 1.1786 vs 1.1748 ms) or loses (aggregation: 1.1913 ms) --
// profiles/r2_pdl.txt.  With the attribute the kernel may be scheduled
// while the preceding kernel of the stream is still running, once every CTA of that kernel has called
// kp_pdl_trigger() (or exited); it must call kp_pdl_wait() before touching anything the predecessor writes.
#define KP_LAUNCH_PDL(kernel, grid, block, smem, strm_, ...)                                                \
  do {                                                                                                      \
    static const bool _pdl = getenv("KP_AGG_PDL") && atoi(getenv("KP_AGG_PDL")) == 1;                     \
    cudaLaunchConfig_t _cfg = {};                                                                           \
    _cfg.gridDim = dim3(grid);                                                                              \
    _cfg.blockDim = dim3(block);                                                                            \
    _cfg.dynamicSmemBytes = (size_t)(smem);                                                                 \
    _cfg.stream = (cudaStream_t)(strm_);                                                                    \
    cudaLaunchAttribute _at[1];                                                                             \
    _at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                         \
    _at[0].val.programmaticStreamSerializationAllowed = 1;                                                  \
    _cfg.attrs = _at;                                                                                       \
    _cfg.numAttrs = _pdl ? 1 : 0;                                                                           \
    cudaError_t _le = cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__);                                       \
    kp::g_launches.fetch_add(1, std::memory_order_relaxed);                                                 \
    KP_CUDA(_le);                                                                                           \
  } while (0)

// Kernels that spin on a grid-wide barrier: a cooperative launch makes the driver guarantee that every CTA of the grid
// is co-resident (it fails with cudaErrorCooperativeLaunchTooLarge instead of deadlocking on a smaller part / MIG
// slice); capturable in CUDA graphs like a plain launch.  `args` is the usual array of pointers to the arguments.
#define KP_LAUNCH_COOP(kernel, grid, block, smem, strm_, args)                                             \
  do {                                                                                                      \
    static const bool _plain = getenv("KP_DENSE_COOP") && atoi(getenv("KP_DENSE_COOP")) == 0;               \
    static const bool _pdl = getenv("KP_DENSE_PDL") && atoi(getenv("KP_DENSE_PDL")) == 1;                   \
    cudaLaunchConfig_t _cfg = {};                                                                           \
    _cfg.gridDim = dim3(grid);                                                                              \
    _cfg.blockDim = dim3(block);                                                                            \
    _cfg.dynamicSmemBytes = (size_t)(smem);                                                                 \
    _cfg.stream = (cudaStream_t)(strm_);                                                                    \
    cudaLaunchAttribute _at[2];                                                                             \
    unsigned _na = 0;                                                                                       \
    if (!_plain) {                                                                                          \
      _at[_na].id = cudaLaunchAttributeCooperative;                                                         \
      _at[_na].val.cooperative = 1;                                                                         \
      ++_na;                                                                                                \
    }                                                                                                       \
    if (_pdl) { /* programmatic dependent launch: the kernel may start while its predecessor drains; it   */ \
      _at[_na].id = cudaLaunchAttributeProgrammaticStreamSerialization; /* calls griddepcontrol.wait before */ \
      _at[_na].val.programmaticStreamSerializationAllowed = 1;          /* touching the predecessor's data  */ \
      ++_na;                                                                                                \
    }                                                                                                       \
    _cfg.attrs = _at;                                                                                       \
    _cfg.numAttrs = _na;                                                                                    \
    cudaError_t _le = cudaLaunchKernelExC(&_cfg, (const void*)(kernel), (args));                            \
    kp::g_launches.fetch_add(1, std::memory_order_relaxed);                                                 \
    if (_le != cudaSuccess) {                                                                               \
      kp::set_error("cooperative launch of %s failed: %s (%s:%d)", #kernel, cudaGetErrorString(_le), __FILE__, __LINE__); \
      return 2;                                                                                             \
    }                                                                                                       \
  } while (0)

// Programmatic dependent launch, device side.  kp_pdl_wait(): every memory operation of the preceding kernel in the stream
// is complete and visible (a no-op when the kernel was launched normally).  kp_pdl_trigger(): kernels launched behind
// this one with the PDL attribute may start now (they still wait in kp_pdl_wait for this grid to finish).
__device__ __forceinline__ void kp_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void kp_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Fork helper: make `to` wait for everything enqueued on `from` so far.  Events come from a small per-thread ring and
// are never destroyed while in flight (create + record + wait + destroy around every fork is legal, but keeping the
// event object alive keeps the dependency valid under every driver / capture mode).
cudaError_t fork_stream(cudaStream_t from, cudaStream_t to);

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

}  // namespace kp
