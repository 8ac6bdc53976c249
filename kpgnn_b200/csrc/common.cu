// common.cu -- error string, ABI version, launch counter.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace kp {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace kp

extern "C" {
const char* kp_last_error(void) { return kp::g_err; }
int kp_abi_version(void) { return KPGNN_ABI_VERSION; }
uint64_t kp_launch_count(void) { return kp::g_launches.load(std::memory_order_relaxed); }
}

namespace kp {
cudaError_t fork_stream(cudaStream_t from, cudaStream_t to) {
  constexpr int kRing = 256;
  static thread_local cudaEvent_t ring[kRing];
  static thread_local int made = 0, next = 0;
  if (made < kRing) {
    cudaError_t e = cudaEventCreateWithFlags(&ring[made], cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    ++made;
  }
  cudaEvent_t ev = ring[next % made];
  next = (next + 1) % kRing;
  cudaError_t e = cudaEventRecord(ev, from);
  if (e != cudaSuccess) return e;
  return cudaStreamWaitEvent(to, ev, 0);
}
}  // namespace kp
