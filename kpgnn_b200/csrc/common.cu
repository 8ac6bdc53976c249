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
