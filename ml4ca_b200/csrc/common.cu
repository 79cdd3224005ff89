// common.cu -- last-error string, launch counter, version.
#include <stdarg.h>
#include <atomic>

#include "common.h"

namespace ml4ca {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace ml4ca

extern "C" {
const char* ml4ca_last_error(void) { return ml4ca::g_err; }
const char* ml4ca_version(void) { return "ml4ca_b200 0.1.0 sm_100a"; }
int64_t ml4ca_launch_count(void) { return ml4ca::g_launches.load(std::memory_order_relaxed); }
}
