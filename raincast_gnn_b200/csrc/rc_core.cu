// Library-wide state: error text, version, launch counter.
#include <stdarg.h>
#include <stdlib.h>

#include "rc_common.cuh"

namespace rc {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("RC_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace rc

extern "C" int rc_version(void) { return 100; }
extern "C" const char* rc_last_error(void) { return rc::g_err; }
extern "C" uint64_t rc_launch_count(void) { return rc::g_launches.load(std::memory_order_relaxed); }
