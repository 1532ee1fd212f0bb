// Error reporting and bookkeeping for the C-ABI (no exceptions cross the boundary).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ps {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace ps

extern "C" const char* ps_last_error(void) { return ps::g_err; }
extern "C" int ps_abi_version(void) { return 1; }
extern "C" int64_t ps_launch_count(void) { return ps::g_launches.load(); }
