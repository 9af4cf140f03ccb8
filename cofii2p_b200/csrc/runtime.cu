// runtime.cu -- error reporting, version, launch counter.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace cofi {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace cofi

extern "C" int cofi_version(void) { return 100; }
extern "C" const char* cofi_last_error(void) { return cofi::g_err; }
extern "C" int64_t cofi_launch_count(void) { return cofi::g_launches.load(std::memory_order_relaxed); }
