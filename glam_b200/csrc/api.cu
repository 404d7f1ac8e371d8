// Library-level entry points: version, error string, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace glam {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace glam

extern "C" int glam_abi_version(void) { return GLAM_B200_ABI_VERSION; }
extern "C" const char* glam_last_error(void) { return glam::g_err; }
extern "C" int64_t glam_launch_count(void) { return glam::g_launches.load(std::memory_order_relaxed); }
