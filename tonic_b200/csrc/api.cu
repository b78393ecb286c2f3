// Library-level entry points: version, last error, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace tb {
static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace tb

extern "C" int tb_version(void) { return TB_VERSION; }
extern "C" const char* tb_last_error(void) { return tb::g_error; }
extern "C" int64_t tb_launch_count(void) { return tb::g_launches.load(); }
