// Library-level entry points: version, last error, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <map>
#include <string>
#include <vector>

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

namespace tb {
// ---- event profiler ----------------------------------------------------------
struct ProfEvent { const char* name; cudaEvent_t start, stop; };
static std::vector<ProfEvent> g_events;
static std::vector<cudaEvent_t> g_pool;
static bool g_profiling = false;

bool profiling_enabled() { return g_profiling; }

static cudaEvent_t take_event() {
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void profile_mark(const char* name, cudaStream_t stream, bool begin) {
    if (begin) {
        ProfEvent ev{name, take_event(), take_event()};
        cudaEventRecord(ev.start, stream);
        g_events.push_back(ev);
    } else if (!g_events.empty()) {
        cudaEventRecord(g_events.back().stop, stream);
    }
}
}  // namespace tb

extern "C" int tb_profile_begin(void) {
    tb::g_events.clear();
    tb::g_profiling = true;
    return 0;
}

// Synchronises, then writes "name count total_ms\n" lines (one per kernel entry
// point) into buf.  Returns the number of bytes needed.
extern "C" int tb_profile_end(char* buf, int32_t size) {
    tb::g_profiling = false;
    cudaDeviceSynchronize();
    std::map<std::string, std::pair<long, double>> acc;
    for (auto& ev : tb::g_events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev.start, ev.stop) == cudaSuccess) {
            auto& a = acc[ev.name];
            a.first += 1;
            a.second += ms;
        }
        tb::g_pool.push_back(ev.start);
        tb::g_pool.push_back(ev.stop);
    }
    tb::g_events.clear();
    cudaGetLastError();
    std::string out;
    char line[256];
    for (auto& kv : acc) {
        snprintf(line, sizeof(line), "%s %ld %.6f\n", kv.first.c_str(), kv.second.first,
                 kv.second.second);
        out += line;
    }
    if (buf && size > 0) {
        strncpy(buf, out.c_str(), (size_t)size - 1);
        buf[size - 1] = 0;
    }
    return (int)out.size() + 1;
}

extern "C" int tb_version(void) { return TB_VERSION; }
extern "C" const char* tb_last_error(void) { return tb::g_error; }
extern "C" int64_t tb_launch_count(void) { return tb::g_launches.load(); }
