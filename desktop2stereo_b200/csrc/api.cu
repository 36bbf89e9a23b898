// Library-wide state of libd2s_b200: error string, launch counter, version.
#include <stdarg.h>

#include "common.cuh"

namespace d2s {
thread_local std::string g_last_error;
std::atomic<long long> g_launch_count{0};
thread_local bool g_pdl = false;

int set_error(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
}  // namespace d2s

extern "C" const char *d2s_last_error(void) { return d2s::g_last_error.c_str(); }
extern "C" const char *d2s_version(void) { return "d2s_b200 0.1 (sm_100a)"; }
extern "C" int64_t d2s_launch_count(void) { return (int64_t)d2s::g_launch_count.load(); }
