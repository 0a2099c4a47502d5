// csrc/snb_api.cu -- version, thread-local error string, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "snb_common.h"

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void snb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void snb_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

extern "C" int snb_version(void) { return SNB_VERSION; }
extern "C" const char *snb_last_error(void) { return g_err; }
extern "C" uint64_t snb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
