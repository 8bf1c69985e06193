// Library-wide bookkeeping for libocr_b200.so: error text, version, launch counter.
#include <atomic>
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace ocr {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace ocr

extern "C" const char* ocr_last_error(void) { return ocr::g_err; }
extern "C" const char* ocr_version(void) { return "0.1;sm_100a;" __DATE__ " " __TIME__; }
extern "C" uint64_t ocr_launch_count(void) { return ocr::g_launches.load(std::memory_order_relaxed); }
