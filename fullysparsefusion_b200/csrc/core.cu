// core.cu — version, error text, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace fsfb {

static thread_local char g_err[512] = {0};
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace fsfb

extern "C" {

int fsfb_version(void) { return 2; }

const char* fsfb_last_error(void) { return fsfb::g_err; }

int64_t fsfb_launch_count(void) { return fsfb::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
