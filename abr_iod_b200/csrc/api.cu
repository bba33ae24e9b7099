// api.cu -- library-wide entry points of libabr_b200: version, thread-local error text, launch counter.
#include "common.cuh"

namespace abr {

static thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

}  // namespace abr

extern "C" {
int abr_version(void) { return ABR_B200_VERSION; }
const char* abr_last_error(void) { return abr::g_error; }
uint64_t abr_launch_count(void) { return abr::g_launches.load(std::memory_order_relaxed); }
}
