// api.cu -- library-wide entry points of libabr_b200: version, thread-local error text, launch counter.
#include <cstdlib>
#include <cstring>

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

// Tuning switches (abr_set_option): kernel-family selection for A/B measurements and tests.  Initial values come from
// the environment once (ABR_ROI_V2, ABR_FWD_TMA, ABR_BWD_TMA, ABR_ARD_CLUSTER); -1 = automatic.
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
Options& options() {
  static Options o = {env_int("ABR_ROI_V2", -1), env_int("ABR_FWD_TMA", 1), env_int("ABR_BWD_TMA", 1),
                      env_int("ABR_ARD_CLUSTER", 1)};
  return o;
}

}  // namespace abr

extern "C" {
int abr_set_option(const char* key, int value) {
  ABR_REQUIRE(key, ABR_ERR_BAD_ARG, "set_option: null key");
  abr::Options& o = abr::options();
  if (!strcmp(key, "roi_v2")) o.roi_v2 = value;
  else if (!strcmp(key, "fwd_tma")) o.fwd_tma = value;
  else if (!strcmp(key, "bwd_tma")) o.bwd_tma = value;
  else if (!strcmp(key, "ard_cluster")) o.ard_cluster = value;
  else ABR_REQUIRE(false, ABR_ERR_BAD_ARG, "set_option: unknown key '%s'", key);
  return ABR_OK;
}
int abr_version(void) { return ABR_B200_VERSION; }
const char* abr_last_error(void) { return abr::g_error; }
uint64_t abr_launch_count(void) { return abr::g_launches.load(std::memory_order_relaxed); }
}
