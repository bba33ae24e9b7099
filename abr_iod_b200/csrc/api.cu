// api.cu -- library-wide entry points of libabr_b200: version, thread-local error text, launch counter.
#include <cstdlib>
#include <cstring>
#include <vector>

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

// ---- stage timing: (ABR_FUSED_STAGES + 1) events per recorded call
static std::vector<cudaEvent_t> g_stage_events;
static int g_stage_max = 0, g_stage_calls = 0, g_stage_every = 1, g_stage_seen = 0;
static bool g_stage_on = false;
constexpr int kMarks = ABR_FUSED_STAGES + 1;

void stage_mark(cudaStream_t st, int k) {
  if (!g_stage_on) return;
  const bool sampled = g_stage_seen % g_stage_every == 0 && g_stage_calls < g_stage_max;  // every g_stage_every-th call
  if (sampled) cudaEventRecord(g_stage_events[(size_t)g_stage_calls * kMarks + k], st);
  if (k == ABR_FUSED_STAGES) {
    if (sampled) g_stage_calls++;
    g_stage_seen++;
  }
}

}  // namespace abr

extern "C" {
int abr_stage_timing_begin(int max_calls) { return abr_stage_timing_begin_every(max_calls, 1); }
int abr_stage_timing_begin_every(int max_calls, int every) {
  using namespace abr;
  ABR_REQUIRE(max_calls > 0 && max_calls <= 4096, ABR_ERR_BAD_ARG, "stage_timing_begin: max_calls %d (1..4096)", max_calls);
  ABR_REQUIRE(every > 0, ABR_ERR_BAD_ARG, "stage_timing_begin: every %d (>= 1)", every);
  g_stage_every = every;
  g_stage_seen = 0;
  while ((int)g_stage_events.size() < max_calls * kMarks) {
    cudaEvent_t e;
    ABR_CUDA_OK(cudaEventCreate(&e));
    g_stage_events.push_back(e);
  }
  g_stage_max = max_calls;
  g_stage_calls = 0;
  g_stage_on = true;
  return ABR_OK;
}
int abr_stage_timing_end(float* avg_ms, int n_stages) {
  using namespace abr;
  g_stage_on = false;
  if (!avg_ms || n_stages < ABR_FUSED_STAGES) { set_error("stage_timing_end: room for %d stages needed", ABR_FUSED_STAGES); return -ABR_ERR_BAD_ARG; }
  for (int k = 0; k < n_stages; k++) avg_ms[k] = 0.f;
  if (g_stage_calls == 0) return 0;
  if (cudaEventSynchronize(g_stage_events[(size_t)g_stage_calls * kMarks - 1]) != cudaSuccess) { set_error("stage_timing_end: event synchronize failed"); return -ABR_ERR_CUDA; }
  for (int c = 0; c < g_stage_calls; c++)
    for (int k = 0; k < ABR_FUSED_STAGES; k++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, g_stage_events[(size_t)c * kMarks + k], g_stage_events[(size_t)c * kMarks + k + 1]);
      avg_ms[k] += ms / g_stage_calls;
    }
  return g_stage_calls;
}
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
