// ABI bookkeeping: version, thread-local error string, process-wide launch counter.
#include <atomic>
#include <cstdlib>

#include "osr_common.cuh"

namespace osr {

static thread_local char g_err[512] = {0};
static std::atomic<long long> g_launches{0};  // process-wide: autograd backward runs on other threads

char* tls_error_buf() { return g_err; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches += n; }

// variant switches: initialised once (static initialiser of this translation unit) from OSR_TUNE_<KEY>
static const char* const kTuneNames[OSR_TUNE_COUNT] = {"OSR_TUNE_BWD_VARIANT", "OSR_TUNE_FWD_VARIANT", "OSR_TUNE_PLN_VARIANT",
                                                       "OSR_TUNE_RPN_VARIANT", "OSR_TUNE_NMS_VARIANT", "OSR_TUNE_BWD_SPLIT"};
static std::atomic<int> g_tune[OSR_TUNE_COUNT];
static const bool g_tune_init = [] {
  for (int k = 0; k < OSR_TUNE_COUNT; ++k) {
    const char* e = getenv(kTuneNames[k]);
    g_tune[k] = e ? atoi(e) : 0;
  }
  return true;
}();
int tuning(int key) { return g_tune[key].load(std::memory_order_relaxed); }

}  // namespace osr

extern "C" {

int osr_version(void) { return OSR_ABI_VERSION; }
const char* osr_last_error(void) { return osr::g_err; }
long long osr_launch_count(void) { return osr::g_launches.load(); }
void osr_reset_launch_count(void) { osr::g_launches = 0; }
int osr_set_tuning(int key, int value) {
  if (key < 0 || key >= OSR_TUNE_COUNT) return osr::fail_arg(OSR_E_ARG, "osr_set_tuning: unknown key %d", key);
  return osr::g_tune[key].exchange(value);
}
int osr_get_tuning(int key) {
  if (key < 0 || key >= OSR_TUNE_COUNT) return osr::fail_arg(OSR_E_ARG, "osr_get_tuning: unknown key %d", key);
  return osr::g_tune[key].load();
}

}  // extern "C"
