// ABI bookkeeping: version, thread-local error string, process-wide launch counter.
#include <atomic>

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

}  // namespace osr

extern "C" {

int osr_version(void) { return OSR_ABI_VERSION; }
const char* osr_last_error(void) { return osr::g_err; }
long long osr_launch_count(void) { return osr::g_launches.load(); }
void osr_reset_launch_count(void) { osr::g_launches = 0; }

}  // extern "C"
