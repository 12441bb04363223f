// Library-wide state: ABI version, last-error string, launch counter, SM count.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "../../include/slam_b200.h"
#include "host.cuh"

namespace slam {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace slam

extern "C" {
int slam_abi_version(void) { return SLAM_B200_ABI_VERSION; }
const char* slam_last_error(void) { return slam::g_err; }
int64_t slam_launch_count(void) { return slam::g_launches.load(std::memory_order_relaxed); }
}
