// espresso_b200 -- C-ABI plumbing: error strings, launch counter, cached device attributes.
#include "common.cuh"
#include "espresso_b200.h"
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace {
thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void esp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void esp_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int esp_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

bool esp_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("ESP_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

extern "C" const char* esp_last_error(void) { return g_err; }
extern "C" int esp_version(void) { return 100; }
extern "C" int64_t esp_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" void esp_note_graph_replay(int64_t launches) { g_launches.fetch_add(launches, std::memory_order_relaxed); }
