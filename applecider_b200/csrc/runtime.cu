// Library management: error string, version, launch counter.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void acb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void acb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static const unsigned long long* g_seed_epoch = nullptr;
const unsigned long long* acb_seed_epoch_ptr() { return g_seed_epoch; }

extern "C" {
const char* acb_last_error(void) { return g_err; }
int acb_version(void) { return 100; }
long long acb_launch_count(void) { return g_launches.load(); }
void acb_reset_launch_count(void) { g_launches.store(0); }
int acb_set_seed_epoch_ptr(const unsigned long long* dev_ptr) {
  g_seed_epoch = dev_ptr;
  return ACB_OK;
}
}
