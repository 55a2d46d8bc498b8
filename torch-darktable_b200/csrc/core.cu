// libtdb200 core: version, thread-local error text, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "tdb_common.cuh"

namespace tdb {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof g_error, fmt, ap);
  va_end(ap);
}

void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char *what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(err));
    return TDB_ECUDA;
  }
  return TDB_OK;
}

}  // namespace tdb

extern "C" {
int tdb_version(void) { return 100; }
const char *tdb_last_error(void) { return tdb::g_error; }
uint64_t tdb_launch_count(void) { return tdb::g_launches.load(std::memory_order_relaxed); }
}
