// libtdb200 core: version, thread-local error text, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

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

// ---- optional per-kernel timing ----------------------------------------------------------------------------------
static thread_local cudaStream_t g_stream = nullptr;
static thread_local bool g_timing = false;
struct Mark {
  const char *name;
  cudaEvent_t ev;
};
static thread_local std::vector<Mark> g_marks;

void note_stream(cudaStream_t s) { g_stream = s; }

static void mark(const char *name) {
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  cudaEventRecord(ev, g_stream);
  g_marks.push_back(Mark{name, ev});
}

void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char *what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_timing) mark(what);
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

void tdb_timing_begin(tdb_stream_t stream) {
  using namespace tdb;
  for (auto &m : g_marks) cudaEventDestroy(m.ev);
  g_marks.clear();
  g_stream = reinterpret_cast<cudaStream_t>(stream);
  g_timing = true;
  mark("<begin>");
}

size_t tdb_timing_end(char *buf, size_t buf_bytes) {
  using namespace tdb;
  g_timing = false;
  std::map<std::string, std::pair<int, double>> agg;
  if (!g_marks.empty()) {
    cudaEventSynchronize(g_marks.back().ev);
    for (size_t i = 1; i < g_marks.size(); i++) {
      float ms = 0.0f;
      cudaEventElapsedTime(&ms, g_marks[i - 1].ev, g_marks[i].ev);
      auto &slot = agg[g_marks[i].name];
      slot.first += 1, slot.second += ms;
    }
  }
  std::string out;
  char line[256];
  for (auto &kv : agg) {
    snprintf(line, sizeof line, "%s,%d,%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  for (auto &m : g_marks) cudaEventDestroy(m.ev);
  g_marks.clear();
  if (buf && buf_bytes > 0) {
    const size_t n = out.size() < buf_bytes - 1 ? out.size() : buf_bytes - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return out.size() + 1;
}
}
