// libtdb200 core: version, thread-local error text, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
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

// ---- side stream -----------------------------------------------------------------------------------------------------
// Small "frame" launches (RCD border tiles, Wiener border tile pairs) are latency-bound and far too small to fill 148 SMs;
// they run on a per-device side stream next to the big interior kernel and are joined back before the entry point returns.
// While the per-kernel timing hook is active everything stays on the caller's stream so that the event pairs stay meaningful.
struct Side {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
};
static Side g_side[64];
static std::mutex g_side_mutex;

cudaStream_t fork_side(cudaStream_t main) {
  // Measured on B200 (tools/gpu_ab.sh, 16 x 4K frames): 6416 MP/s with the side stream against 6735 MP/s without -- the small
  // launches steal CTA slots and shared memory from the interior kernels instead of filling idle ones.  Off unless asked for.
  static const bool enabled = getenv("TDB_SIDE_STREAM") != nullptr;
  if (g_timing || !enabled) return main;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return main;
  Side &sd = g_side[dev];
  {
    std::lock_guard<std::mutex> lock(g_side_mutex);
    if (!sd.stream) {
      if (cudaStreamCreateWithFlags(&sd.stream, cudaStreamNonBlocking) != cudaSuccess) return main;
      cudaEventCreateWithFlags(&sd.fork_ev, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&sd.join_ev, cudaEventDisableTiming);
    }
  }
  cudaEventRecord(sd.fork_ev, main);
  cudaStreamWaitEvent(sd.stream, sd.fork_ev, 0);
  return sd.stream;
}

void join_side(cudaStream_t main, cudaStream_t side) {
  if (side == main) return;
  int dev = 0;
  cudaGetDevice(&dev);
  Side &sd = g_side[dev];
  cudaEventRecord(sd.join_ev, side);
  cudaStreamWaitEvent(main, sd.join_ev, 0);
}

void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static thread_local int g_concurrent_lanes = 0;
int concurrent_lanes() { return g_concurrent_lanes; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool make_tensor_map_f32(CUtensorMap *map, const float *plane, int width, int height, int box_w, int box_h) {
  static const EncodeTiledFn encode = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  memset(map, 0, sizeof *map);
  if (!encode || !plane || (width & 3) || (reinterpret_cast<uintptr_t>(plane) & 15) || box_w > 256 || box_h > 256 || ((box_w * 4) & 15)) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)height}, strides[1] = {(cuuint64_t)width * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}, estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(plane), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

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
void tdb_set_concurrency_hint(int lanes) { tdb::g_concurrent_lanes = lanes > 0 ? lanes : 0; }
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
