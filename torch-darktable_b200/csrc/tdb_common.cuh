// Shared host/device helpers of libtdb200 (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <atomic>
#include <mutex>

#include "../../include/tdb200.h"

namespace tdb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of this

// MUFU.RCP: what a / b compiles to under --use_fast_math is a * rcp_approx(b); written out where the reciprocal of a loop
// invariant should be taken once (bit-identical to the division it replaces)
__device__ __forceinline__ float rcp_approx(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// ---- error plumbing -------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int check_launch(const char *what);  // cudaGetLastError() -> TDB_OK / TDB_ECUDA, bumps the launch counter
// Kernel attributes (cudaFuncSetAttribute: dynamic shared memory above 48 KB) belong to a device's context: they are set on the
// first launch per DEVICE, not per process, so that one process may drive several GPUs -- and from several host threads (ctypes
// releases the GIL; pipeline/tiled.py runs one thread per band).  The device's bit is published only AFTER the attributes have been
// applied, under a mutex, so that no thread can launch a kernel whose opt-in is still pending.
class DeviceOnce {
 public:
  template <class F>
  void run(F &&apply) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (done_.load(std::memory_order_acquire) & bit) return;
    std::lock_guard<std::mutex> lock(mutex_);
    if (done_.load(std::memory_order_relaxed) & bit) return;
    apply();
    done_.fetch_or(bit, std::memory_order_release);
  }

 private:
  std::atomic<unsigned long long> done_{0};
  std::mutex mutex_;
};
void count_launches(int n);
// > 1 while the calling thread issues work for several frames in flight on different streams (tdb_set_concurrency_hint)
int concurrent_lanes();
// Tensor map of a row-major float32 (height, width) plane for boxes of box_w x box_h elements (cuTensorMapEncodeTiled, reached through
// the runtime's driver entry point query: libtdb200 does not link libcuda).  Needs a 16-byte aligned base and width % 4 == 0; returns
// false when the plane does not qualify or the driver call is unavailable -- callers then stage with ordinary loads.
bool make_tensor_map_f32(CUtensorMap *map, const float *plane, int width, int height, int box_w, int box_h);
// a per-device side stream ordered after everything already queued on `main` (returns `main` itself when unavailable)
cudaStream_t fork_side(cudaStream_t main);
void join_side(cudaStream_t main, cudaStream_t side);

#define TDB_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::tdb::set_error(__VA_ARGS__);  \
      return TDB_EINVAL;              \
    }                                 \
  } while (0)

void note_stream(cudaStream_t s);  // remembers the stream of the entry point for the optional per-kernel timing hook
static inline cudaStream_t as_stream(tdb_stream_t s) {
  cudaStream_t cs = reinterpret_cast<cudaStream_t>(s);
  note_stream(cs);
  return cs;
}
static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- CFA helpers ------------------------------------------------------------------------------------------
// colour of a CFA site: 0 = R, 1 = G, 2 = B (the four supported filter words never yield 3)
__host__ __device__ __forceinline__ int fc(int row, int col, uint32_t filters) {
  return (filters >> ((((row << 1) & 14) + (col & 1)) << 1)) & 3u;
}

// ---- packed 12-bit pairs ---------------------------------------------------------------------------------
// three bytes b0,b1,b2 (little end of `w`) -> two 12-bit samples
template <bool kIds>
__device__ __forceinline__ void unpack_pair(uint32_t w, uint32_t &p0, uint32_t &p1) {
  const uint32_t b0 = w & 0xffu, b1 = (w >> 8) & 0xffu, b2 = (w >> 16) & 0xffu;
  if (kIds) {
    p0 = (b0 << 4) | (b2 & 0xfu);
    p1 = (b1 << 4) | (b2 >> 4);
  } else {
    p0 = ((b1 & 0xfu) << 8) | b0;
    p1 = (b2 << 4) | (b1 >> 4);
  }
}

template <bool kIds>
__device__ __forceinline__ uint32_t pack_pair(uint32_t p0, uint32_t p1) {
  if (kIds) return (p0 >> 4) | ((p1 >> 4) << 8) | ((((p0 & 0xfu) << 4) | (p1 & 0xfu)) << 16);
  return (p0 & 0xffu) | ((((p1 & 0xfu) << 4) | (p0 >> 8)) << 8) | ((p1 >> 4) << 16);
}

// sample `i` (0-based) of a packed stream through byte loads; used on halo / unaligned paths only
template <bool kIds>
__device__ __forceinline__ uint32_t packed_sample(const uint8_t *__restrict__ p, int64_t i) {
  const uint8_t *b = p + (i >> 1) * 3;
  const uint32_t w = (uint32_t)__ldg(b) | ((uint32_t)__ldg(b + 1) << 8) | ((uint32_t)__ldg(b + 2) << 16);
  uint32_t p0, p1;
  unpack_pair<kIds>(w, p0, p1);
  return (i & 1) ? p1 : p0;
}

// ---- small math ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
__device__ __forceinline__ float sqr(float x) { return x * x; }
__device__ __forceinline__ float mixf(float a, float b, float t) { return (1.0f - t) * a + t * b; }

// streaming 128-bit accesses that do not pollute L1 (data is touched once)
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4 *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ void st_stream(uint4 *p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

// ---- bulk asynchronous copies (the TMA engine's 1-D form: cp.async.bulk, SASS UBLKCP) with mbarrier completion ---------------
// One thread arms the barrier with the byte count, any threads issue row copies (16-byte aligned, multiples of 16 bytes), every
// thread waits on the barrier's phase; the data is then visible to the generic proxy without further fences.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
// 2-D tensor-map load (cp.async.bulk.tensor.2d, SASS UTMALDG): the box whose first element has coordinates (c0, c1) = (column, row)
// lands densely in shared memory (row pitch = box width; destination 128-byte aligned); elements outside the tensor arrive as zeros,
// which is what the zero-filled halos of the reference's demosaic kernels hold (csrc/debayer/ppg.cu:61,159,270)
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_addr(dst_smem)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(smem_addr(bar))
               : "memory");
}
// shared -> global: the writers of the shared-memory tile call bulk_store_fence() before the barrier that precedes the copy
__device__ __forceinline__ void bulk_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_addr(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_commit_and_drain() {  // until the shared-memory source may be overwritten / released
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(smem_addr(bar)),
      "r"(phase)
      : "memory");
}

// float atomic min / max through the ordered-int trick (no CAS loop)
__device__ __forceinline__ void atomic_min_float(float *addr, float v) {
  if (v >= 0.0f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace tdb
