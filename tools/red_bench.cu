// Micro-benchmark: throughput of red.global.add.f32 (scalar / v2 / v4) and plain stores on an L2-resident accumulator,
// with the access pattern of the Wiener overlap-add (a warp adds 32 consecutive floats of a row, 32 rows per tile).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/red_bench tools/red_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__device__ __forceinline__ void red(float *p, float v);
template <> __device__ __forceinline__ void red<1>(float *p, float v) { atomicAdd(p, v); }
template <> __device__ __forceinline__ void red<2>(float *p, float v) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%1};" ::"l"(p), "f"(v) : "memory");
}
template <> __device__ __forceinline__ void red<4>(float *p, float v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(v) : "memory");
}
template <> __device__ __forceinline__ void red<0>(float *p, float v) { *p = v; }

// every warp: tiles of 32 rows x (32*max(V,1)) floats, tile origins advance by 8 columns (overlap 4)
template <int V>
__global__ void k(float *acc, int width, int height, int tiles_x, int tiles_y) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  constexpr int W = V == 0 ? 1 : V;
  for (int t = warp; t < tiles_x * tiles_y; t += nwarps) {
    const int tx = t % tiles_x, ty = t / tiles_x;
    const int ox = tx * 8 * W, oy = ty * 8;
#pragma unroll 8
    for (int r = 0; r < 32; r++) {
      const int y = oy + r, x = ox + lane * W;
      if (y < height && x + W <= width) red<V>(acc + (size_t)y * width + x, 1.0f);
    }
  }
}

template <int V>
void run(const char *name, float *acc, int width, int height) {
  constexpr int W = V == 0 ? 1 : V;
  const int tiles_x = (width - 32 * W) / (8 * W) + 1, tiles_y = (height - 32) / 8 + 1;
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  for (int it = 0; it < 3; it++) k<V><<<148 * 8, 256>>>(acc, width, height, tiles_x, tiles_y);
  cudaEventRecord(a);
  for (int it = 0; it < 10; it++) k<V><<<148 * 8, 256>>>(acc, width, height, tiles_x, tiles_y);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 10;
  const double lanes = (double)tiles_x * tiles_y * 32 * 32;
  printf("%-10s %8.4f ms  %7.2f G lane-requests/s  %7.2f G floats/s  (%.3f cyc/lane-request/SM @1.9GHz)\n", name, ms, lanes / ms / 1e6,
         lanes * W / ms / 1e6, ms * 1e-3 * 1.9e9 * 148 / lanes);
}

int main() {
  const int width = 3840, height = 2160;
  float *acc;
  cudaMalloc(&acc, (size_t)width * height * 4);
  cudaMemset(acc, 0, (size_t)width * height * 4);
  run<0>("st.f32", acc, width, height);
  run<1>("red.f32", acc, width, height);
  run<2>("red.v2", acc, width, height);
  run<4>("red.v4", acc, width, height);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
