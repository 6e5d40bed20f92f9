// Event-bracketed duration of (nearly) empty kernels: what a persistent one-CTA-per-SM launch costs
// before it does any work, by dynamic shared memory size, launch kind and parameter size.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>
struct Big { char b[1024]; };
__global__ void __launch_bounds__(512, 1) k_empty(const __grid_constant__ Big p, int* out) {
  extern __shared__ char smem[];
  if (threadIdx.x == 0 && p.b[0] == 77) out[blockIdx.x] = smem[0];
}
__global__ void __launch_bounds__(512, 1) k_small(int x, int* out) {
  extern __shared__ char smem[];
  if (threadIdx.x == 0 && x == 77) out[blockIdx.x] = smem[0];
}
int main() {
  int* d; cudaMalloc(&d, 4096);
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  Big big{}; int x = 0;
  for (int coop = 0; coop < 2; ++coop)
    for (int bigp = 0; bigp < 2; ++bigp)
      for (size_t smem : {0ul, 48ul << 10, 139ul << 10, 200ul << 10}) {
        const void* fn = bigp ? (const void*)k_empty : (const void*)k_small;
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        std::vector<float> ms;
        for (int i = 0; i < 60; ++i) {
          void* a1[] = {&big, &d}; void* a2[] = {&x, &d};
          void** args = bigp ? a1 : a2;
          cudaEventRecord(e0, s);
          if (coop) cudaLaunchCooperativeKernel(fn, dim3(148), dim3(512), args, smem, s);
          else cudaLaunchKernel(fn, dim3(148), dim3(512), args, smem, s);
          cudaEventRecord(e1, s);
          cudaStreamSynchronize(s);
          float t; cudaEventElapsedTime(&t, e0, e1); if (i >= 10) ms.push_back(t);
        }
        std::sort(ms.begin(), ms.end());
        printf("coop=%d params=%s smem=%3zuKB  median %.2f us  min %.2f us  err=%d\n", coop, bigp ? "1KB" : "8B ", smem >> 10,
               ms[ms.size() / 2] * 1e3, ms[0] * 1e3, (int)cudaGetLastError());
      }
  return 0;
}
