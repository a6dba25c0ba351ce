// Micro-benchmark (development aid): issue rate of scalar FFMA / FADD against the packed FFMA2 / FADD2 of sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/f32x2_rate tools/micro/f32x2_rate.cu && /tmp/f32x2_rate
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float a[16];
  u64 p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
  const u64 ps = pk(s, s * 0.5f);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 0.25f);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(ps));
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = a[i] + s;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
    }
  }
  float acc = 0;
  for (int i = 0; i < 16; ++i) acc += a[i];
  for (int i = 0; i < 8; ++i) { float x, y; upk(p[i], x, y); acc += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE>
void run(const char* name, int per_iter_flops_per_thread) {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  const int iters = 20000;
  k<MODE><<<148 * 8, 256>>>(d, 100, 1.0001f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 8 * 256 * iters * per_iter_flops_per_thread;
  printf("%-8s %8.3f ms  %7.2f T elem-ops/s (%s)\n", name, ms, ops / ms / 1e9, "one FMA or ADD on one float = 1");
  cudaFree(d);
}
int main() {
  run<0>("FFMA", 16); run<1>("FFMA2", 16); run<2>("FADD", 16); run<3>("FADD2", 16);
  return 0;
}
