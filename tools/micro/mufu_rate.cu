// Micro-benchmark (development aid): MUFU (ex2 / rcp) issue rate per SM sub-partition on sm_100a, alone and with 2 / 4 warps
// per scheduler.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_rate tools/micro/mufu_rate.cu && /tmp/mufu_rate
#include <cuda_runtime.h>
#include <stdio.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i * 0.01f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      else if (MODE == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      else if (MODE == 2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      else asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  const long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < 16; ++i) acc += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int threads : {32, 128, 256, 512, 1024}) {
    k<MODE><<<148, threads>>>(out, iters, cyc);
    k<MODE><<<148, threads>>>(out, iters, cyc);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_clk = (double)threads * 16 * iters / (double)c;
    printf("%s: %4d threads/SM  %8lld cycles  -> %.2f lane-ops/clk/SM  (%.1f clk per warp instruction per scheduler)\n", name, threads, c, per_clk,
           (double)c / ((double)((threads + 127) / 128 > 0 ? (threads < 128 ? 1 : threads / 128) : 1) * 16 * iters));
  }
}
int main() {
  run<0>("ex2"); run<1>("rcp"); run<2>("lg2"); run<3>("tanh");
  return 0;
}
