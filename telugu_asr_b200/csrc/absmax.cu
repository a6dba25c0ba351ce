// Per-utterance peak |x| (src/speech_featurizer.py:70: tf.reduce_max(tf.abs(signal))).
// HBM-bound streaming reduction: 128-bit coalesced loads, warp-shuffle + shared reduction,
// one atomicMax per CTA on the int view of the (non-negative) float.
#include "common.cuh"

using namespace tasr;

namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 8;                           // 8 x float4 in flight per thread
constexpr int kChunk = kThreads * kVecPerThread * 4;       // 8192 samples per CTA

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(kThreads)
absmax_kernel(const float* __restrict__ wav, const int32_t* __restrict__ len, int64_t row_stride,
              float* __restrict__ peak) {
  const int b = blockIdx.y;
  const int n = len[b];
  const int start = blockIdx.x * kChunk;
  if (start >= n) return;
  const float* row = wav + (size_t)b * row_stride;
  const int end = min(n, start + kChunk);
  // The maximum is taken over the BIT PATTERNS of |x| as unsigned integers: they order like the values for finite
  // samples and put a NaN (0x7fc00000 and up) above +inf, so a NaN sample makes the peak NaN exactly like
  // tf.reduce_max(tf.abs(x)) (src/speech_featurizer.py:70) - fmaxf would silently drop it.
  unsigned m = 0u;
  // row is 16-byte aligned and start is a multiple of 4 -> float4 loads; tail handled scalar.
  const int nvec = (end - start) >> 2;
  const float4* v4 = reinterpret_cast<const float4*>(row + start);
  float4 v[kVecPerThread];
#pragma unroll
  for (int i = 0; i < kVecPerThread; ++i) {
    int idx = threadIdx.x + i * kThreads;
    v[i] = (idx < nvec) ? ldg_stream(v4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < kVecPerThread; ++i)
    m = max(m, max(max(abs_bits(v[i].x), abs_bits(v[i].y)), max(abs_bits(v[i].z), abs_bits(v[i].w))));
  for (int i = start + (nvec << 2) + threadIdx.x; i < end; i += kThreads) m = max(m, abs_bits(row[i]));
  m = __reduce_max_sync(0xffffffffu, m);
  __shared__ unsigned sm[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = (threadIdx.x < kThreads / 32) ? sm[threadIdx.x] : 0u;
    m = __reduce_max_sync(0xffffffffu, m);
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned*>(peak + b), m);
  }
}

}  // namespace

extern "C" int tasr_absmax_f32(const float* wav, const int32_t* len, int32_t B, int64_t row_stride,
                               float* peak, tasr_stream_t stream) {
  if (!wav || !len || !peak) return fail(TASR_ERR_BAD_ARG, "tasr_absmax_f32: null argument");
  if (B < 0 || row_stride < 0) return fail(TASR_ERR_BAD_ARG, "tasr_absmax_f32: negative size");
  if (!aligned16(wav) || (row_stride & 3))
    return fail(TASR_ERR_MISALIGNED, "tasr_absmax_f32: wav must be 16-byte aligned and row_stride a multiple of 4 samples");
  if (B == 0) return TASR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  TASR_CUDA(cudaMemsetAsync(peak, 0, (size_t)B * sizeof(float), st));
  if (row_stride == 0) return TASR_OK;
  dim3 grid((unsigned)((row_stride + kChunk - 1) / kChunk), (unsigned)B);
  if (grid.y > 65535) return fail(TASR_ERR_UNSUPPORTED, "tasr_absmax_f32: batch > 65535");
  absmax_kernel<<<grid, kThreads, 0, st>>>(wav, len, row_stride, peak);
  TASR_LAUNCH_CHECK("absmax_kernel");
  return TASR_OK;
}
