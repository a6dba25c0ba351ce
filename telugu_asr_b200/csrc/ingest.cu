// Device-side collate of raw audio: ragged, packed host batches -> the padded
// [B, row_stride] float32 layout the featurizer kernels read.
//
// Replaces the waveform half of the reference's loader: tf.audio.decode_wav's int16 -> float32
// conversion (src/utils/data_util.py:31: sample / 32768, exact in float32) and the per-utterance
// hand-off of src/dataset.py:167-175.  The host ships only the valid samples (no padding crosses
// PCIe), optionally still as 16-bit PCM (half the bytes again); offsets are in samples and must be
// multiples of 8 so that every utterance starts 16-byte aligned in either format.
// Samples beyond len[b] are left untouched: no kernel of this library reads them.
#include "common.cuh"

using namespace tasr;

namespace {

constexpr int kThreads = 256;
constexpr int kPerCta = kThreads * 8 * 4;   // samples per CTA: 8 x 128-bit stores per thread

__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ldg_stream2(const uint2* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float pcm(int v) { return __fmul_rn((float)v, 1.0f / 32768.0f); }

template <bool PCM16>
__global__ void __launch_bounds__(kThreads)
unpack_kernel(const void* __restrict__ packed, const int64_t* __restrict__ offset, const int32_t* __restrict__ len,
              float* __restrict__ wav, int64_t row_stride) {
  const int b = blockIdx.y;
  const int n = min(len[b], (int)min(row_stride, (int64_t)0x7fffffff));
  const int start = blockIdx.x * kPerCta;
  if (start >= n) return;
  const int64_t off = offset[b];
  float* dst = wav + (size_t)b * row_stride;
  const int end = min(n, start + kPerCta);
  const int nvec = (end - start) >> 2;
  float4 v[8];
  if (PCM16) {
    const uint2* src = reinterpret_cast<const uint2*>(reinterpret_cast<const int16_t*>(packed) + off + start);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = threadIdx.x + i * kThreads;
      uint2 r = make_uint2(0u, 0u);
      if (idx < nvec) r = ldg_stream2(src + idx);
      v[i] = make_float4(pcm((short)(r.x & 0xffffu)), pcm((short)(r.x >> 16)), pcm((short)(r.y & 0xffffu)), pcm((short)(r.y >> 16)));
    }
  } else {
    const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(packed) + off + start);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = threadIdx.x + i * kThreads;
      v[i] = (idx < nvec) ? ldg_stream4(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = threadIdx.x + i * kThreads;
    if (idx < nvec) *reinterpret_cast<float4*>(dst + start + 4 * idx) = v[i];
  }
  for (int i = start + (nvec << 2) + threadIdx.x; i < end; i += kThreads) {   // < 4 tail samples
    dst[i] = PCM16 ? pcm(reinterpret_cast<const int16_t*>(packed)[off + i]) : reinterpret_cast<const float*>(packed)[off + i];
  }
}

}  // namespace

static int unpack_common(const char* who, bool pcm16, const void* packed, const int64_t* offset, const int32_t* len,
                         int32_t B, int32_t max_len, float* wav, int64_t row_stride, tasr_stream_t stream) {
  if (!packed || !offset || !len || !wav) return fail(TASR_ERR_BAD_ARG, "%s: null argument", who);
  if (B < 0 || max_len < 0 || row_stride < 0) return fail(TASR_ERR_BAD_ARG, "%s: negative size", who);
  if (max_len > row_stride) return fail(TASR_ERR_BAD_ARG, "%s: max_len=%d exceeds row_stride=%lld", who, max_len, (long long)row_stride);
  if (!aligned16(packed) || !aligned16(wav) || (row_stride & 3))
    return fail(TASR_ERR_MISALIGNED, "%s: packed/wav must be 16-byte aligned and row_stride a multiple of 4 samples", who);
  if (B == 0 || max_len == 0) return TASR_OK;
  if (B > 65535) return fail(TASR_ERR_UNSUPPORTED, "%s: batch > 65535", who);
  dim3 grid((unsigned)((max_len + kPerCta - 1) / kPerCta), (unsigned)B);
  if (pcm16) unpack_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(packed, offset, len, wav, row_stride);
  else unpack_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(packed, offset, len, wav, row_stride);
  TASR_LAUNCH_CHECK("unpack_kernel");
  return TASR_OK;
}

extern "C" int tasr_unpack_f32(const float* packed, const int64_t* offset, const int32_t* len, int32_t B,
                               int32_t max_len, float* wav, int64_t row_stride, tasr_stream_t stream) {
  return unpack_common("tasr_unpack_f32", false, packed, offset, len, B, max_len, wav, row_stride, stream);
}

extern "C" int tasr_unpack_pcm16(const int16_t* packed, const int64_t* offset, const int32_t* len, int32_t B,
                                 int32_t max_len, float* wav, int64_t row_stride, tasr_stream_t stream) {
  return unpack_common("tasr_unpack_pcm16", true, packed, offset, len, B, max_len, wav, row_stride, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// Valid rows only on the way back (the other end of the device collate): x [B, T, C] with a valid prefix of len[b] rows per
// utterance -> packed [sum_b len[b], C], utterance b at row offset[b] = sum_{b' < b} len[b'] (offset[B] = the total).  Nearly
// half of the padded [B, T3, 192] encoder input of a ragged batch is collate padding that no consumer reads; a D2H copy of
// the packed rows halves the bytes on the host link.  One CTA per (utterance, 64-row slab); the offset is a block-wide sum.
namespace {
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ len, int B, int T, int C,
                                                        float* __restrict__ packed, int64_t* __restrict__ offset) {
  __shared__ long long part[8];
  __shared__ long long base_s;
  const int b = blockIdx.y, tid = threadIdx.x;
  long long s = 0;
  for (int i = tid; i < b; i += 256) s += max(0, min(len[i], T));
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((tid & 31) == 0) part[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    long long t = 0;
    for (int w = 0; w < 8; ++w) t += part[w];
    base_s = t;
    if (blockIdx.x == 0) {
      offset[b] = t;
      if (b == B - 1) offset[B] = t + max(0, min(len[b], T));
    }
  }
  __syncthreads();
  const int L = max(0, min(len[b], T));
  const int r0 = blockIdx.x * 64, r1 = min(L, r0 + 64);
  if (r0 >= L) return;
  const int c4 = C >> 2;
  const float4* src = reinterpret_cast<const float4*>(x + ((size_t)b * T + r0) * C);
  float4* dst = reinterpret_cast<float4*>(packed + ((size_t)base_s + r0) * C);
  const int n4 = (r1 - r0) * c4;
  for (int i = tid; i < n4; i += 256) dst[i] = __ldg(src + i);
}
}  // namespace

extern "C" int tasr_pack_valid_rows(const float* x, const int32_t* len, int32_t B, int32_t T, int32_t C, float* packed,
                                    int64_t* offset, tasr_stream_t stream) {
  if (!x || !len || !packed || !offset) return fail(TASR_ERR_BAD_ARG, "tasr_pack_valid_rows: null argument");
  if (B < 0 || T < 0 || C < 1) return fail(TASR_ERR_BAD_ARG, "tasr_pack_valid_rows: bad shape");
  if ((C & 3) || !aligned16(x) || !aligned16(packed))
    return fail(TASR_ERR_MISALIGNED, "tasr_pack_valid_rows: C must be a multiple of 4 and x / packed 16-byte aligned");
  if (B > 65535) return fail(TASR_ERR_UNSUPPORTED, "tasr_pack_valid_rows: batch > 65535");
  if (B == 0) return TASR_OK;
  dim3 grid((unsigned)(T > 0 ? (T + 63) / 64 : 1), (unsigned)B);
  pack_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, len, B, T, C, packed, offset);
  TASR_LAUNCH_CHECK("pack_rows_kernel");
  return TASR_OK;
}
