// The non-default tails of SpeechFeaturizer.call and its "waveform" feature type.
//
//   mfcc                 src/speech_featurizer.py:128-130: tf.signal.mfccs_from_log_mel_spectrograms =
//                        DCT-II (unnormalised, factor 2) of the log-mel vector * rsqrt(2*80);
//   normalize_zscore     :82-85  (x - mean) / sqrt(var + 1e-9), mean/variance over axis=1 of the per-utterance
//                        [T, 80] feature, i.e. over the 80 bins of each frame (population variance);
//   normalize_min_max    :86-91  (x - min) / (max - min) per frame; for feature_type "spectrogram" the minimum is
//                        the constant logarithm(output_floor);
//   waveform             :132-133 the normalised, pre-emphasised signal itself.
// One warp per valid frame, in place on [B, T_max, 80]; padded rows are not touched (they stay 0.0).  All of
// this is off in config/model.yaml, so it is a separate small launch after logmel_kernel rather than more
// code in the hot kernel.
#include "common.cuh"

using namespace tasr;

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowsPerCta = 64;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

__global__ void __launch_bounds__(kThreads)
feature_post_kernel(float* __restrict__ feat, const int32_t* __restrict__ n_frames, int32_t T_max,
                    const float* __restrict__ dct, int32_t zscore, int32_t minmax, int32_t fixed_min, float min_const) {
  extern __shared__ float sm[];
  float* D = sm;                      // [80][80] when dct != nullptr
  float* X = sm + (dct ? kMel * kMel : 0);   // [kWarps][80]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int T = min(n_frames[b], T_max);
  const int r0 = blockIdx.x * kRowsPerCta;
  if (r0 >= T) return;
  if (dct) {
    for (int i = tid; i < kMel * kMel; i += kThreads) D[i] = dct[i];
    __syncthreads();
  }
  float* xs = X + warp * kMel;
  for (int r = r0 + warp; r < min(r0 + kRowsPerCta, T); r += kWarps) {
    float* row = feat + ((size_t)b * T_max + r) * kMel;
    float v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = (lane + 32 * i < kMel) ? row[lane + 32 * i] : 0.0f;
    if (dct) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 3; ++i) if (lane + 32 * i < kMel) xs[lane + 32 * i] = v[i];
      __syncwarp();
      float acc[3] = {0.0f, 0.0f, 0.0f};
      for (int n = 0; n < kMel; ++n) {
        const float x = xs[n];
#pragma unroll
        for (int i = 0; i < 3; ++i) if (lane + 32 * i < kMel) acc[i] = fmaf(x, D[n * kMel + lane + 32 * i], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) v[i] = acc[i];
    }
    const bool has2 = lane + 64 < kMel;   // lanes 0..15 hold a third value
    if (zscore) {
      const float mean = warp_sum(v[0] + v[1] + (has2 ? v[2] : 0.0f)) * (1.0f / kMel);
      float d0 = v[0] - mean, d1 = v[1] - mean, d2 = has2 ? v[2] - mean : 0.0f;
      const float var = warp_sum(d0 * d0 + d1 * d1 + d2 * d2) * (1.0f / kMel);
      const float sd = sqrtf(var + 1e-9f);
      v[0] = d0 / sd; v[1] = d1 / sd; v[2] = d2 / sd;
    } else if (minmax) {
      const float mx = warp_max(fmaxf(fmaxf(v[0], v[1]), has2 ? v[2] : -INFINITY));
      const float mn = fixed_min ? min_const : warp_min(fminf(fminf(v[0], v[1]), has2 ? v[2] : INFINITY));
      const float den = mx - mn;
      v[0] = (v[0] - mn) / den; v[1] = (v[1] - mn) / den; v[2] = (v[2] - mn) / den;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) if (lane + 32 * i < kMel) row[lane + 32 * i] = v[i];
  }
}

__global__ void __launch_bounds__(256)
waveform_kernel(const float* __restrict__ wav, const int32_t* __restrict__ len, const float* __restrict__ peak,
                int64_t row_stride, int32_t normalize, float c, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int n = len[b];
  const float* row = wav + (size_t)b * row_stride;
  float* orow = out + (size_t)b * row_stride;
  float g = 1.0f;
  if (normalize) g = __fdiv_rn(1.0f, __fadd_rn(peak[b], 1e-9f));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = __fmul_rn(row[i], g);
    float y = x;
    if (c > 0.0f && i > 0) y = __fsub_rn(x, __fmul_rn(c, __fmul_rn(row[i - 1], g)));
    orow[i] = y;
  }
}

}  // namespace

int tasr_feature_post_launch(const TasrFeaturizer* f, float* feat, const int32_t* n_frames, int32_t B, int32_t T_max,
                             cudaStream_t st) {
  if (B == 0 || T_max == 0) return TASR_OK;
  if (B > 65535) return fail(TASR_ERR_UNSUPPORTED, "tasr_logmel_f32: batch > 65535 with mfcc / feature normalisation");
  const float* dct = (f->p.feature_type == TASR_FEAT_MFCC) ? f->d_dct : nullptr;
  const size_t smem = ((dct ? kMel * kMel : 0) + kWarps * kMel) * sizeof(float);
  // reference: min_value = self.logarithm(self.output_floor) for spectrograms (src/speech_featurizer.py:87-88)
  const int fixed_min = (f->p.feature_type == TASR_FEAT_SPECTROGRAM) ? 1 : 0;
  const float min_const = f->p.log_base_e ? logf(f->p.output_floor) : logf(f->p.output_floor) / logf(10.0f);
  dim3 grid((T_max + kRowsPerCta - 1) / kRowsPerCta, B);
  feature_post_kernel<<<grid, kThreads, smem, st>>>(feat, n_frames, T_max, dct, f->p.normalize_zscore ? 1 : 0,
                                                    f->p.normalize_min_max ? 1 : 0, fixed_min, min_const);
  TASR_LAUNCH_CHECK("feature_post_kernel");
  return TASR_OK;
}

namespace {
__global__ void apply_gain_kernel(float* __restrict__ feat, const int32_t* __restrict__ n_frames, int T_max, int F,
                                  const float* __restrict__ peak, float scale2, float floor_) {
  const int b = blockIdx.y;
  const int n = min(n_frames[b], T_max) * F;
  float lg;
  const float g = __fdiv_rn(1.0f, __fadd_rn(peak[b], 1e-9f));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(g));
  const float c = scale2 * lg;
  if (c != c) floor_ = c;     // a NaN peak (a NaN sample) makes the whole utterance NaN, like the reference; fmaxf alone would drop it
  float* row = feat + (size_t)b * T_max * F;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) row[i] = fmaxf(row[i] + c, floor_);
}
}  // namespace

extern "C" int tasr_apply_deferred_gain(float* feat, const int32_t* n_frames, int32_t B, int32_t T_max, int32_t F,
                                        const TasrDeferredGain* gain, tasr_stream_t stream) {
  if (!feat || !n_frames || !gain || !gain->peak) return fail(TASR_ERR_BAD_ARG, "tasr_apply_deferred_gain: null argument");
  if (B < 0 || T_max < 0 || F < 0) return fail(TASR_ERR_BAD_ARG, "tasr_apply_deferred_gain: negative size");
  if (B == 0 || T_max == 0 || F == 0) return TASR_OK;
  if (B > 65535) return fail(TASR_ERR_UNSUPPORTED, "tasr_apply_deferred_gain: batch > 65535");
  const long long per = ((long long)T_max * F + 255) / 256;
  dim3 grid((unsigned)(per > 64 ? 64 : per), (unsigned)B);
  apply_gain_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feat, n_frames, T_max, F, gain->peak, gain->log_scale_x2, gain->log_floor);
  TASR_LAUNCH_CHECK("apply_gain_kernel");
  return TASR_OK;
}

extern "C" int tasr_waveform_f32(const TasrFeaturizer* f, const float* wav, const int32_t* len, const float* peak,
                                 int32_t B, int64_t row_stride, float* out, tasr_stream_t stream) {
  if (!f || !wav || !len || !out) return fail(TASR_ERR_BAD_ARG, "tasr_waveform_f32: null argument");
  if (B < 0 || row_stride < 0) return fail(TASR_ERR_BAD_ARG, "tasr_waveform_f32: negative size");
  if (f->p.normalize_signal && !peak)
    return fail(TASR_ERR_BAD_ARG, "tasr_waveform_f32: normalize_signal is set but peak is NULL (run tasr_absmax_f32 first)");
  if (B == 0 || row_stride == 0) return TASR_OK;
  if (B > 65535) return fail(TASR_ERR_UNSUPPORTED, "tasr_waveform_f32: batch > 65535");
  const long long per = (row_stride + 255) / 256;
  dim3 grid((unsigned)(per > 256 ? 256 : per), (unsigned)B);
  waveform_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(wav, len, peak, row_stride, f->p.normalize_signal ? 1 : 0,
                                                          f->p.preemphasis, out);
  TASR_LAUNCH_CHECK("waveform_kernel");
  return TASR_OK;
}
