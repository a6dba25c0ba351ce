// SpecAugment on the device: frequency and time masking of the batched features.
//
// Replaces FreqMasking.augment / TimeMasking.augment (src/augmentations/specaugment.py:6-62), which the
// reference's loader applies per utterance on the CPU (src/dataset.py:172): the spectrogram is multiplied
// by a 0/1 mask that is 0 on the bins [f0, f0+f) of every frame, resp. on the frames [t0, t0+t).  Here the
// masks of the whole batch are applied in place on [B, T_max, F] in one launch.  The random draws stay on
// the host side of the ABI (telugu_asr_b200/augmentation.py mirrors the reference's distributions); this
// kernel is the deterministic part.  A masked value is x * 0.0f, not a stored 0.0f, so the sign of zero is
// what the reference's multiply produces (-0.0 for the negative log-mel values).
#include "common.cuh"

using namespace tasr;

namespace {

constexpr int kMaxMasks = 8;

__global__ void __launch_bounds__(256)
specaugment_kernel(float* __restrict__ feat, const int32_t* __restrict__ n_frames, int32_t T_max, int32_t F,
                   const int32_t* __restrict__ time_masks, int32_t n_time,
                   const int32_t* __restrict__ freq_masks, int32_t n_freq) {
  const int b = blockIdx.y;
  const int T = min(n_frames[b], T_max);
  __shared__ int32_t tm[2 * kMaxMasks], fm[2 * kMaxMasks];
  if (threadIdx.x < 2 * n_time) tm[threadIdx.x] = time_masks[(size_t)b * 2 * n_time + threadIdx.x];
  if (threadIdx.x < 2 * n_freq) fm[threadIdx.x] = freq_masks[(size_t)b * 2 * n_freq + threadIdx.x];
  __syncthreads();
  bool any = false;
  for (int i = 0; i < n_time; ++i) any |= tm[2 * i + 1] > 0;
  for (int i = 0; i < n_freq; ++i) any |= fm[2 * i + 1] > 0;
  if (!any) return;                                   // this utterance drew "no augmentation"
  float* base = feat + (size_t)b * T_max * F;
  const size_t n = (size_t)T * F;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / F), f = (int)(i - (size_t)t * F);
    bool hit = false;
    for (int m = 0; m < n_time; ++m) hit |= (t >= tm[2 * m] && t < tm[2 * m] + tm[2 * m + 1]);
    for (int m = 0; m < n_freq; ++m) hit |= (f >= fm[2 * m] && f < fm[2 * m] + fm[2 * m + 1]);
    if (hit) base[i] = __fmul_rn(base[i], 0.0f);
  }
}

}  // namespace

extern "C" int tasr_specaugment_f32(float* feat, const int32_t* n_frames, int32_t B, int32_t T_max, int32_t F,
                                    const int32_t* time_masks, int32_t n_time, const int32_t* freq_masks,
                                    int32_t n_freq, tasr_stream_t stream) {
  if (!feat || !n_frames) return fail(TASR_ERR_BAD_ARG, "tasr_specaugment_f32: null argument");
  if (B < 0 || T_max < 0 || F < 1 || n_time < 0 || n_freq < 0) return fail(TASR_ERR_BAD_ARG, "tasr_specaugment_f32: bad shape");
  if (n_time > kMaxMasks || n_freq > kMaxMasks)
    return fail(TASR_ERR_UNSUPPORTED, "tasr_specaugment_f32: at most %d masks of each kind per utterance", kMaxMasks);
  if ((n_time > 0 && !time_masks) || (n_freq > 0 && !freq_masks)) return fail(TASR_ERR_BAD_ARG, "tasr_specaugment_f32: null mask table");
  if (B == 0 || T_max == 0 || (n_time == 0 && n_freq == 0)) return TASR_OK;
  if (B > 65535) return fail(TASR_ERR_UNSUPPORTED, "tasr_specaugment_f32: batch > 65535");
  const long long per = ((long long)T_max * F + 255) / 256;
  dim3 grid((unsigned)(per > 64 ? 64 : per), (unsigned)B);
  specaugment_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feat, n_frames, T_max, F, time_masks, n_time, freq_masks, n_freq);
  TASR_LAUNCH_CHECK("specaugment_kernel");
  return TASR_OK;
}
