// Log-mel / spectrogram featurizer for frame geometries other than the specialised 400 / 160 / 512 / 80
// (SpeechFeaturizer derives frame_length = round(sample_rate * frame_ms / 1000) and frame_step from arbitrary
// constructor arguments, src/speech_featurizer.py:46-49; tf.signal.stft takes fft_length = the enclosing power of two).
// Same contract as logmel_kernel (normalize_signal -> preemphasis_signal -> stft -> |X|^2 -> mel matmul -> log, zero rows
// beyond n_frames[b]); a plain, general kernel — one CTA per frame, radix-2 FFT in shared memory, dense mel product — for
// configurations off the hot path.  No mfcc / per-frame normalisation here (feature_post.cu is written for 80 bins).
#include "logmel_common.cuh"
#include <math.h>
#include <vector>

using namespace tasr;

namespace {

constexpr int kGThreads = 128;

struct GenArgs {
  const float* wav;
  const int32_t* len;
  const float* peak;
  const float* win;       // [frame_length]
  const float* mel;       // [bins, n_mel] row-major
  const float2* tw;       // [fft/2] exp(-2 pi i j / fft)
  float* out;             // [B, T_max, n_mel]
  int32_t* n_frames;
  int64_t row_stride;
  int32_t B, T_max, frame_length, frame_step, fft, log2fft, n_mel, normalize, pad_end, mode;
  float preemph, floor_, log_scale;
};

__device__ __forceinline__ int gen_frames_of(int n, const GenArgs& a) {   // src/speech_featurizer.py:163-166
  const int Tb = a.pad_end ? (n > 0 ? (n + a.frame_step - 1) / a.frame_step : 0)
                           : ((n >= a.frame_length) ? 1 + (n - a.frame_length) / a.frame_step : 0);
  return min(Tb, a.T_max);
}

__global__ void __launch_bounds__(kGThreads) logmel_generic_kernel(const GenArgs a) {
  extern __shared__ __align__(16) float gsm[];
  float* re = gsm;
  float* im = gsm + a.fft;
  const int b = blockIdx.x / a.T_max, t = blockIdx.x - b * a.T_max, tid = threadIdx.x;
  const int n = a.len[b];
  const int Tb = gen_frames_of(n, a);
  if (t == 0 && tid == 0) a.n_frames[b] = Tb;
  float* orow = a.out + ((size_t)b * a.T_max + t) * a.n_mel;
  if (t >= Tb) {                                   // collate padding (src/dataset.py:241)
    for (int m = tid; m < a.n_mel; m += kGThreads) orow[m] = 0.0f;
    return;
  }
  const float* row = a.wav + (size_t)b * a.row_stride;
  float g = 1.0f;
  if (a.normalize) g = __fdiv_rn(1.0f, __fadd_rn(a.peak[b], 1e-9f));          // :70
  const float floor_b = (g != g) ? g : a.floor_;                                // NaN peak -> NaN features, as in the reference
  const int s0 = t * a.frame_step;
  const int shift = 32 - a.log2fft;
  for (int i = tid; i < a.fft; i += kGThreads) {
    float v = 0.0f;
    const int s = s0 + i;
    if (i < a.frame_length && s < n) {                                          // (pad_end: zeros appended after pre-emphasis)
      const float x = __fmul_rn(row[s], g);                                     // :71
      v = x;
      if (a.preemph > 0.0f && s > 0) v = __fsub_rn(x, __fmul_rn(a.preemph, __fmul_rn(row[s - 1], g)));   // :77-79
      v *= a.win[i];
    }
    const int j = (int)(__brev((unsigned)i) >> shift);
    re[j] = v;
    im[j] = 0.0f;
  }
  __syncthreads();
  for (int st = 1; st <= a.log2fft; ++st) {
    const int half = 1 << (st - 1);
    const int tstep = a.fft >> st;
    for (int j = tid; j < a.fft / 2; j += kGThreads) {
      const int k = j & (half - 1);
      const int i0 = ((j >> (st - 1)) << st) + k, i1 = i0 + half;
      const float2 w = a.tw[k * tstep];
      const float xr = re[i1], xi = im[i1];
      const float tr = xr * w.x - xi * w.y, ti = xr * w.y + xi * w.x;
      const float ur = re[i0], ui = im[i0];
      re[i0] = ur + tr; im[i0] = ui + ti;
      re[i1] = ur - tr; im[i1] = ui - ti;
    }
    __syncthreads();
  }
  const int bins = a.fft / 2 + 1;
  for (int k = tid; k < bins; k += kGThreads) {
    const float xr = re[k], xi = im[k];
    im[k] = xr * xr + xi * xi;                     // power spectrum (distinct array from the reads of other threads: re untouched)
  }
  __syncthreads();
  const float* P = im;
  for (int m = tid; m < a.n_mel; m += kGThreads) {
    float acc;
    if (a.mode == 1) {                             // "spectrogram": log power of the first n_mel FFT bins (:124-126)
      acc = P[m];
    } else {
      acc = 0.0f;
      for (int k = 0; k < bins; ++k) acc = fmaf(P[k], a.mel[(size_t)k * a.n_mel + m], acc);
    }
    orow[m] = __log2f(fmaxf(acc, floor_b)) * a.log_scale;
  }
}

}  // namespace

int tasr_logmel_generic_create(TasrFeaturizer* f, const float* hann_host, const float* mel_w_host) {
  const TasrFeatParams& p = f->p;
  int lg = 0;
  while ((1 << lg) < p.fft_length) ++lg;
  if ((1 << lg) != p.fft_length || p.fft_length < p.frame_length || p.fft_length > 4096 || p.fft_length < 8)
    return fail(TASR_ERR_UNSUPPORTED, "tasr_featurizer_create: fft_length must be a power of two in [8, 4096] and >= frame_length; got %d / %d",
                p.fft_length, p.frame_length);
  if (p.frame_length < 1 || p.frame_step < 1 || p.num_mel_bins < 1 || p.num_mel_bins > 1024)
    return fail(TASR_ERR_BAD_ARG, "tasr_featurizer_create: bad frame geometry %d / %d / %d", p.frame_length, p.frame_step, p.num_mel_bins);
  if (p.feature_type == TASR_FEAT_MFCC || p.normalize_zscore || p.normalize_min_max)
    return fail(TASR_ERR_UNSUPPORTED, "tasr_featurizer_create: mfcc and per-frame normalisation are built for the 400/160/512/80 geometry only");
  if (p.feature_type == TASR_FEAT_SPECTROGRAM && p.num_mel_bins > p.fft_length / 2 + 1)
    return fail(TASR_ERR_BAD_ARG, "tasr_featurizer_create: spectrogram needs num_feature_bins <= fft_length/2+1");
  const int bins = p.fft_length / 2 + 1;
  std::vector<float2> tw((size_t)p.fft_length / 2);
  const double two_pi = 6.283185307179586476925286766559;
  for (int j = 0; j < p.fft_length / 2; ++j) tw[j] = make_float2((float)cos(two_pi * j / p.fft_length), (float)(-sin(two_pi * j / p.fft_length)));
  int rc = check_cuda(cudaMalloc(&f->d_gwin, (size_t)p.frame_length * sizeof(float)), "cudaMalloc window");
  if (rc == TASR_OK) rc = check_cuda(cudaMalloc(&f->d_gmel, (size_t)bins * p.num_mel_bins * sizeof(float)), "cudaMalloc mel");
  if (rc == TASR_OK) rc = check_cuda(cudaMalloc(&f->d_gtw, tw.size() * sizeof(float2)), "cudaMalloc twiddles");
  if (rc == TASR_OK) rc = check_cuda(cudaMemcpy(f->d_gwin, hann_host, (size_t)p.frame_length * sizeof(float), cudaMemcpyHostToDevice), "copy window");
  if (rc == TASR_OK) rc = check_cuda(cudaMemcpy(f->d_gmel, mel_w_host, (size_t)bins * p.num_mel_bins * sizeof(float), cudaMemcpyHostToDevice), "copy mel");
  if (rc == TASR_OK) rc = check_cuda(cudaMemcpy(f->d_gtw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice), "copy twiddles");
  f->generic = 1;
  f->g_log2fft = lg;
  return rc;
}

int tasr_logmel_generic_launch(const TasrFeaturizer* f, const float* wav, const int32_t* len, const float* peak, int32_t B,
                               int64_t row_stride, float* out, int32_t T_max, int32_t* n_frames, cudaStream_t st) {
  if (B == 0) return TASR_OK;
  if (T_max == 0) return check_cuda(cudaMemsetAsync(n_frames, 0, (size_t)B * sizeof(int32_t), st), "cudaMemsetAsync n_frames");
  if ((long long)B * T_max > 0x7fffffffLL) return fail(TASR_ERR_UNSUPPORTED, "tasr_logmel_f32: too many frames");
  GenArgs a;
  a.wav = wav; a.len = len; a.peak = peak; a.win = f->d_gwin; a.mel = f->d_gmel; a.tw = f->d_gtw; a.out = out; a.n_frames = n_frames;
  a.row_stride = row_stride; a.B = B; a.T_max = T_max;
  a.frame_length = f->p.frame_length; a.frame_step = f->p.frame_step; a.fft = f->p.fft_length; a.log2fft = f->g_log2fft;
  a.n_mel = f->p.num_mel_bins; a.normalize = f->p.normalize_signal ? 1 : 0; a.pad_end = f->p.pad_end ? 1 : 0;
  a.mode = (f->p.feature_type == TASR_FEAT_SPECTROGRAM) ? 1 : 0;
  a.preemph = f->p.preemphasis; a.floor_ = f->p.output_floor; a.log_scale = f->log_scale;
  logmel_generic_kernel<<<(unsigned)((long long)B * T_max), kGThreads, (size_t)2 * a.fft * sizeof(float), st>>>(a);
  TASR_LAUNCH_CHECK("logmel_generic_kernel");
  return TASR_OK;
}
