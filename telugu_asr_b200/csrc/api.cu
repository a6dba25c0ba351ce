// tasr C-ABI: error plumbing, featurizer handle, length / mask kernels.
#include "common.cuh"

#include <math.h>
#include <string.h>
#include <atomic>
#include <vector>

namespace tasr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace tasr

using namespace tasr;

extern "C" int tasr_version(void) { return 100; }
extern "C" const char* tasr_last_error(void) { return tasr::g_err; }
extern "C" int64_t tasr_launch_count(void) { return (int64_t)tasr::launches(); }

// ---------------------------------------------------------------------------------------
// Featurizer handle
// ---------------------------------------------------------------------------------------
extern "C" int tasr_featurizer_create(const TasrFeatParams* p, const float* hann_host,
                                      const float* mel_w_host, TasrFeaturizer** out) {
  if (!p || !hann_host || !mel_w_host || !out) return fail(TASR_ERR_BAD_ARG, "tasr_featurizer_create: null argument");
  *out = nullptr;
  if (p->feature_type < TASR_FEAT_LOG_MEL || p->feature_type > TASR_FEAT_WAVEFORM)
    return fail(TASR_ERR_BAD_ARG, "tasr_featurizer_create: unknown feature_type %d", p->feature_type);
  if (!(p->output_floor > 0.0f)) return fail(TASR_ERR_BAD_ARG, "tasr_featurizer_create: output_floor must be > 0");
  if (p->output_floor < 1.17549435e-38f)
    return fail(TASR_ERR_UNSUPPORTED, "tasr_featurizer_create: output_floor below FLT_MIN (the log uses a flush-to-zero MUFU)");
  if (p->frame_length != kFrameLen || p->frame_step != kFrameStep || p->fft_length != kFft || p->num_mel_bins != kMel) {
    // Off the specialised geometry (config/model.yaml: 25 ms / 10 ms at 16 kHz, 80 bins): the general kernel of logmel_generic.cu
    TasrFeaturizer* f = new TasrFeaturizer();
    memset(f, 0, sizeof(*f));
    f->p = *p;
    f->log_scale = p->log_base_e ? 0.69314718055994530942f : 0.30102999566398119521f;
    int rc = tasr_logmel_generic_create(f, hann_host, mel_w_host);      // validates the geometry before touching the device
    if (rc == TASR_OK) rc = check_cuda(cudaGetDevice(&f->device), "cudaGetDevice");
    if (rc != TASR_OK) { tasr_featurizer_destroy(f); return rc; }
    *out = f;
    return TASR_OK;
  }

  // Banded mel structure from the dense matrix the caller built (values are used verbatim).
  MelBands bands;
  std::vector<float> bw;
  for (int m = 0; m < kMel; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < kBins; ++k) {
      float w = mel_w_host[k * kMel + m];
      if (!(w == w)) return fail(TASR_ERR_BAD_ARG, "tasr_featurizer_create: NaN in mel matrix");
      if (w != 0.0f) { if (lo < 0) lo = k; hi = k; }
    }
    if (lo < 0) { lo = 0; hi = -1; }
    int n = hi - lo + 1;
    int n4 = (n + 3) / 4;
    // Zero-weight padding reads P[k] up to k = lo+4*n4-1; the kernel's P rows are kBins+4 wide
    // and zero filled beyond kBins-1, so shift the window down only if it would run past that.
    if (lo + 4 * n4 > kBins + 4) lo = kBins + 4 - 4 * n4;
    bands.k0[m] = lo;
    bands.n4[m] = n4;
    bands.off4[m] = (int)(bw.size() / 4);
    for (int i = 0; i < 4 * n4; ++i) {
      int k = lo + i;
      bw.push_back((k < kBins) ? mel_w_host[k * kMel + m] : 0.0f);
    }
  }
  bands.total4 = (int)(bw.size() / 4);
  if (bands.total4 > kMelBandMaxW4)
    return fail(TASR_ERR_UNSUPPORTED,
                "tasr_featurizer_create: mel matrix is not banded enough (%d float4 > %d); only "
                "triangular filterbanks are supported", bands.total4, kMelBandMaxW4);
  bw.resize((size_t)kMelBandMaxW4 * 4, 0.0f);

  std::vector<float> hwin(kFft, 0.0f);
  for (int i = 0; i < kFrameLen; ++i) hwin[i] = 0.5f * hann_host[i];  // exact scaling
  std::vector<float2> tw256(256), tw512(256);
  const double two_pi = 6.283185307179586476925286766559;
  for (int j = 0; j < 256; ++j) {
    tw256[j] = make_float2((float)cos(two_pi * j / 256.0), (float)(-sin(two_pi * j / 256.0)));
    tw512[j] = make_float2((float)cos(two_pi * j / 512.0), (float)(-sin(two_pi * j / 512.0)));
  }

  TasrFeaturizer* f = new TasrFeaturizer();
  memset(f, 0, sizeof(*f));
  f->p = *p;
  f->bands = bands;
  f->log_scale = p->log_base_e ? 0.69314718055994530942f : 0.30102999566398119521f;
  f->mel_fixed = tasr_mel_fixed_from_dense(mel_w_host, f->mel_fixed_w) ? 1 : 0;
  int rc = check_cuda(cudaGetDevice(&f->device), "cudaGetDevice");
  if (rc == TASR_OK) rc = check_cuda(cudaMalloc(&f->d_hwin, kFft * sizeof(float)), "cudaMalloc hwin");
  if (rc == TASR_OK) rc = check_cuda(cudaMalloc(&f->d_tw256, 256 * sizeof(float2)), "cudaMalloc tw256");
  if (rc == TASR_OK) rc = check_cuda(cudaMalloc(&f->d_tw512, 256 * sizeof(float2)), "cudaMalloc tw512");
  if (rc == TASR_OK) rc = check_cuda(cudaMalloc(&f->d_band_w, kMelBandMaxW4 * sizeof(float4)), "cudaMalloc band_w");
  if (rc == TASR_OK) rc = check_cuda(cudaMalloc(&f->d_bands, sizeof(MelBands)), "cudaMalloc bands");
  if (rc == TASR_OK) rc = check_cuda(cudaMemcpy(f->d_hwin, hwin.data(), kFft * sizeof(float), cudaMemcpyHostToDevice), "copy hwin");
  if (rc == TASR_OK) rc = check_cuda(cudaMemcpy(f->d_tw256, tw256.data(), 256 * sizeof(float2), cudaMemcpyHostToDevice), "copy tw256");
  if (rc == TASR_OK) rc = check_cuda(cudaMemcpy(f->d_tw512, tw512.data(), 256 * sizeof(float2), cudaMemcpyHostToDevice), "copy tw512");
  if (rc == TASR_OK) rc = check_cuda(cudaMemcpy(f->d_band_w, bw.data(), kMelBandMaxW4 * sizeof(float4), cudaMemcpyHostToDevice), "copy band_w");
  if (rc == TASR_OK) rc = check_cuda(cudaMemcpy(f->d_bands, &bands, sizeof(MelBands), cudaMemcpyHostToDevice), "copy bands");
  if (rc == TASR_OK && f->mel_fixed) {   // operand images of the tensor-core log-mel kernel (logmel_tc.cu)
    std::vector<unsigned char> img(16384, 0);
    tasr_logmel_tc_build_dft32(img.data());
    rc = check_cuda(cudaMalloc(&f->d_dft32, img.size()), "cudaMalloc dft32");
    if (rc == TASR_OK) rc = check_cuda(cudaMemcpy(f->d_dft32, img.data(), img.size(), cudaMemcpyHostToDevice), "copy dft32");
  }
  if (rc == TASR_OK && p->feature_type == TASR_FEAT_MFCC) {
    // tf.signal.mfccs_from_log_mel_spectrograms: dct(type 2, unnormalised: 2 sum x_n cos(pi k (2n+1)/(2M))) * rsqrt(2M)
    std::vector<float> dct((size_t)kMel * kMel);
    const double pi = 3.14159265358979323846, sc = 2.0 / sqrt(2.0 * kMel);
    for (int n = 0; n < kMel; ++n)
      for (int k = 0; k < kMel; ++k) dct[(size_t)n * kMel + k] = (float)(sc * cos(pi * k * (2.0 * n + 1.0) / (2.0 * kMel)));
    rc = check_cuda(cudaMalloc(&f->d_dct, dct.size() * sizeof(float)), "cudaMalloc dct");
    if (rc == TASR_OK) rc = check_cuda(cudaMemcpy(f->d_dct, dct.data(), dct.size() * sizeof(float), cudaMemcpyHostToDevice), "copy dct");
  }
  if (rc != TASR_OK) { tasr_featurizer_destroy(f); return rc; }
  *out = f;
  return TASR_OK;
}

extern "C" int tasr_featurizer_destroy(TasrFeaturizer* f) {
  if (!f) return TASR_OK;
  cudaFree(f->d_hwin);
  cudaFree(f->d_tw256);
  cudaFree(f->d_tw512);
  cudaFree(f->d_band_w);
  cudaFree(f->d_bands);
  cudaFree(f->d_dct);
  cudaFree(f->d_dft32);
  cudaFree(f->d_gwin);
  cudaFree(f->d_gmel);
  cudaFree(f->d_gtw);
  delete f;
  return TASR_OK;
}

extern "C" int tasr_featurizer_uses_fixed_mel(const TasrFeaturizer* f) { return f ? f->mel_fixed : -1; }

// ---------------------------------------------------------------------------------------
// Lengths after each conv layer + padding mask (src/utils/math_util.py:20-32, encoder.py:43-48)
// ---------------------------------------------------------------------------------------
namespace {

constexpr int kMaxLayers = 8;
struct ConvGeom {
  int32_t n;
  int32_t k[kMaxLayers], s[kMaxLayers], same[kMaxLayers];
};

// The reference does this arithmetic in float32 and truncates toward zero on the int32 cast;
// the same IEEE operations are issued here (no FMA contraction, explicit rounding intrinsics).
__device__ __forceinline__ int32_t conv_len_f32(int32_t L, int32_t k, int32_t s, int32_t same) {
  float l = (float)L, kf = (float)k, sf = (float)s;
  float r = same ? ceilf(__fdiv_rn(l, sf)) : __fadd_rn(__fdiv_rn(__fsub_rn(l, kf), sf), 1.0f);
  return (int32_t)r;  // cvt.rzi: truncation, like tf.cast(float32 -> int32)
}

__global__ void conv_lengths_kernel(const int32_t* __restrict__ len_in, int32_t B, ConvGeom g,
                                    int32_t* __restrict__ len_out) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int32_t L = len_in[b];
  for (int i = 0; i < g.n; ++i) {
    L = conv_len_f32(L, g.k[i], g.s[i], g.same[i]);
    len_out[(size_t)i * B + b] = L;
  }
}

// Lengths and padding mask in ONE launch: thread (b, t) recomputes the short float32 length chain of its
// utterance (a handful of instructions) and writes mask[b,t]; the t == 0 thread also stores the lengths.
__global__ void lengths_mask_kernel(const int32_t* __restrict__ len_in, int32_t B, ConvGeom g, int32_t W,
                                    int32_t* __restrict__ len_out, float* __restrict__ mask) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * W) return;
  const int b = (int)(i / W), t = (int)(i - (size_t)b * W);
  int32_t L = len_in[b];
  for (int l = 0; l < g.n; ++l) {
    L = conv_len_f32(L, g.k[l], g.s[l], g.same[l]);
    if (t == 0) len_out[(size_t)l * B + b] = L;
  }
  mask[i] = (t < L) ? 1.0f : 0.0f;
}

// ASRModel.create_masks' audio half (model.py:80): mask[b,t,f] = any_v (audio[b,t,f,v] != pad) as float32.
__global__ void audio_mask_kernel(const float* __restrict__ x, size_t n, int32_t V, float pad, float* __restrict__ mask) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int any = 0;
    for (int v = 0; v < V; ++v) any |= (x[i * V + v] != pad);
    mask[i] = any ? 1.0f : 0.0f;
  }
}

// n_frames[b] = #{t : any_f feat[b,t,f] != 0}  (model.py:80 + encoder.py:53-56).
// One warp per frame row, block-level count, one atomicAdd per block.
__global__ void count_nonzero_frames_kernel(const float* __restrict__ feat, int32_t T, int32_t F,
                                            int32_t* __restrict__ n_frames) {
  int b = blockIdx.y;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  int t = blockIdx.x * nwarp + warp;
  int any = 0;
  if (t < T) {
    const float* row = feat + ((size_t)b * T + t) * F;
    for (int f = lane; f < F; f += 32) any |= (row[f] != 0.0f);
  }
  any = __any_sync(0xffffffffu, any);
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  if (lane == 0 && any) atomicAdd(&cnt, 1);
  __syncthreads();
  if (threadIdx.x == 0 && cnt) atomicAdd(&n_frames[b], cnt);
}

}  // namespace

extern "C" int tasr_conv_lengths_mask(const int32_t* len_in, int32_t B, int32_t n_layers,
                                      const int32_t* k_host, const int32_t* s_host,
                                      const int32_t* same_host, int32_t* len_out, float* mask,
                                      int32_t mask_width, tasr_stream_t stream) {
  if (!len_in || !len_out || !k_host || !s_host || !same_host)
    return fail(TASR_ERR_BAD_ARG, "tasr_conv_lengths_mask: null argument");
  if (B < 0 || mask_width < 0) return fail(TASR_ERR_BAD_ARG, "tasr_conv_lengths_mask: negative size");
  if (n_layers < 1 || n_layers > kMaxLayers)
    return fail(TASR_ERR_BAD_ARG, "tasr_conv_lengths_mask: n_layers must be in [1,%d]", kMaxLayers);
  ConvGeom g;
  g.n = n_layers;
  for (int i = 0; i < n_layers; ++i) {
    if (k_host[i] < 1 || s_host[i] < 1) return fail(TASR_ERR_BAD_ARG, "tasr_conv_lengths_mask: kernel/stride must be >= 1");
    g.k[i] = k_host[i]; g.s[i] = s_host[i]; g.same[i] = same_host[i] ? 1 : 0;
  }
  if (B == 0) return TASR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (mask && mask_width > 0) {
    const size_t n = (size_t)B * mask_width;
    lengths_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(len_in, B, g, mask_width, len_out, mask);
    TASR_LAUNCH_CHECK("lengths_mask_kernel");
  } else {
    conv_lengths_kernel<<<(B + 127) / 128, 128, 0, st>>>(len_in, B, g, len_out);
    TASR_LAUNCH_CHECK("conv_lengths_kernel");
  }
  return TASR_OK;
}

extern "C" int tasr_audio_mask(const float* x, int64_t n, int32_t V, float pad_value, float* mask, tasr_stream_t stream) {
  if (!x || !mask) return fail(TASR_ERR_BAD_ARG, "tasr_audio_mask: null argument");
  if (n < 0 || V < 1) return fail(TASR_ERR_BAD_ARG, "tasr_audio_mask: bad shape");
  if (n == 0) return TASR_OK;
  const long long blocks = (n + 255) / 256;
  audio_mask_kernel<<<(unsigned)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, (cudaStream_t)stream>>>(x, (size_t)n, V, pad_value, mask);
  TASR_LAUNCH_CHECK("audio_mask_kernel");
  return TASR_OK;
}

extern "C" int tasr_count_nonzero_frames(const float* feat, int32_t B, int32_t T, int32_t F,
                                         int32_t* n_frames, tasr_stream_t stream) {
  if (!feat || !n_frames) return fail(TASR_ERR_BAD_ARG, "tasr_count_nonzero_frames: null argument");
  if (B < 0 || T < 0 || F < 1) return fail(TASR_ERR_BAD_ARG, "tasr_count_nonzero_frames: bad shape");
  if (B == 0) return TASR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  TASR_CUDA(cudaMemsetAsync(n_frames, 0, (size_t)B * sizeof(int32_t), st));
  if (T == 0) return TASR_OK;
  dim3 grid((T + 7) / 8, B);
  count_nonzero_frames_kernel<<<grid, 256, 0, st>>>(feat, T, F, n_frames);
  TASR_LAUNCH_CHECK("count_nonzero_frames_kernel");
  return TASR_OK;
}
