// Shared helpers for the tasr C-ABI implementation (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/tasr.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "telugu_asr_b200 kernels are written for sm_100a (B200) only"
#endif

namespace tasr {

void set_error(const char* fmt, ...);
void count_launch();

inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  set_error("%s", buf);
  return code;
}

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return TASR_OK;
  return fail(TASR_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define TASR_CUDA(call)                                            \
  do {                                                             \
    int _rc = ::tasr::check_cuda((call), #call);                   \
    if (_rc != TASR_OK) return _rc;                                \
  } while (0)

#define TASR_LAUNCH_CHECK(name)                                    \
  do {                                                             \
    ::tasr::count_launch();                                        \
    int _rc = ::tasr::check_cuda(cudaGetLastError(), name);        \
    if (_rc != TASR_OK) return _rc;                                \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sm_count();  // SMs of the current device (cached per device)

// |x| as its bit pattern.  Unsigned-integer max over these orders finite values like fmaxf does and ranks a NaN above
// +inf, so max|x| reductions propagate NaN the way tf.reduce_max(tf.abs(x)) does (src/speech_featurizer.py:70).
__device__ __forceinline__ unsigned abs_bits(float x) { return __float_as_uint(x) & 0x7fffffffu; }

// Fixed front-end geometry the kernels are specialised for (config/model.yaml:1-17).
constexpr int kFrameLen = 400;
constexpr int kFrameStep = 160;
constexpr int kFft = 512;
constexpr int kBins = 257;
constexpr int kMel = 80;

// Banded view of the [257,80] mel matrix: for mel bin m the non-zero weights are the
// contiguous FFT bins [k0[m], k0[m]+4*n4[m]) (zero padded to a multiple of four).
struct MelBands {
  int32_t k0[kMel];
  int32_t n4[kMel];
  int32_t off4[kMel];  // offset of the first float4 of bin m in `w`
  int32_t total4;
};
constexpr int kMelBandMaxW4 = 256;  // capacity in float4 (1024 weights; dense HTK needs ~160)

}  // namespace tasr

// Device-side handle contents.
struct TasrFeaturizer {
  TasrFeatParams p;
  int device;
  float* d_hwin;        // [512]  0.5*hann zero padded (0.5 folds the real-FFT split's 1/2)
  float2* d_tw256;      // [256]  exp(-2*pi*i*j/256), float64-derived
  float2* d_tw512;      // [256]  exp(-2*pi*i*j/512)
  float4* d_band_w;     // [kMelBandMaxW4]
  tasr::MelBands* d_bands;  // device copy of `bands`
  tasr::MelBands bands; // host copy
  float log_scale;      // log10(2) or ln(2)
  float* d_dct;         // [80][80] mfcc basis D[n][k] = 2 cos(pi k (2n+1) / 160) / sqrt(160), or null
  int mel_fixed;        // 1: the matrix has the compiled-in config/model.yaml structure (mel_geometry.inc)
  float mel_fixed_w[512];  // wr[256] | wf[256], per FFT bin (kernel-parameter constants of the unrolled projection)
  unsigned char* d_dft32;  // [16 KB] UMMA B images of the DFT-32 matrix, FP16 high | low parts (logmel_tc.cu)
  // any other frame geometry (logmel_generic.cu): window [frame_length], dense mel matrix [fft/2+1, n_mel], twiddles [fft/2]
  int generic, g_log2fft;
  float* d_gwin;
  float* d_gmel;
  float2* d_gtw;
};

// feature_post.cu: mfcc DCT and/or per-frame z-score / min-max normalisation, in place on [B, T_max, 80].
int tasr_feature_post_launch(const TasrFeaturizer* f, float* feat, const int32_t* n_frames, int32_t B, int32_t T_max,
                             cudaStream_t st);

// logmel_generic.cu: featurizer for frame geometries other than 400 / 160 / 512 / 80
int tasr_logmel_generic_create(TasrFeaturizer* f, const float* hann_host, const float* mel_w_host);
int tasr_logmel_generic_launch(const TasrFeaturizer* f, const float* wav, const int32_t* len, const float* peak, int32_t B,
                               int64_t row_stride, float* out, int32_t T_max, int32_t* n_frames, cudaStream_t st);

// logmel_tc.cu: host builder of the DFT-32 operand images (16 KB)
void tasr_logmel_tc_build_dft32(unsigned char* img16k);

// logmel.cu: true (and wr_wf_512 filled) when the dense [257,80] matrix has the compiled-in structure.
bool tasr_mel_fixed_from_dense(const float* mel_w_host, float* wr_wf_512);
