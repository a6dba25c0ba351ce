/* Plain-C consumer of include/tasr.h: proves the boundary is a C ABI (no C++ / torch types), that the header
 * compiles as C, that libtasr_b200.so can be dlopen'ed by a C program and that argument validation works
 * without a GPU (every check below fails before any CUDA call).  Built and run by tests/test_c_abi.py. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "tasr.h"

#define LOAD(name)                                                   \
  do {                                                               \
    *(void**)(&p_##name) = dlsym(h, #name);                          \
    if (!p_##name) { printf("missing symbol %s\n", #name); return 2; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage: abi_check /path/to/libtasr_b200.so\n"); return 2; }
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { printf("dlopen failed: %s\n", dlerror()); return 2; }

  int (*p_tasr_version)(void);
  const char* (*p_tasr_last_error)(void);
  int (*p_tasr_featurizer_create)(const TasrFeatParams*, const float*, const float*, TasrFeaturizer**);
  int (*p_tasr_absmax_f32)(const float*, const int32_t*, int32_t, int64_t, float*, tasr_stream_t);
  int (*p_tasr_logmel_f32)(const TasrFeaturizer*, const float*, const int32_t*, const float*, int32_t, int64_t, float*,
                           int32_t, int32_t*, tasr_stream_t);
  int (*p_tasr_sepconv1d_f32)(const float*, int32_t, int32_t, const TasrSepConvLayer*, float*, int32_t, tasr_stream_t);
  int (*p_tasr_conv_lengths_mask)(const int32_t*, int32_t, int32_t, const int32_t*, const int32_t*, const int32_t*,
                                  int32_t*, float*, int32_t, tasr_stream_t);
  int (*p_tasr_unpack_pcm16)(const int16_t*, const int64_t*, const int32_t*, int32_t, int32_t, float*, int64_t, tasr_stream_t);
  int (*p_tasr_specaugment_f32)(float*, const int32_t*, int32_t, int32_t, int32_t, const int32_t*, int32_t, const int32_t*,
                                int32_t, tasr_stream_t);
  int (*p_tasr_logmel_f32_single_pass)(const TasrFeaturizer*, const float*, const int32_t*, int32_t, int64_t, float*, int32_t,
                                       int32_t*, int32_t, float*, TasrDeferredGain*, tasr_stream_t);
  int (*p_tasr_sepconv1d_tf32_ragged_lean)(const TasrSepConvPlan*, const float*, const int32_t*, int32_t, int32_t, int32_t,
                                           float*, int32_t, int32_t, const TasrDeferredGain*, tasr_stream_t);
  int32_t (*p_tasr_sepconv_ragged_margin)(void);
  int (*p_tasr_conv2d_plan_create)(const float*, const float*, const float*, const float*, int32_t, TasrConv2dPlan**, tasr_stream_t);
  int (*p_tasr_conv2d_output_shape)(int32_t, int32_t, int32_t*, int32_t*, int32_t*, int32_t*);
  int (*p_tasr_conv2d_subsample_ragged)(const TasrConv2dPlan*, const float*, const int32_t*, int32_t, int32_t, int32_t, void*,
                                        float*, const TasrDeferredGain*, tasr_stream_t);
  LOAD(tasr_logmel_f32_single_pass); LOAD(tasr_sepconv1d_tf32_ragged_lean); LOAD(tasr_sepconv_ragged_margin);
  LOAD(tasr_conv2d_plan_create); LOAD(tasr_conv2d_output_shape); LOAD(tasr_conv2d_subsample_ragged);
  LOAD(tasr_version); LOAD(tasr_last_error); LOAD(tasr_featurizer_create); LOAD(tasr_absmax_f32); LOAD(tasr_logmel_f32);
  LOAD(tasr_sepconv1d_f32); LOAD(tasr_conv_lengths_mask); LOAD(tasr_unpack_pcm16); LOAD(tasr_specaugment_f32);

  if (p_tasr_version() < 100) { printf("bad version\n"); return 1; }

  /* null arguments are refused with TASR_ERR_BAD_ARG and a message, nothing is launched */
  if (p_tasr_absmax_f32(0, 0, 1, 4, 0, 0) != TASR_ERR_BAD_ARG) { printf("absmax null check\n"); return 1; }
  if (strlen(p_tasr_last_error()) == 0) { printf("empty error text\n"); return 1; }
  if (p_tasr_logmel_f32(0, 0, 0, 0, 1, 4, 0, 1, 0, 0) != TASR_ERR_BAD_ARG) { printf("logmel null check\n"); return 1; }
  if (p_tasr_sepconv1d_f32(0, 1, 9, 0, 0, 1, 0) != TASR_ERR_BAD_ARG) { printf("sepconv null check\n"); return 1; }
  if (p_tasr_unpack_pcm16(0, 0, 0, 1, 4, 0, 4, 0) != TASR_ERR_BAD_ARG) { printf("unpack null check\n"); return 1; }
  if (p_tasr_specaugment_f32(0, 0, 1, 1, 80, 0, 0, 0, 0, 0) != TASR_ERR_BAD_ARG) { printf("specaugment null check\n"); return 1; }

  /* the round-1 additions: single-pass featurizer, lean ragged conv, Conv2dSubsampling */
  if (p_tasr_logmel_f32_single_pass(0, 0, 0, 1, 4, 0, 1, 0, -1, 0, 0, 0) != TASR_ERR_BAD_ARG) { printf("single-pass null check\n"); return 1; }
  if (p_tasr_sepconv1d_tf32_ragged_lean(0, 0, 0, 0, 1, 9, 0, 1, 39, 0, 0) != TASR_ERR_BAD_ARG) { printf("lean null check\n"); return 1; }
  if (p_tasr_sepconv_ragged_margin() < 38) { printf("ragged margin\n"); return 1; }
  { TasrConv2dPlan* cp = 0;
    if (p_tasr_conv2d_plan_create(0, 0, 0, 0, 144, &cp, 0) != TASR_ERR_BAD_ARG || cp != 0) { printf("conv2d plan null check\n"); return 1; }
    int32_t h1, w1, h2, w2;
    if (p_tasr_conv2d_output_shape(1498, 80, &h1, &w1, &h2, &w2) != TASR_OK || h1 != 749 || w1 != 40 || h2 != 375 || w2 != 20) {
      printf("conv2d output shape %d %d %d %d\n", h1, w1, h2, w2); return 1; }
    if (p_tasr_conv2d_subsample_ragged(0, 0, 0, 1, 8, 80, 0, 0, 0, 0) != TASR_ERR_BAD_ARG) { printf("conv2d ragged null check\n"); return 1; } }

  /* a frame geometry the kernels are not built for is TASR_ERR_UNSUPPORTED at handle creation */
  TasrFeatParams p;
  memset(&p, 0, sizeof(p));
  p.sample_rate = 16000; p.frame_length = 320; p.frame_step = 160; p.fft_length = 500; p.num_mel_bins = 80;   /* fft_length must be a power of two */
  p.preemphasis = 0.97f; p.output_floor = 1e-9f;
  static float hann[400], mel[257 * 80];
  TasrFeaturizer* f = 0;
  if (p_tasr_featurizer_create(&p, hann, mel, &f) != TASR_ERR_UNSUPPORTED || f != 0) { printf("geometry check\n"); return 1; }

  /* the length arithmetic entry point validates its layer count */
  int32_t k[1] = {9}, s[1] = {2}, same[1] = {0}, in[1] = {10}, out[1];
  if (p_tasr_conv_lengths_mask(in, 1, 0, k, s, same, out, 0, 0, 0) != TASR_ERR_BAD_ARG) { printf("layer count check\n"); return 1; }

  printf("C ABI OK: version %d, sizeof(TasrFeatParams)=%zu, sizeof(TasrSepConvLayer)=%zu\n", p_tasr_version(),
         sizeof(TasrFeatParams), sizeof(TasrSepConvLayer));
  dlclose(h);
  return 0;
}
