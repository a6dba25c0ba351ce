/* tasr.h — C ABI of the B200-native Telugu-ASR front-end hot path.
 *
 * The drop-in boundary for the reference's data-parallel front end:
 *   waveform -> log-mel            (src/speech_featurizer.py:136-161, per utterance on CPU)
 *   log-mel  -> subsampled [B,T3,d] (src/models/moonshine/encoder.py:50-71, Keras/cuDNN)
 *   lengths  -> padding mask        (src/utils/math_util.py:20-32, encoder.py:43-48)
 *
 * Plain pointers and sizes only; no torch / C++ types.  All data pointers are DEVICE
 * pointers unless the name ends in `_host`.  Every launch entry point takes the CUDA
 * stream to enqueue on and returns immediately (asynchronous); nothing allocates per
 * call.  Return value: 0 = ok, otherwise a TASR_ERR_* code; `tasr_last_error()` gives
 * the message for the calling thread.  Nothing throws across this boundary.
 *
 * Built for sm_100a only (libtasr_b200.so); there is no CPU fallback.
 */
#ifndef TASR_H_
#define TASR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* tasr_stream_t; /* == cudaStream_t */

enum {
  TASR_OK = 0,
  TASR_ERR_BAD_ARG = 1,      /* null pointer, negative size, inconsistent shapes            */
  TASR_ERR_UNSUPPORTED = 2,  /* parameter combination the sm_100a kernels are not built for */
  TASR_ERR_MISALIGNED = 3,   /* pointer / row stride not 16-byte aligned                     */
  TASR_ERR_CUDA = 4          /* a CUDA runtime call failed; message carries cudaGetErrorString */
};

/* Activation codes of tasr_sepconv1d_f32 (Keras names: encoder.py:25,36). */
enum { TASR_ACT_NONE = 0, TASR_ACT_TANH = 1, TASR_ACT_GELU_ERF = 2, TASR_ACT_RELU = 3 };

/* Arithmetic of the pointwise (1x1) contraction. */
enum {
  TASR_MATH_FP32 = 0, /* FP32 FMA on CUDA cores (bit-for-bit deterministic, slowest)          */
  TASR_MATH_TF32 = 1  /* tcgen05.mma kind::tf32, operands rounded to TF32 (rna), FP32 accum.  */
};

/* Feature parameters: the `speech_config` block of config/model.yaml:1-17 as the
 * SpeechFeaturizer constructor derives them (src/speech_featurizer.py:41-65). */
typedef struct TasrFeatParams {
  int32_t sample_rate;      /* 16000                                                        */
  int32_t frame_length;     /* int(round(sample_rate*frame_ms/1000))  = 400  (:46)           */
  int32_t frame_step;       /* int(round(sample_rate*stride_ms/1000)) = 160  (:49)           */
  int32_t fft_length;       /* enclosing power of two of frame_length = 512 (tf.signal.stft) */
  int32_t num_mel_bins;     /* num_feature_bins = 80                                         */
  int32_t normalize_signal; /* 1: x *= 1/(max|x|+1e-9) per utterance (:68-72)                */
  int32_t log_base_e;       /* 0: log10 (log_base "10"), 1: natural log (:107-110)           */
  int32_t pad_end;          /* 0 (config/model.yaml:8); 1: tf.signal.stft(pad_end=True), ceil(N/step) frames */
  float preemphasis;        /* 0.97; <= 0 disables (:74-79)                                  */
  float output_floor;       /* 1e-9 (:109)                                                   */
  int32_t feature_type;     /* TASR_FEAT_* (:136-153); 0 = log_mel_spectrogram               */
  int32_t normalize_zscore; /* 1: per frame (x-mean)/sqrt(var+1e-9) over the feature axis (:82-85)   */
  int32_t normalize_min_max;/* 1: per frame (x-min)/(max-min) (:86-91); zscore wins if both are set  */
} TasrFeatParams;

/* feature_type codes (FeaturizerConfig, src/speech_featurizer.py:10-15). */
enum { TASR_FEAT_LOG_MEL = 0, TASR_FEAT_SPECTROGRAM = 1, TASR_FEAT_MFCC = 2, TASR_FEAT_WAVEFORM = 3 };

/* One SeparableConv1D layer (encoder.py:31-40), weights on the DEVICE, float32:
 * dw [kernel, c_in] (Keras depthwise_kernel (k,Cin,1)), pw [c_in, c_out] (pointwise_kernel
 * (1,Cin,Cout)), bias [c_out]. */
typedef struct TasrSepConvLayer {
  const float* dw;
  const float* pw;
  const float* bias;
  int32_t c_in;
  int32_t c_out;
  int32_t kernel;     /* 9 */
  int32_t stride;     /* 2 */
  int32_t same;       /* 0 = "valid" (config/model.yaml:26); 1 = "same" (the reference constructor's default, encoder.py:24;
                         TensorFlow's rule for the tensor length t_in): dense entry points only, the ragged ones refuse it */
  int32_t activation; /* TASR_ACT_* */
} TasrSepConvLayer;

typedef struct TasrFeaturizer TasrFeaturizer; /* opaque, owns the per-device constant tables */
typedef struct TasrSepConvPlan TasrSepConvPlan; /* opaque, owns TF32-packed pointwise weights */

int tasr_version(void);
const char* tasr_last_error(void);
/* Number of CUDA kernels this library has launched in this process (all threads, all devices);
 * bench.py reports the difference over the timed region as `gpu_launches`. */
int64_t tasr_launch_count(void);

/* Replaces SpeechFeaturizer.__init__ (src/speech_featurizer.py:19-66) plus the per-call
 * tf.signal.hann_window / linear_to_mel_weight_matrix construction (:96-101, :114-120):
 * the caller builds both tables ONCE in float32 on the host (same op order as TF) and the
 * handle uploads them to the current device together with float64-derived FFT twiddles.
 * hann_host [frame_length], mel_w_host [fft_length/2+1, num_mel_bins] row-major. */
int tasr_featurizer_create(const TasrFeatParams* params, const float* hann_host,
                           const float* mel_w_host, TasrFeaturizer** out);
int tasr_featurizer_destroy(TasrFeaturizer* f);
/* 1 when the handle's mel matrix has the config/model.yaml sparsity structure and tasr_logmel_f32
 * runs the fully unrolled projection, 0 when it runs the generic banded loop, < 0 on a null handle. */
int tasr_featurizer_uses_fixed_mel(const TasrFeaturizer* f);

/* Device-side collate of raw audio (src/dataset.py:167-175 hands the featurizer one decoded
 * utterance at a time; src/utils/data_util.py:31 decodes int16 PCM to float32 = sample/32768).
 * `packed` holds the B utterances back to back on the DEVICE (only valid samples, so padding never
 * crosses PCIe), utterance b at sample offset[b] (a multiple of 8, int64, device) with len[b]
 * samples; they are written to wav[b*row_stride ...].  Samples beyond len[b] are not written (no
 * kernel reads them).  max_len >= max_b len[b] sizes the grid.  The pcm16 variant converts exactly
 * like tf.audio.decode_wav: float32(sample) * 2^-15. */
int tasr_unpack_f32(const float* packed, const int64_t* offset, const int32_t* len, int32_t batch,
                    int32_t max_len, float* wav, int64_t row_stride, tasr_stream_t stream);
int tasr_unpack_pcm16(const int16_t* packed, const int64_t* offset, const int32_t* len, int32_t batch,
                      int32_t max_len, float* wav, int64_t row_stride, tasr_stream_t stream);

/* The way back of the device collate: x [batch, t, c] whose rows t >= len[b] are collate padding -> packed [sum_b len[b], c]
 * (utterance b at row offset[b]; offset has batch + 1 entries, the last is the total; int64, device).  A consumer that reads
 * only valid rows (src/models/moonshine/encoder.py:237-247 masks the rest) can fetch the packed rows instead of the padded
 * tensor: about half the device->host bytes on a ragged batch.  packed must hold sum_b min(len[b], t) rows; c % 4 == 0. */
int tasr_pack_valid_rows(const float* x, const int32_t* len, int32_t batch, int32_t t, int32_t c, float* packed,
                         int64_t* offset, tasr_stream_t stream);

/* Replaces tf.reduce_max(tf.abs(signal)) (src/speech_featurizer.py:70), batched:
 * peak[b] = max_{n < len[b]} |wav[b*row_stride + n]|.  peak is overwritten. */
int tasr_absmax_f32(const float* wav, const int32_t* len, int32_t batch, int64_t row_stride,
                    float* peak, tasr_stream_t stream);

/* Replaces SpeechFeaturizer.call with feature_type "waveform" (src/speech_featurizer.py:132-133,142-143):
 * out[b, n] = preemphasis(normalize_signal(wav[b, :len[b]]))[n] for n < len[b]; samples beyond len[b] of
 * out are not written.  out has the layout of wav (same row stride). */
int tasr_waveform_f32(const TasrFeaturizer* f, const float* wav, const int32_t* len, const float* peak_or_null,
                      int32_t batch, int64_t row_stride, float* out, tasr_stream_t stream);

/* Replaces SpeechFeaturizer.call on each utterance (src/speech_featurizer.py:136-161:
 * normalize_signal -> preemphasis_signal -> stft -> mel matmul -> logarithm; for the other feature types
 * of the handle: spectrogram = log power of the first num_mel_bins FFT bins (:124-126), mfcc = DCT-II of
 * the log-mel (:128-130); then normalize_audio_feature (:81-93) when the handle asks for it) AND the
 * zero-padded collate that follows it (src/dataset.py:173-175, 236-252).
 * wav [batch, row_stride] (samples beyond len[b] are never read); peak from
 * tasr_absmax_f32 (may be NULL when params.normalize_signal == 0);
 * out [batch, t_max, num_mel_bins] float32: rows t < n_frames[b] hold log-mel, rows
 * t >= n_frames[b] are written as 0.0; n_frames[b] = max(0, 1+(len[b]-frame_length)/frame_step)
 * (src/speech_featurizer.py:163-166), clamped to t_max. */
int tasr_logmel_f32(const TasrFeaturizer* f, const float* wav, const int32_t* len,
                    const float* peak_or_null, int32_t batch, int64_t row_stride, float* out,
                    int32_t t_max, int32_t* n_frames, tasr_stream_t stream);

/* Lean variant for a feature tensor that is only an INTERMEDIATE (its one reader is the ragged separable
 * convolution below, which never looks further than tasr_sepconv_ragged_margin() rows past the data): of the
 * collate padding only rows n_frames[b] <= t < n_frames[b] + pad_fill_rows are written as 0.0; the rest of
 * out[b] is left untouched.  Rows t < n_frames[b] and n_frames are identical to tasr_logmel_f32.  mfcc and
 * per-frame normalisation handles are TASR_ERR_UNSUPPORTED (they post-process the whole tensor). */
int tasr_logmel_f32_lean(const TasrFeaturizer* f, const float* wav, const int32_t* len,
                         const float* peak_or_null, int32_t batch, int64_t row_stride, float* out,
                         int32_t t_max, int32_t* n_frames, int32_t pad_fill_rows, tasr_stream_t stream);

/* Single-pass variant (the waveform is read ONCE; tasr_absmax_f32 is not needed): the kernel accumulates
 * peak_out[b] = max|x_b| while it stages the samples, featurises the UN-normalised signal and applies no floor, so
 * out holds log(mel power of x) — -inf where that power is 0.  Because the featurizer is scale-covariant,
 *   log-mel(x / (peak + 1e-9))[t, m] = max( out[t, m] + 2*log(1 / (peak + 1e-9)),  log(output_floor) )
 * (src/speech_featurizer.py:68-72, 107-110), which the READER applies: tasr_sepconv1d_tf32_ragged_lean takes the
 * TasrDeferredGain this call fills in (host struct, device peak pointer), tasr_apply_deferred_gain materialises the
 * reference's feature values in place.  Differences to the two-pass result are float32 rounding only (the gain is
 * applied after the arithmetic instead of before it).  pad_fill_rows as in tasr_logmel_f32_lean (< 0: all rows).
 * Needs normalize_signal = 1, pad_end = 0, feature_type log-mel or spectrogram, no per-frame normalisation. */
typedef struct TasrDeferredGain {
  const float* peak;   /* device [batch]: max|x_b|                                   */
  float log_scale_x2;  /* 2*log10(2) or 2*ln(2): the gain enters the power squared   */
  float log_floor;     /* log(output_floor) in the handle's base                     */
} TasrDeferredGain;
int tasr_logmel_f32_single_pass(const TasrFeaturizer* f, const float* wav, const int32_t* len, int32_t batch,
                                int64_t row_stride, float* out, int32_t t_max, int32_t* n_frames,
                                int32_t pad_fill_rows, float* peak_out, TasrDeferredGain* gain_host,
                                tasr_stream_t stream);
/* feat[b, t, :] = max(feat[b, t, :] + 2*log(1/(peak[b]+1e-9)), log_floor) for t < n_frames[b], in place. */
int tasr_apply_deferred_gain(float* feat, const int32_t* n_frames, int32_t batch, int32_t t_max, int32_t f,
                             const TasrDeferredGain* gain_host, tasr_stream_t stream);

/* Replaces one tf.keras.layers.SeparableConv1D forward (encoder.py:31-40, called at :60):
 * y[b,t,o] = act( sum_c ( sum_k x[b, stride*t+k, c] * dw[k,c] ) * pw[c,o] + bias[o] ),
 * "valid" padding, t < t_out where t_out <= (t_in-kernel)/stride+1.  x [batch,t_in,c_in],
 * y [batch,t_out,c_out].  The convolution runs over the whole zero-padded tensor, like the
 * reference (no masking between layers).  math = TASR_MATH_FP32 here. */
int tasr_sepconv1d_f32(const float* x, int32_t batch, int32_t t_in, const TasrSepConvLayer* layer,
                       float* y, int32_t t_out, tasr_stream_t stream);

/* TF32 tensor-core variant: the plan packs pw (rounded to TF32, UMMA K-major tiles) once. */
int tasr_sepconv_plan_create(const TasrSepConvLayer* layer, TasrSepConvPlan** out, tasr_stream_t stream);
int tasr_sepconv_plan_destroy(TasrSepConvPlan* plan);
int tasr_sepconv1d_tf32(const TasrSepConvPlan* plan, const float* x, int32_t batch, int32_t t_in,
                        float* y, int32_t t_out, tasr_stream_t stream);

/* Ragged form of the same layer for zero-padded batches (src/dataset.py:236-252 pads with 0.0 and the
 * reference then convolves the padding too, encoder.py:60 — far from the data that just reproduces one
 * constant row per layer).  tasr_sepconv_plan_set_pad_row tells the plan what every padding row of its
 * INPUT looks like (device [c_in]; NULL = all zeros, i.e. the collate's padding) and computes, with the
 * layer's own arithmetic, the row it produces from a receptive field made of that row;
 * tasr_sepconv_plan_pad_row returns it (device [c_out]) so the next layer's plan can be chained.
 * tasr_sepconv1d_tf32_ragged then takes len0 [batch] (device int32) and shift: rows
 * t >= ceil(len0[b] / 2^shift) of x[b] must all equal that padding row (len0 = n_frames of the features,
 * shift = index of the layer in the stride-2 stack).  Tiles whose receptive field lies entirely there are
 * filled with the constant row instead of being computed; every value of y — valid and padded — is
 * bit-identical to what tasr_sepconv1d_tf32 writes for the same input. */
int tasr_sepconv_plan_set_pad_row(TasrSepConvPlan* plan, const float* pad_row_in, tasr_stream_t stream);
const float* tasr_sepconv_plan_pad_row(const TasrSepConvPlan* plan);
int tasr_sepconv1d_tf32_ragged(const TasrSepConvPlan* plan, const float* x, const int32_t* len0, int32_t shift,
                               int32_t batch, int32_t t_in, float* y, int32_t t_out, tasr_stream_t stream);

/* Lean variant for an INTERMEDIATE activation tensor whose only reader is the next ragged layer of the stack
 * (src/models/moonshine/encoder.py:58-68 keeps no intermediate): of the rows that only repeat the constant padding
 * row — t >= ceil(len0[b] / 2^(shift+1)) — at least the first fill_rows are written (whole 128-row tiles), the
 * tiles beyond are left untouched (fill_rows < 0: every row is written).  All other rows are identical to
 * tasr_sepconv1d_tf32_ragged.
 * tasr_sepconv_ragged_margin() is the number of input rows past ceil(len0[b] / 2^shift) that the ragged kernels
 * may read: a producer in lean mode must be given fill_rows / pad_fill_rows >= that. */
int32_t tasr_sepconv_ragged_margin(void);
int tasr_sepconv1d_tf32_ragged_lean(const TasrSepConvPlan* plan, const float* x, const int32_t* len0, int32_t shift,
                                    int32_t batch, int32_t t_in, float* y, int32_t t_out, int32_t fill_rows,
                                    const TasrDeferredGain* input_gain_or_null, tasr_stream_t stream);
/* input_gain_or_null (host struct; first layer only, shift == 0): x comes from tasr_logmel_f32_single_pass and every
 * row t < len0[b] is read as max(x + 2*log(1/(peak[b]+1e-9)), log_floor); fill_rows < 0 writes every row of y. */

/* Replaces math_util.get_conv_length applied per layer (src/utils/math_util.py:20-32,
 * encoder.py:60-68) and lengths_to_padding_mask (encoder.py:43-48).
 * len_out [n_layers, batch] int32 gets the length after every layer, computed exactly as the
 * reference does: int32(trunc(float32((L-k)/s + 1))) for valid, int32(ceil(L/s)) for same.
 * mask (may be NULL) [batch, mask_width] float32 gets (t < len_last[b]).  The reference's
 * mask width is max_b(len_last[b]); the caller passes that (or any upper bound). */
int tasr_conv_lengths_mask(const int32_t* len_in, int32_t batch, int32_t n_layers,
                           const int32_t* kernel_host, const int32_t* stride_host,
                           const int32_t* same_host, int32_t* len_out, float* mask,
                           int32_t mask_width, tasr_stream_t stream);

/* Conv2dSubsampling of the conformer configuration (src/models/conformer/encoder.py:9-67, config/conformer.yaml:22-27):
 * two tf.keras.layers.Conv2D(filters, 3, strides 2, padding "same") each followed by ReLU, then merge_two_last_dims
 * (src/utils/math_util.py:34-46).  Weights in Keras shapes, float32 on the DEVICE: w1 [3,3,1,filters], b1 [filters],
 * w2 [3,3,filters,filters], b2 [filters]; the plan copies / packs them (w2 rounded to FP16 as UMMA tiles); filters
 * must be a multiple of 8.
 * feat [batch, t, w] (the [B,T,80,1] features) -> out [batch, h2, w2*filters] with h1 = ceil(t/2), h2 = ceil(h1/2),
 * w1 = ceil(w/2), w2 = ceil(w1/2) (tasr_conv2d_output_shape); h1_workspace [batch, h1, w1, filters] of 2-byte
 * elements (FP16; 16-byte aligned) is caller-provided scratch.  TensorFlow "SAME" padding (the odd row / column goes after).  The second convolution is an
 * implicit GEMM on tcgen05 (kind::f16: the first layer's ReLU output and w2 as FP16 — 11 significant bits, like
 * TF32 — with FP32 accumulation).  The reference passes the lengths through
 * get_conv_length ONCE (encoder.py:59-64: ceil(L/2)): use tasr_conv_lengths_mask with one "same" layer (k 3, s 2). */
typedef struct TasrConv2dPlan TasrConv2dPlan;
int tasr_conv2d_plan_create(const float* w1, const float* b1, const float* w2, const float* b2, int32_t filters,
                            TasrConv2dPlan** out, tasr_stream_t stream);
int tasr_conv2d_plan_destroy(TasrConv2dPlan* plan);
int tasr_conv2d_output_shape(int32_t t, int32_t w, int32_t* h1, int32_t* w1, int32_t* h2, int32_t* w2);
int tasr_conv2d_subsample(const TasrConv2dPlan* plan, const float* feat, int32_t batch, int32_t t, int32_t w,
                               void* h1_workspace, float* out, tasr_stream_t stream);

/* Ragged form for zero-padded batches (src/dataset.py:236-252 pads the features with 0.0 and the reference convolves the
 * padding too): n_frames [batch] (device int32) says that feature rows t >= n_frames[b] are zero — they are then never
 * read, so the features may be a lean tensor.  An output row whose whole receptive field is padding repeats one
 * [w2, filters] pattern (the bottom border row excepted); tasr_conv2d_plan_prepare_ragged computes that pattern for
 * feature width w with these very kernels on a zero input (synchronises the stream; call once per width, outside any
 * graph capture), and tasr_conv2d_subsample_ragged fills the 128-position tiles that consist only of such rows instead
 * of computing them, and skips the first-layer rows nobody reads.  Every value of `out` is bit-identical to
 * tasr_conv2d_subsample on the same (zero-padded) input. */
int tasr_conv2d_plan_prepare_ragged(TasrConv2dPlan* plan, int32_t w, tasr_stream_t stream);
int tasr_conv2d_subsample_ragged(const TasrConv2dPlan* plan, const float* feat, const int32_t* n_frames, int32_t batch,
                                 int32_t t, int32_t w, void* h1_workspace, float* out,
                                 const TasrDeferredGain* input_gain_or_null, tasr_stream_t stream);
/* input_gain_or_null (host struct): feat comes from tasr_logmel_f32_single_pass and every row t < n_frames[b] is read as
 * max(feat + 2*log(1/(peak[b]+1e-9)), log_floor) by the first convolution. */

/* SpecAugment, deterministic half (replaces FreqMasking.augment / TimeMasking.augment,
 * src/augmentations/specaugment.py:6-62, applied per utterance at src/dataset.py:172): in place on
 * feat [batch, t_max, f]: feat[b,t,:] *= 0 for t in [t0, t0+t) of every time mask, feat[b,:,k] *= 0 for
 * k in [f0, f0+f) of every frequency mask, rows t < n_frames[b] only.  time_masks [batch, n_time, 2] =
 * (t0, t), freq_masks [batch, n_freq, 2] = (f0, f), device int32; a mask of width 0 is "not applied"
 * (the reference applies each augmentation with probability `prob`).  The draws are the caller's. */
int tasr_specaugment_f32(float* feat, const int32_t* n_frames, int32_t batch, int32_t t_max, int32_t f,
                         const int32_t* time_masks, int32_t n_time, const int32_t* freq_masks, int32_t n_freq,
                         tasr_stream_t stream);

/* Replaces the audio half of ASRModel.create_masks (model.py:80): mask[i] = any_v (x[i*V + v] != pad_value)
 * as float32 0/1, i < n.  For audio_inputs [B,T,F,1]: n = B*T*F, V = 1. */
int tasr_audio_mask(const float* x, int64_t n, int32_t v, float pad_value, float* mask, tasr_stream_t stream);

/* Replaces ASRModel.create_masks + the mask->length reduction (model.py:80, encoder.py:53-56):
 * n_frames[b] = #{t : any_f feat[b,t,f] != 0.0}.  feat [batch, t, f]. */
int tasr_count_nonzero_frames(const float* feat, int32_t batch, int32_t t, int32_t f,
                              int32_t* n_frames, tasr_stream_t stream);

/* First encoder block.  Replaces EncoderBlock.call (src/models/moonshine/encoder.py:151-154) =
 * MHSAModule (src/models/layers/attention.py:519-602; MultiHeadAttention :44-230 with RoPE, positional_encoding.py:20-93)
 * followed by FFNModule (src/models/layers/mlp.py:9-60), inference mode (dropout = identity):
 *   q, k, v = x Wq, x Wk, x Wv (no bias); RoPE on q and k; softmax(q k^T / sqrt(head_dim) + (1 - mask) * -1e9) v; . Wo;
 *   h1 = LayerNorm(x + .); out = LayerNorm(gelu(h1 W1 + b1) W2 + b2 + h1)       (exact-erf GELU, LayerNorm epsilon ln_eps).
 * Weights float32 on the DEVICE in Keras shapes: wq / wk / wv [d_model, num_heads*head_dim], wo [num_heads*head_dim, d_model],
 * w1 [d_model, d_model*fc_factor], b1 [d_model*fc_factor], w2 [d_model*fc_factor, d_model], b2 [d_model], LayerNorm gamma / beta
 * [d_model].  The plan packs the four kernels as TF32 UMMA tiles; bias and LayerNorm pointers are borrowed (must outlive it).
 * Built for the config/model.yaml shape: d_model 192 = 6 heads x 32 (rot_dim = max(head_dim // 2, 32) = head_dim), fc_factor 1..8;
 * anything else returns TASR_ERR_UNSUPPORTED.  `MHSAModule.call` unpacks `inputs, pos = inputs` (attention.py:572) although
 * EncoderBlock hands it one tensor; `pos` is unused by the 'sdpa' attention the encoder builds, so x is the whole [batch, t, d_model]
 * tensor here (oracle/encoder_block_ref.py states the resolution).
 * x [batch, t, d_model]; len [batch] device int32 = valid tokens per utterance (the padding mask of encoder.py:43-48 as a prefix
 * length; NULL = no mask); workspace = tasr_encoder_block_workspace_floats(plan, batch, t) floats, 16-byte aligned; out [batch, t,
 * d_model].  Rows t >= len[b] attend uniformly to all t keys, as the reference's float32 "-1e9" masking does.  Enqueue-only once
 * tasr_encoder_block_prepare(plan, t_max) has sized the RoPE table (otherwise the first call with a larger t allocates). */
typedef struct TasrEncoderBlockWeights {
  const float *wq, *wk, *wv, *wo, *ln1_gamma, *ln1_beta, *w1, *b1, *w2, *b2, *ln2_gamma, *ln2_beta;
  int32_t d_model, num_heads, head_dim, fc_factor;
  float ln_eps;            /* tf.keras.layers.LayerNormalization default: 1e-3 */
} TasrEncoderBlockWeights;
typedef struct TasrEncoderBlockPlan TasrEncoderBlockPlan;
int tasr_encoder_block_plan_create(const TasrEncoderBlockWeights* weights, TasrEncoderBlockPlan** out, tasr_stream_t stream);
int tasr_encoder_block_plan_destroy(TasrEncoderBlockPlan* plan);
int tasr_encoder_block_prepare(TasrEncoderBlockPlan* plan, int32_t t_max);
int64_t tasr_encoder_block_workspace_floats(const TasrEncoderBlockPlan* plan, int32_t batch, int32_t t);
int tasr_encoder_block_f32(TasrEncoderBlockPlan* plan, const float* x, const int32_t* len, int32_t batch, int32_t t,
                           int32_t use_causal_mask, float* workspace, float* out, tasr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TASR_H_ */
