// Shared pieces of the log-mel kernel (logmel.cu; also used by the experiments under tools/experiments/):
// argument block, frame count, the compiled-in mel filterbank structure
// and its unrolled projection, small PTX helpers.
#pragma once
#include "common.cuh"

namespace tasr_lm {

using namespace tasr;

#include "mel_geometry.inc"

constexpr int kTileFrames = 32;
constexpr int kOutStride = kMel + 1;    // 81
constexpr int kPadChunkRows = 128;      // rows of collate padding zero-filled per work item


// Per-FFT-bin weights of the fixed-geometry mel projection, passed BY VALUE as a kernel parameter so
// that they sit in the constant bank and FFMA reads them as operands (no load instruction):
// wr[k] = W[k, seg(k)] (rising side of bin seg(k)), wf[k] = W[k, seg(k)-1] (falling side of the bin below).
struct MelFixedW {
  float wr[256];
  float wf[256];
};

__device__ __forceinline__ void st_global_v4(float* p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// log2 of a NORMAL positive float (the argument is clamped to output_floor >= FLT_MIN first, so the
// denormal rescue sequence of __log2f is dead weight): one MUFU.
__device__ __forceinline__ float lg2_normal(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}


struct LogmelArgs {
  const float* wav;
  const int32_t* len;
  const float* peak;       // may be null when !normalize
  float* peak_out;         // single-pass mode (non-null): the kernel itself accumulates max|x| per utterance here (zeroed
                           // by the launcher), runs on the UN-normalised signal and applies no floor; the reader adds
                           // 2*log(1/(peak+1e-9)) and clamps (TasrDeferredGain)
  float* out;
  int32_t* n_frames;
  const float* hwin;
  const float2* tw256;
  const float2* tw512;
  const float4* band_w;
  const MelBands* bands;
  int64_t row_stride;
  int32_t B, T_max, tiles_per_row;
  int32_t normalize;
  int32_t pad_end;         // tf.signal.stft(pad_end=True): ceil(N/160) frames, the tail zero padded
  int32_t mode;            // 0: mel projection + log; 1: log of the first 80 power bins ("spectrogram")
  int32_t tma;             // 1: the fixed-filterbank kernel stages the next tile's raw samples by cp.async.bulk (default; TASR_LOGMEL_TMA=0
                           // at launch selects the register-staged global loads, bit-identical)
  int32_t pad_fill_rows;   // < 0: every collate padding row (t >= n_frames[b]) is written as 0.0; >= 0: only the first
                           // pad_fill_rows of them are guaranteed (lean mode: the rest of out[b] is left untouched)
  float preemph, floor_, log_scale;
};

__device__ __forceinline__ int frames_of(int n, const LogmelArgs& a) {   // src/speech_featurizer.py:163-166
  const int Tb = a.pad_end ? (n > 0 ? (n + kFrameStep - 1) / kFrameStep : 0)
                           : ((n >= kFrameLen) ? 1 + (n - kFrameLen) / kFrameStep : 0);
  return min(Tb, a.T_max);
}

// First row of out[b] that need not be written: T_max, or (lean mode) pad_fill_rows past the utterance's frames.
__device__ __forceinline__ int pad_limit(int Tu, const LogmelArgs& a) {
  return a.pad_fill_rows < 0 ? a.T_max : min(a.T_max, Tu + a.pad_fill_rows);
}

// ---- fixed-geometry mel projection (config/model.yaml filterbank), fully unrolled ------------------
// Warp W owns mel bins [kMelGrp[W], kMelGrp[W+1]).  It walks segments m = first..last+1; in segment m
// each power bin k is loaded once and accumulated into bin m (rising weight) and bin m-1 (falling
// weight); bin m-1 is complete when segment m ends.  Summation is in ascending k, like a dot product.
template <int M, int M0, int M1>
struct MelSeg {
  static __device__ __forceinline__ void run(const float* __restrict__ Prow, const MelFixedW& w, float (&res)[M1 - M0], float acc_prev) {
    float acc_cur = 0.0f;
    constexpr int kBegin = kMelSegStart[M], kEnd = kMelSegStart[M + 1];
#pragma unroll
    for (int k = kBegin; k < kEnd; ++k) {
      const float p = Prow[k];
      if (M < M1) acc_cur = fmaf(p, w.wr[k], acc_cur);
      if (M > M0) acc_prev = fmaf(p, w.wf[k], acc_prev);
    }
    if constexpr (M > M0) res[M - 1 - M0] = acc_prev;
    if constexpr (M < M1) MelSeg<M + 1, M0, M1>::run(Prow, w, res, acc_cur);
  }
};

// The sums of a group stay in registers until the whole group is done: a store to the staging row after every bin would fence
// the loads of the power row behind it (both are shared memory), one exposed LDS latency per mel bin.
template <int W>
__device__ __forceinline__ void mel_fixed_group(const float* Prow, const MelFixedW& w, float* srow, float floor_, float scale) {
  constexpr int M0 = kMelGrp[W], M1 = kMelGrp[W + 1];
  float res[M1 - M0];
  MelSeg<M0, M0, M1>::run(Prow, w, res, 0.0f);
#pragma unroll
  for (int i = 0; i < M1 - M0; ++i) srow[M0 + i] = lg2_normal(fmaxf(res[i], floor_)) * scale;
}


}  // namespace tasr_lm

// logmel_tc.cu: launches the tensor-core kernel; returns a negative value (and launches nothing) when the handle /
// arguments are outside its scope, so that the caller falls back to logmel_kernel.
int tasr_logmel_tc_launch(const TasrFeaturizer* f, const tasr_lm::LogmelArgs& a, cudaStream_t st);
