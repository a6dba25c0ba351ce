// Fused waveform -> log-mel kernel (sm_100a).
//
// Replaces, per utterance, src/speech_featurizer.py:136-161 (normalize_signal ->
// preemphasis_signal -> tf.signal.stft(400/160, periodic Hann, rFFT-512) -> |X|^2 -> HTK mel
// matmul -> log10(max(.,1e-9))) and the zero-padded collate of src/dataset.py:236-252.
//
// Work decomposition
//   tile  = 32 consecutive frames of one utterance (5360 samples, staged once in shared memory
//           with gain and pre-emphasis applied in exactly the reference's float32 op order);
//   FFT   = 16 lanes per frame (two frames per warp).  The 512-point real FFT is a 256-point
//           complex FFT of z[m] = y[2m] + i*y[2m+1] done as 16x16: radix-16 in registers,
//           twiddle, 16x16 transpose through a padded per-warp scratch, radix-16 again; then the
//           real-FFT split, where lane t and lane 16-t exchange eight values by warp shuffle and
//           each forms |X[k]|^2 and |X[256-k]|^2 for its eight k;
//   mel   = lane <-> frame, warp <-> 10 mel bins; banded FP32 FMAs over the 502 non-zero weights
//           (weights broadcast from shared memory, power rows read conflict-free, stride 261);
//   out   = log, staged through shared memory, written with coalesced 128-bit stores; rows
//           t >= n_frames[b] are written as 0.0 (the collate padding value).
//
// Per-lane constants (half-window, transpose twiddles, split twiddle) live in registers for the
// whole persistent loop over tiles.  Twiddles are float64-derived tables.
#include "common.cuh"

using namespace tasr;

namespace {

constexpr int kTileFrames = 32;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kWavSmem = 5376;          // (32-1)*160+400 = 5360, +16 floats lanes 8..15 touch at m2=12
constexpr int kScrStride = 17;          // float2 units; odd -> conflict-free transposed reads
constexpr int kScrPerFrame = 16 * kScrStride;
constexpr int kPStride = kBins + 4;     // 261, odd; columns 257..260 stay zero (band padding)
constexpr int kOutStride = kMel + 1;    // 81

struct __align__(16) Smem {
  float wav[kWavSmem];
  float2 scr[kWarps * 2 * kScrPerFrame];   // also the [32][81] output staging tile
  float P[kTileFrames * kPStride];
  float4 band_w[kMelBandMaxW4];
  MelBands bands;
};
static_assert(sizeof(float) * kTileFrames * kOutStride <= sizeof(float2) * kWarps * 2 * kScrPerFrame,
              "output staging must fit in the transpose scratch");

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}

// Forward 4-point DFT in place (W4 = -i).
__device__ __forceinline__ void fft4(float2& p0, float2& p1, float2& p2, float2& p3) {
  float2 s0 = cadd(p0, p2), s1 = csub(p0, p2), s2 = cadd(p1, p3), s3 = csub(p1, p3);
  p0 = cadd(s0, s2);
  p2 = csub(s0, s2);
  p1 = make_float2(s1.x + s3.y, s1.y - s3.x);
  p3 = make_float2(s1.x - s3.y, s1.y + s3.x);
}

// Forward 16-point DFT, radix 4x4, fully in registers.  Input natural order v[n]; on return
// X[4c+d] is stored at v[c+4d]; use X16(v,k).
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128675613f;  // cos(pi/8)
  constexpr float S1 = 0.38268343236508977173f;  // sin(pi/8)
  constexpr float H = 0.70710678118654752440f;
#pragma unroll
  for (int a = 0; a < 4; ++a) fft4(v[a], v[a + 4], v[a + 8], v[a + 12]);
  // y[a][d] at v[a+4d]  *=  W16^(a*d)
  float2 x;
  x = v[1 + 4];  v[1 + 4]  = make_float2(x.x * C1 + x.y * S1, x.y * C1 - x.x * S1);      // W^1
  x = v[1 + 8];  v[1 + 8]  = make_float2((x.x + x.y) * H, (x.y - x.x) * H);              // W^2
  x = v[1 + 12]; v[1 + 12] = make_float2(x.x * S1 + x.y * C1, x.y * S1 - x.x * C1);      // W^3
  x = v[2 + 4];  v[2 + 4]  = make_float2((x.x + x.y) * H, (x.y - x.x) * H);              // W^2
  x = v[2 + 8];  v[2 + 8]  = make_float2(x.y, -x.x);                                     // W^4
  x = v[2 + 12]; v[2 + 12] = make_float2((x.y - x.x) * H, -(x.x + x.y) * H);             // W^6
  x = v[3 + 4];  v[3 + 4]  = make_float2(x.x * S1 + x.y * C1, x.y * S1 - x.x * C1);      // W^3
  x = v[3 + 8];  v[3 + 8]  = make_float2((x.y - x.x) * H, -(x.x + x.y) * H);             // W^6
  x = v[3 + 12]; v[3 + 12] = make_float2(-(x.x * C1 + x.y * S1), x.x * S1 - x.y * C1);   // W^9
#pragma unroll
  for (int d = 0; d < 4; ++d) fft4(v[4 * d], v[4 * d + 1], v[4 * d + 2], v[4 * d + 3]);
}
#define X16(v, k) (v)[((k) >> 2) + 4 * ((k) & 3)]

// exp(-2*pi*i*j/32), j = 0..7 (folded to immediates after unrolling).
__device__ __forceinline__ float2 w32(int j) {
  switch (j) {
    case 1: return make_float2(0.98078528040323044913f, -0.19509032201612826785f);
    case 2: return make_float2(0.92387953251128675613f, -0.38268343236508977173f);
    case 3: return make_float2(0.83146961230254523708f, -0.55557023301960222474f);
    case 4: return make_float2(0.70710678118654752440f, -0.70710678118654752440f);
    case 5: return make_float2(0.55557023301960222474f, -0.83146961230254523708f);
    case 6: return make_float2(0.38268343236508977173f, -0.92387953251128675613f);
    case 7: return make_float2(0.19509032201612826785f, -0.98078528040323044913f);
    default: return make_float2(1.0f, 0.0f);
  }
}

__device__ __forceinline__ void st_global_v4(float* p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct LogmelArgs {
  const float* wav;
  const int32_t* len;
  const float* peak;       // may be null when !normalize
  float* out;
  int32_t* n_frames;
  const float* hwin;
  const float2* tw256;
  const float2* tw512;
  const float4* band_w;
  const MelBands* bands;
  int64_t row_stride;
  int32_t B, T_max, tiles_per_row, total_tiles;
  int32_t normalize;
  float preemph, floor_, log_scale;
};

__global__ void __launch_bounds__(kThreads, 2) logmel_kernel(const LogmelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = lane & 15, half = lane >> 4;

  // ---- per-lane constants ----------------------------------------------------------------
  float2 hw[13];
#pragma unroll
  for (int m2 = 0; m2 < 13; ++m2) hw[m2] = *reinterpret_cast<const float2*>(a.hwin + 2 * (t + 16 * m2));
  float2 tw[16];
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) tw[k2] = a.tw256[(t * k2) & 255];
  const float2 base = a.tw512[t];
  const float2 w0 = (t == 0) ? make_float2(0.0f, -1.0f) : base;
  const int partner = (lane & 16) | ((16 - t) & 15);

  for (int i = tid; i < kMelBandMaxW4; i += kThreads) S.band_w[i] = a.band_w[i];
  {
    const int32_t* src = reinterpret_cast<const int32_t*>(a.bands);
    int32_t* dst = reinterpret_cast<int32_t*>(&S.bands);
    for (int i = tid; i < (int)(sizeof(MelBands) / 4); i += kThreads) dst[i] = src[i];
  }
  for (int i = tid; i < kTileFrames * 4; i += kThreads) S.P[(i >> 2) * kPStride + kBins + (i & 3)] = 0.0f;
  __syncthreads();

  float2* scr = S.scr + (warp * 2 + half) * kScrPerFrame;
  float* stage = reinterpret_cast<float*>(S.scr);

  for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_row;
    const int tf = tile - b * a.tiles_per_row;
    const int n = a.len[b];
    int Tb = (n >= kFrameLen) ? 1 + (n - kFrameLen) / kFrameStep : 0;
    Tb = min(Tb, a.T_max);
    if (tf == 0 && tid == 0) a.n_frames[b] = Tb;
    const int f0 = tf * kTileFrames;
    const int rows = min(kTileFrames, a.T_max - f0);
    const int nvalid = max(0, min(kTileFrames, Tb - f0));
    float* orow = a.out + ((size_t)b * a.T_max + f0) * kMel;

    if (nvalid == 0) {  // whole tile is collate padding
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = tid; i < rows * (kMel / 4); i += kThreads) st_global_v4(orow + 4 * i, z);
      continue;
    }

    // ---- stage the tile: gain, pre-emphasis (reference float32 op order), to shared --------
    {
      const float* row = a.wav + (size_t)b * a.row_stride;
      const int s0 = f0 * kFrameStep;
      const int count = (nvalid - 1) * kFrameStep + kFrameLen;  // multiple of 4, s0+count <= n
      float g = 1.0f;
      if (a.normalize) g = __fdiv_rn(1.0f, __fadd_rn(a.peak[b], 1e-9f));  // :70
      const float c = a.preemph;
      for (int i4 = tid; i4 < kWavSmem / 4; i4 += kThreads) {
        const int s = s0 + 4 * i4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool in = (4 * i4 < count);
        if (in) x = *reinterpret_cast<const float4*>(row + s);
        x.x = __fmul_rn(x.x, g); x.y = __fmul_rn(x.y, g); x.z = __fmul_rn(x.z, g); x.w = __fmul_rn(x.w, g);  // :71
        float xp = __shfl_up_sync(0xffffffffu, x.w, 1);
        if (lane == 0) xp = (in && s > 0) ? __fmul_rn(row[s - 1], g) : 0.0f;
        float4 y = x;
        if (c > 0.0f) {  // :75-79  y[0]=x[0]; y[n]=x[n]-c*x[n-1], product and difference rounded separately
          y.x = (s > 0) ? __fsub_rn(x.x, __fmul_rn(c, xp)) : x.x;
          y.y = __fsub_rn(x.y, __fmul_rn(c, x.x));
          y.z = __fsub_rn(x.z, __fmul_rn(c, x.y));
          y.w = __fsub_rn(x.w, __fmul_rn(c, x.z));
        }
        if (!in) y = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(S.wav + 4 * i4) = y;
      }
    }
    __syncthreads();

    // ---- FFT + power: two frames per warp per pass ------------------------------------------
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int fA = pass * 8 + warp;
      if (fA >= nvalid) continue;  // both of this warp's frames are padding (warp-uniform)
      const int fr = fA + 16 * half;
      const float* frp = S.wav + fr * kFrameStep + 2 * t;
      float2 v[16];
#pragma unroll
      for (int m2 = 0; m2 < 13; ++m2) {
        const float2 s = *reinterpret_cast<const float2*>(frp + 32 * m2);
        v[m2] = make_float2(s.x * hw[m2].x, s.y * hw[m2].y);
      }
      v[13] = v[14] = v[15] = make_float2(0.f, 0.f);
      fft16(v);
      // transpose twiddle W256^(t*k2) and scatter: scratch[k2][t]
      __syncwarp();
      scr[t] = X16(v, 0);
#pragma unroll
      for (int k2 = 1; k2 < 16; ++k2) scr[k2 * kScrStride + t] = cmul(X16(v, k2), tw[k2]);
      __syncwarp();
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) v[n1] = scr[t * kScrStride + n1];
      fft16(v);  // X16(v,k1) = Z[t + 16*k1] (half scaled)

      float* Prow = S.P + fr * kPStride;
      // real-FFT split: pairs (k, 256-k), k = t+16j, j=0..7; partner lane holds Z[256-k] at k1=15-j
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float2 za = X16(v, j);
        const float2 zq = X16(v, 15 - j);
        float2 zb = make_float2(__shfl_sync(0xffffffffu, zq.x, partner), __shfl_sync(0xffffffffu, zq.y, partner));
        if (t == 0) {  // residue 0 pairs with itself: (16j, 256-16j); slot j=0 takes the self-paired k=128
          if (j == 0) { za = X16(v, 8); zb = za; }
          else zb = X16(v, 16 - j);
        }
        const float er = za.x + zb.x, ei = za.y - zb.y;      // E' = Z[k] + conj(Z[256-k])
        const float dr = za.x - zb.x, di = za.y + zb.y;      // D  = Z[k] - conj(Z[256-k])
        float2 o = make_float2(di, -dr);                     // O' = -i*D
        if (j > 0) o = cmul(o, w32(j));
        const float2 tt = cmul(o, (j == 0) ? w0 : base);     // W512^k * O'
        const float ar = er + tt.x, ai = ei + tt.y;          // X[k]
        const float br = er - tt.x, bi = ei - tt.y;          // conj(X[256-k])
        const int ka = (j == 0) ? ((t == 0) ? 128 : t) : t + 16 * j;
        Prow[ka] = ar * ar + ai * ai;
        Prow[256 - ka] = br * br + bi * bi;
      }
      if (t == 0) {
        const float2 z0 = X16(v, 0);
        const float p = 2.0f * (z0.x + z0.y), q = 2.0f * (z0.x - z0.y);
        Prow[0] = p * p;
        Prow[256] = q * q;
      }
    }
    __syncthreads();

    // ---- mel projection + log: lane = frame, warp = mel bins {warp, warp+8, ...} ------------
    {
      const float* Prow = S.P + lane * kPStride;
#pragma unroll 1
      for (int m = warp; m < kMel; m += kWarps) {
        const int k0 = S.bands.k0[m], n4 = S.bands.n4[m];
        const float4* wp = S.band_w + S.bands.off4[m];
        const float* pp = Prow + k0;
        float acc = 0.0f;
        for (int i = 0; i < n4; ++i) {
          const float4 w = wp[i];
          acc = fmaf(pp[4 * i + 0], w.x, acc);
          acc = fmaf(pp[4 * i + 1], w.y, acc);
          acc = fmaf(pp[4 * i + 2], w.z, acc);
          acc = fmaf(pp[4 * i + 3], w.w, acc);
        }
        stage[lane * kOutStride + m] = __log2f(fmaxf(acc, a.floor_)) * a.log_scale;
      }
    }
    __syncthreads();

    // ---- coalesced store; rows beyond n_frames[b] are the collate's 0.0 ----------------------
    for (int i = tid; i < rows * (kMel / 4); i += kThreads) {
      const int r = i / (kMel / 4), m4 = (i - r * (kMel / 4)) * 4;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nvalid) {
        const float* sp = stage + r * kOutStride + m4;
        o = make_float4(sp[0], sp[1], sp[2], sp[3]);
      }
      st_global_v4(orow + 4 * i, o);
    }
    __syncthreads();  // stage (= scratch) and wav are reused by the next tile
  }
}

__global__ void nframes_kernel(const int32_t* __restrict__ len, int32_t B, int32_t T_max, int32_t* __restrict__ n_frames) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int n = len[b];
  int Tb = (n >= kFrameLen) ? 1 + (n - kFrameLen) / kFrameStep : 0;
  n_frames[b] = min(Tb, T_max);
}

}  // namespace

extern "C" int tasr_logmel_f32(const TasrFeaturizer* f, const float* wav, const int32_t* len,
                               const float* peak, int32_t B, int64_t row_stride, float* out,
                               int32_t T_max, int32_t* n_frames, tasr_stream_t stream) {
  if (!f || !wav || !len || !out || !n_frames) return fail(TASR_ERR_BAD_ARG, "tasr_logmel_f32: null argument");
  if (B < 0 || T_max < 0 || row_stride < 0) return fail(TASR_ERR_BAD_ARG, "tasr_logmel_f32: negative size");
  if (f->p.normalize_signal && !peak)
    return fail(TASR_ERR_BAD_ARG, "tasr_logmel_f32: normalize_signal is set but peak is NULL (run tasr_absmax_f32 first)");
  if (!aligned16(wav) || (row_stride & 3) || !aligned16(out))
    return fail(TASR_ERR_MISALIGNED, "tasr_logmel_f32: wav/out must be 16-byte aligned and row_stride a multiple of 4 samples");
  if (B == 0) return TASR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  TASR_CUDA(cudaGetDevice(&dev));
  if (dev != f->device) return fail(TASR_ERR_BAD_ARG, "tasr_logmel_f32: featurizer was created on device %d, current device is %d", f->device, dev);

  const int tiles_per_row = (T_max + kTileFrames - 1) / kTileFrames;
  const long long total = (long long)tiles_per_row * B;
  if (total > 0x7fffffffLL) return fail(TASR_ERR_UNSUPPORTED, "tasr_logmel_f32: too many tiles");
  if (total == 0) {
    nframes_kernel<<<(B + 127) / 128, 128, 0, st>>>(len, B, T_max, n_frames);
    TASR_LAUNCH_CHECK("nframes_kernel");
    return TASR_OK;
  }
  static bool attr_set[64] = {false};
  const size_t smem = sizeof(Smem);
  if (dev < 64 && !attr_set[dev]) {
    TASR_CUDA(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[dev] = true;
  }
  LogmelArgs a;
  a.wav = wav; a.len = len; a.peak = peak; a.out = out; a.n_frames = n_frames;
  a.hwin = f->d_hwin; a.tw256 = f->d_tw256; a.tw512 = f->d_tw512; a.band_w = f->d_band_w; a.bands = f->d_bands;
  a.row_stride = row_stride; a.B = B; a.T_max = T_max; a.tiles_per_row = tiles_per_row; a.total_tiles = (int)total;
  a.normalize = f->p.normalize_signal ? 1 : 0;
  a.preemph = f->p.preemphasis; a.floor_ = f->p.output_floor; a.log_scale = f->log_scale;
  const int grid = (int)((total < (long long)2 * sm_count()) ? total : (long long)2 * sm_count());
  logmel_kernel<<<grid, kThreads, smem, st>>>(a);
  TASR_LAUNCH_CHECK("logmel_kernel");
  return TASR_OK;
}
