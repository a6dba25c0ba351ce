// Fused waveform -> log-mel kernel (sm_100a).
//
// Replaces, per utterance, src/speech_featurizer.py:136-161 (normalize_signal ->
// preemphasis_signal -> tf.signal.stft(400/160, periodic Hann, rFFT-512) -> |X|^2 -> HTK mel
// matmul -> log10(max(.,1e-9))) and the zero-padded collate of src/dataset.py:236-252.
//
// Work decomposition
//   items = (a) the VALID 32-frame tiles of every utterance, dealt round-robin to the persistent
//           CTAs (every CTA gets the same number of them +-1: a ragged batch stays balanced), and
//           (b) the collate padding (rows t >= n_frames[b]), zero-filled in 128-row chunks that are
//           dealt round-robin the same way, first, so the stores drain under the FFT work;
//   tile  = 32 consecutive frames of one utterance (5360 samples, staged once in shared memory
//           with gain and pre-emphasis applied in exactly the reference's float32 op order; the
//           next tile's samples are prefetched into L2 while this one computes);
//   FFT   = 16 lanes per frame (two frames per warp).  The 512-point real FFT is a 256-point
//           complex FFT of z[m] = y[2m] + i*y[2m+1] done as 16x16: radix-16 in registers,
//           twiddle, 16x16 transpose through a padded per-warp scratch, radix-16 again; then the
//           real-FFT split, where lane t and lane 16-t exchange eight values by warp shuffle and
//           each forms |X[k]|^2 and |X[256-k]|^2 for its eight k;
//   mel   = lane <-> frame, warp <-> a contiguous group of mel bins.  For the config/model.yaml
//           filterbank the sparsity structure is compiled in (mel_geometry.inc): every power bin
//           is loaded once and feeds its two adjacent triangles with weights read straight from
//           the kernel-parameter constant bank, fully unrolled (3 instructions per FFT bin).  Any
//           other triangular filterbank takes the generic banded loop;
//   out   = log, staged through shared memory, written with coalesced 128-bit stores.
//
// Per-lane constants (half-window, transpose twiddles) live in registers for the whole
// persistent loop.  Twiddles are float64-derived tables.
#include "logmel_common.cuh"
#include <stdlib.h>
#include <math.h>

using namespace tasr;

using namespace tasr_lm;

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kWavSmem = 5376;          // (32-1)*160+400 = 5360, +16 floats lanes 8..15 touch at m2=12
constexpr int kWavSlots = kWavSmem / 4 / kThreads + 1;   // float4 slots per thread (6; 1344 = 5.25 * 256)
constexpr int kScrStride = 17;          // float2 units; odd -> conflict-free transposed reads
constexpr int kScrPerFrame = 16 * kScrStride;
constexpr int kPStride = kBins + 4;     // 261, odd; columns 257..260 stay zero (band padding of the generic path)
constexpr int kChunkUtt = 256;          // utterances whose work items are indexed at a time (one per thread)
constexpr int kListN = 32;              // located tiles per refill of the shared list
constexpr int kRawFloats = 5368;        // TMA landing buffer of the next tile: 4 floats ahead of the tile's first sample + 5360

struct __align__(16) Smem {
  float wav[kWavSmem];
  float2 scr[kWarps * 2 * kScrPerFrame];   // also the [32][81] output staging tile
  float P[kTileFrames * kPStride];
  float2 tw512[136];                       // W512^k, k = 0..128
  int32_t vcum[kChunkUtt + 1];             // exclusive prefix of valid 32-frame tiles per utterance of the chunk
  int32_t pcum[kChunkUtt + 1];             // exclusive prefix of 128-row padding chunks per utterance
  int32_t wsum[2][kWarps];
  int4 list[kListN];                       // this CTA's next valid tiles: (utterance within chunk, tile index, len, -)
  unsigned long long bar;                  // mbarrier of the raw-sample copies
  union {
    struct {
      float4 band_w[kMelBandMaxW4];        // generic mel path only (no TMA staging there)
      MelBands bands;
    } gen;
    float raw[kRawFloats];                 // fixed mel path: the next tile's raw samples, written by cp.async.bulk (TMA)
  };
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

// ---- TMA (cp.async.bulk) staging of the raw samples ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void raw_bar_init(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void raw_bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
// One thread: bulk-copy the samples a tile stages — [s0 - 4, s0 + count) of the utterance's row (the 4 floats ahead carry the
// pre-emphasis neighbour of the first sample; none for the utterance's first tile) — so that sample s0 lands at raw[4].
__device__ __forceinline__ void raw_issue(const LogmelArgs& a, float* raw, uint32_t bar, int b, int tf, int n) {
  const int Tb = frames_of(n, a);
  const int f0 = tf * kTileFrames;
  const int nvalid = min(kTileFrames, Tb - f0);
  const int s0 = f0 * kFrameStep;
  const int count = (nvalid - 1) * kFrameStep + kFrameLen;
  const int pre = s0 > 0 ? 4 : 0;
  const uint32_t bytes = (uint32_t)(count + pre) * 4u;
  const float* src = a.wav + (size_t)b * a.row_stride + (s0 - pre);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(raw + 4 - pre)), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <bool FIXED>
__global__ void __launch_bounds__(kThreads, 2)
logmel_kernel(const __grid_constant__ LogmelArgs a, const __grid_constant__ MelFixedW mw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = lane & 15, half = lane >> 4;

  // ---- n_frames (src/speech_featurizer.py:163-166) -------------------------------------------
  for (int b = blockIdx.x * kThreads + tid; b < a.B; b += gridDim.x * kThreads) a.n_frames[b] = frames_of(a.len[b], a);

  // ---- per-lane constants ----------------------------------------------------------------
  float2 hw[13];
#pragma unroll
  for (int m2 = 0; m2 < 13; ++m2) hw[m2] = *reinterpret_cast<const float2*>(a.hwin + 2 * (t + 16 * m2));
  float2 tw[16];
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) tw[k2] = a.tw256[(t * k2) & 255];
  const int partner = (lane & 16) | ((16 - t) & 15);

  // TMA staging (fixed mel path, frames fully inside the signal): the next tile's raw samples are bulk-copied into S.raw while this
  // tile computes; the staging pass then reads them from shared memory instead of waiting for global loads.
  const bool use_tma = FIXED && !a.pad_end && a.tma != 0;
  const uint32_t raw_bar = smem_addr(&S.bar);
  uint32_t raw_phase = 0;
  if (use_tma && tid == 0) raw_bar_init(raw_bar);
  for (int i = tid; i <= 128; i += kThreads) S.tw512[i] = a.tw512[i];
  if (!FIXED) {
    for (int i = tid; i < kMelBandMaxW4; i += kThreads) S.gen.band_w[i] = a.band_w[i];
    const int32_t* src = reinterpret_cast<const int32_t*>(a.bands);
    int32_t* dst = reinterpret_cast<int32_t*>(&S.gen.bands);
    for (int i = tid; i < (int)(sizeof(MelBands) / 4); i += kThreads) dst[i] = src[i];
    for (int i = tid; i < kTileFrames * 4; i += kThreads) S.P[(i >> 2) * kPStride + kBins + (i & 3)] = 0.0f;
  }
  __syncthreads();

  float2* scr = S.scr + (warp * 2 + half) * kScrPerFrame;
  float* stage = reinterpret_cast<float*>(S.scr);
  const float2* twp = S.tw512 + t;           // W512^(t+16j) at twp[16j]; lane t=0 uses W512^128 for j=0
  const int tw0 = (t == 0) ? 128 : 0;

  // Work items are indexed per chunk of kChunkUtt utterances: two block-wide prefix sums (valid tiles,
  // padding chunks) in shared memory, then item j -> (utterance, index) by binary search.  jv / jp are
  // this CTA's next global item indices; they keep striding by gridDim.x across chunks.
  int jv = blockIdx.x, jp = blockIdx.x, voff = 0, poff = 0;
#pragma unroll 1
  for (int cb = 0; cb < a.B; cb += kChunkUtt) {
  const int nu = min(kChunkUtt, a.B - cb);
  {
    constexpr int kPer = kChunkUtt / kThreads;     // utterances per thread
    static_assert(kPer >= 1 && kPer * kThreads == kChunkUtt, "kChunkUtt must be a multiple of the block size");
    int vt[kPer], pt[kPer], vs = 0, ps = 0;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int u = kPer * tid + i;
      vt[i] = pt[i] = 0;
      if (u < nu) {
        const int Tu = frames_of(a.len[cb + u], a);
        vt[i] = (Tu + kTileFrames - 1) / kTileFrames;
        const int pad_rows = pad_limit(Tu, a) - vt[i] * kTileFrames;
        pt[i] = pad_rows > 0 ? (pad_rows + kPadChunkRows - 1) / kPadChunkRows : 0;
      }
      vs += vt[i]; ps += pt[i];
    }
    int vi = vs, pi = ps;   // inclusive scan over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v2 = __shfl_up_sync(0xffffffffu, vi, d), p2 = __shfl_up_sync(0xffffffffu, pi, d);
      if (lane >= d) { vi += v2; pi += p2; }
    }
    if (lane == 31) { S.wsum[0][warp] = vi; S.wsum[1][warp] = pi; }
    __syncthreads();
    int vb = vi - vs, pb = pi - ps;
    for (int w = 0; w < warp; ++w) { vb += S.wsum[0][w]; pb += S.wsum[1][w]; }
    if (tid == 0) { S.vcum[0] = 0; S.pcum[0] = 0; }
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      vb += vt[i]; pb += pt[i];
      S.vcum[kPer * tid + i + 1] = vb;
      S.pcum[kPer * tid + i + 1] = pb;
    }
    __syncthreads();
  }
  const int vtot = S.vcum[nu], ptot = S.pcum[nu];
  auto find = [&](const int32_t* cum, int x) -> int {   // largest u in [0,nu) with cum[u] <= x
    int lo = 0, hi = nu;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (cum[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
  };

  // ---- (b) collate padding: rows beyond the last valid tile of every utterance, 128-row chunks ----
  for (; jp < poff + ptot; jp += gridDim.x) {
    const int u = find(S.pcum, jp - poff);
    const int vt = S.vcum[u + 1] - S.vcum[u];
    const int r0 = vt * kTileFrames + (jp - poff - S.pcum[u]) * kPadChunkRows;
    const int rows = min(kPadChunkRows, pad_limit(frames_of(a.len[cb + u], a), a) - r0);
    float* dst = a.out + ((size_t)(cb + u) * a.T_max + r0) * kMel;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < rows * (kMel / 4); i += kThreads) st_global_v4(dst + 4 * i, z);
  }

  // ---- (a) valid tiles, round-robin ------------------------------------------------------------
  // This CTA's tiles of the chunk are jv, jv+grid, ...; they are located once, one per thread, into a
  // shared list (utterance, tile index) that the tile loop then just reads.
#pragma unroll 1
  while (jv < voff + vtot) {
  const int nlist = min(kListN, (voff + vtot - jv + (int)gridDim.x - 1) / (int)gridDim.x);
  if (tid < nlist) {
    const int x = jv - voff + tid * (int)gridDim.x;
    const int u = find(S.vcum, x);
    S.list[tid] = make_int4(u, x - S.vcum[u], a.len[cb + u], 0);   // the length rides along: no global load in the tile loop
  }
  __syncthreads();
  if (use_tma && tid == 0) { const int4 it0 = S.list[0]; raw_issue(a, S.raw, raw_bar, cb + it0.x, it0.y, it0.z); }
#pragma unroll 1
  for (int li = 0; li < nlist; ++li) {
    const int4 item = S.list[li];
    const int b = cb + item.x;
    const int tf = item.y;

    if (!use_tma && li + 1 < nlist && tid < 168) {  // next tile -> L2: 5360 samples = 167.5 lines of 128 B
      const int4 nxt = S.list[li + 1];
      const float* nrow = a.wav + (size_t)(cb + nxt.x) * a.row_stride;
      const int ns = nxt.y * kTileFrames * kFrameStep + tid * 32;
      if (ns < nxt.z) prefetch_l2(nrow + ns);
    }

    const int n = item.z;
    const int Tb = frames_of(n, a);
    const int f0 = tf * kTileFrames;
    const int rows = min(kTileFrames, a.T_max - f0);
    const int nvalid = min(kTileFrames, Tb - f0);   // >= 1: only valid tiles are enumerated
    float* orow = a.out + ((size_t)b * a.T_max + f0) * kMel;

    // a NaN peak (a NaN sample somewhere in the utterance) makes every feature of the utterance NaN, as in the
    // reference; fmaxf(NaN, floor) alone would return the floor
    float floor_b = a.floor_;
    // ---- stage the tile: gain, pre-emphasis (reference float32 op order), to shared --------
    {
      const float* row = a.wav + (size_t)b * a.row_stride;
      const int s0 = f0 * kFrameStep;
      const int count = (nvalid - 1) * kFrameStep + kFrameLen;  // multiple of 4; s0+count <= n unless pad_end
      const int lim = a.pad_end ? min(count, n - s0) : count;   // samples of the tile that exist
      float g = 1.0f;
      if (a.normalize && a.peak_out == nullptr) g = __fdiv_rn(1.0f, __fadd_rn(a.peak[b], 1e-9f));  // :70
      if (g != g) floor_b = g;
      const float c = a.preemph;
      float4 x[kWavSlots];
      float xp[kWavSlots];
      if (use_tma) {
        raw_bar_wait(raw_bar, raw_phase);       // this tile's samples have landed in S.raw (copied while the previous tile computed)
        raw_phase ^= 1u;
#pragma unroll
        for (int u = 0; u < kWavSlots; ++u) {
          const int i4 = tid + u * kThreads;
          x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          xp[u] = 0.0f;
          if (4 * i4 < lim) {
            x[u] = *reinterpret_cast<const float4*>(S.raw + 4 + 4 * i4);
            if (s0 + 4 * i4 > 0) xp[u] = S.raw[3 + 4 * i4];
          }
        }
      } else {
#pragma unroll
      for (int u = 0; u < kWavSlots; ++u) {     // all loads first: 6 x 128-bit + 6 x 32-bit in flight per thread
        const int i4 = tid + u * kThreads;
        const int s = s0 + 4 * i4;
        x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        xp[u] = 0.0f;
        if (4 * i4 < lim) {
          x[u] = *reinterpret_cast<const float4*>(row + s);   // (rows are padded to 4 samples: in bounds)
          if (s > 0) xp[u] = row[s - 1];
        }
      }
      }
      if (a.peak_out != nullptr) {
        // Single pass: max|x| of the samples this tile stages (slots beyond `lim` hold zeros); the tile that holds the
        // utterance's last frame also takes the samples no frame covers, [s0+count, n).  Non-negative floats order like
        // their bit patterns, so the warp maximum is one integer REDUX and the merge one atomicMax per warp.
        unsigned m = 0u;   // integer max over the bit patterns of |x|: finite values order like floats, a NaN sample wins
#pragma unroll
        for (int u = 0; u < kWavSlots; ++u)
          m = max(max(m, max(abs_bits(x[u].x), abs_bits(x[u].y))), max(abs_bits(x[u].z), abs_bits(x[u].w)));
        if (f0 + nvalid >= Tb) {
          for (int i = s0 + count + 4 * tid; i < n; i += 4 * kThreads) {
            const float4 t4 = *reinterpret_cast<const float4*>(row + i);
            m = max(m, abs_bits(t4.x));
            if (i + 1 < n) m = max(m, abs_bits(t4.y));
            if (i + 2 < n) m = max(m, abs_bits(t4.z));
            if (i + 3 < n) m = max(m, abs_bits(t4.w));
          }
        }
        const unsigned mb = __reduce_max_sync(0xffffffffu, m);
        if (lane == 0 && mb != 0u) atomicMax(reinterpret_cast<unsigned*>(a.peak_out) + b, mb);
      }
#pragma unroll
      for (int u = 0; u < kWavSlots; ++u) {
        const int i4 = tid + u * kThreads;
        if (4 * i4 < count + 16 && i4 < kWavSmem / 4) {     // [count, count+16) must be finite zeros (window tail)
          float4 v = x[u];
          v.x = __fmul_rn(v.x, g); v.y = __fmul_rn(v.y, g); v.z = __fmul_rn(v.z, g); v.w = __fmul_rn(v.w, g);  // :71
          float4 y = v;
          if (c > 0.0f) {  // :75-79  y[0]=x[0]; y[n]=x[n]-c*x[n-1], product and difference rounded separately
            const float vp = __fmul_rn(xp[u], g);
            y.x = (s0 + 4 * i4 > 0) ? __fsub_rn(v.x, __fmul_rn(c, vp)) : v.x;
            y.y = __fsub_rn(v.y, __fmul_rn(c, v.x));
            y.z = __fsub_rn(v.z, __fmul_rn(c, v.y));
            y.w = __fsub_rn(v.w, __fmul_rn(c, v.z));
          }
          if (a.pad_end) {   // the zero padding is appended AFTER pre-emphasis (tf.signal.frame pads the signal it is given)
            if (4 * i4 + 0 >= lim) y.x = 0.0f;
            if (4 * i4 + 1 >= lim) y.y = 0.0f;
            if (4 * i4 + 2 >= lim) y.z = 0.0f;
            if (4 * i4 + 3 >= lim) y.w = 0.0f;
          }
          *reinterpret_cast<float4*>(S.wav + 4 * i4) = y;
        }
      }
    }
    __syncthreads();
    if (use_tma && tid == 0 && li + 1 < nlist) {   // every thread has its slots of S.raw in registers: the next tile's copy may overwrite it
      const int4 nxt = S.list[li + 1];
      raw_issue(a, S.raw, raw_bar, cb + nxt.x, nxt.y, nxt.z);
    }

    // ---- FFT + power: two frames per warp per pass ------------------------------------------
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int fA = pass * 8 + warp;
      if (fA >= nvalid) continue;  // both of this warp's frames are padding (warp-uniform)
      const int fr = fA + 16 * half;
      // frame fA+16 may be beyond nvalid: its samples are then stale but finite (a previous tile's, or the
      // zero-initialised buffer) and its P row is never read by a valid output row.
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

      float* Pa = S.P + fr * kPStride + t;          // P[k],     k = t + 16j
      float* Pb = S.P + fr * kPStride + 256 - t;    // P[256-k]
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
        const float2 wk = (j == 0) ? twp[tw0] : twp[16 * j];  // W512^k
        const float2 tt = cmul(make_float2(di, -dr), wk);    // W512^k * (-i*D)
        const float ar = er + tt.x, ai = ei + tt.y;          // X[k]
        const float br = er - tt.x, bi = ei - tt.y;          // conj(X[256-k])
        const float pa = ar * ar + ai * ai, pb = br * br + bi * bi;
        if (j == 0) {
          const int ka = (t == 0) ? 128 : 0;                 // lane 0: k = 128 (both stores hit P[128])
          Pa[ka] = pa;
          Pb[-ka] = pb;
        } else {
          Pa[16 * j] = pa;
          Pb[-16 * j] = pb;
        }
      }
      if (t == 0) {
        const float2 z0 = X16(v, 0);
        const float p = 2.0f * (z0.x + z0.y), q = 2.0f * (z0.x - z0.y);
        Pa[0] = p * p;
        Pb[0] = q * q;
      }
    }
    __syncthreads();

    // ---- mel projection + log: lane = frame ---------------------------------------------------
    {
      const float* Prow = S.P + lane * kPStride;
      float* srow = stage + lane * kOutStride;
      if (a.mode == 1) {   // "spectrogram": log power of the first 80 FFT bins (src/speech_featurizer.py:124-126)
        for (int k = warp; k < kMel; k += kWarps) srow[k] = lg2_normal(fmaxf(Prow[k], floor_b)) * a.log_scale;
      } else if (FIXED) {
        switch (warp) {
          case 0: mel_fixed_group<0>(Prow, mw, srow, floor_b, a.log_scale); break;
          case 1: mel_fixed_group<1>(Prow, mw, srow, floor_b, a.log_scale); break;
          case 2: mel_fixed_group<2>(Prow, mw, srow, floor_b, a.log_scale); break;
          case 3: mel_fixed_group<3>(Prow, mw, srow, floor_b, a.log_scale); break;
          case 4: mel_fixed_group<4>(Prow, mw, srow, floor_b, a.log_scale); break;
          case 5: mel_fixed_group<5>(Prow, mw, srow, floor_b, a.log_scale); break;
          case 6: mel_fixed_group<6>(Prow, mw, srow, floor_b, a.log_scale); break;
          default: mel_fixed_group<7>(Prow, mw, srow, floor_b, a.log_scale); break;
        }
      } else {
#pragma unroll 1
        for (int m = warp; m < kMel; m += kWarps) {
          const int k0 = S.gen.bands.k0[m], n4 = S.gen.bands.n4[m];
          const float4* wp = S.gen.band_w + S.gen.bands.off4[m];
          const float* pp = Prow + k0;
          float acc = 0.0f;
          for (int i = 0; i < n4; ++i) {
            const float4 w = wp[i];
            acc = fmaf(pp[4 * i + 0], w.x, acc);
            acc = fmaf(pp[4 * i + 1], w.y, acc);
            acc = fmaf(pp[4 * i + 2], w.z, acc);
            acc = fmaf(pp[4 * i + 3], w.w, acc);
          }
          srow[m] = lg2_normal(fmaxf(acc, floor_b)) * a.log_scale;
        }
      }
    }
    __syncthreads();

    // ---- coalesced store; rows beyond n_frames[b] inside this tile are the collate's 0.0 ---------
    for (int i = tid; i < rows * (kMel / 4); i += kThreads) {
      const int r = i / (kMel / 4), m4 = (i - r * (kMel / 4)) * 4;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nvalid) {
        const float* sp = stage + r * kOutStride + m4;
        o = make_float4(sp[0], sp[1], sp[2], sp[3]);
      }
      st_global_v4(orow + 4 * i, o);
    }
    // No barrier here: the next tile first writes only `wav` (last read before the post-FFT barrier above), and its
    // staging barrier orders these reads of `stage` before the next FFT phase overwrites the scratch.
  }
  jv += nlist * (int)gridDim.x;   // (the tile loop ends on a barrier, so the list can be rewritten)
  }
  voff += vtot; poff += ptot;
  __syncthreads();   // the prefix tables are rebuilt for the next chunk
  }
}

}  // namespace

static int logmel_launch(const TasrFeaturizer* f, const float* wav, const int32_t* len,
                         const float* peak, int32_t B, int64_t row_stride, float* out,
                         int32_t T_max, int32_t* n_frames, int32_t pad_fill_rows, tasr_stream_t stream,
                         float* peak_out = nullptr) {
  if (!f || !wav || !len || !out || !n_frames) return fail(TASR_ERR_BAD_ARG, "tasr_logmel_f32: null argument");
  if (B < 0 || T_max < 0 || row_stride < 0) return fail(TASR_ERR_BAD_ARG, "tasr_logmel_f32: negative size");
  if (f->p.normalize_signal && !peak && !peak_out)
    return fail(TASR_ERR_BAD_ARG, "tasr_logmel_f32: normalize_signal is set but peak is NULL (run tasr_absmax_f32 first)");
  if (!aligned16(wav) || (row_stride & 3) || !aligned16(out))
    return fail(TASR_ERR_MISALIGNED, "tasr_logmel_f32: wav/out must be 16-byte aligned and row_stride a multiple of 4 samples");
  if (f->p.feature_type == TASR_FEAT_WAVEFORM)
    return fail(TASR_ERR_BAD_ARG, "tasr_logmel_f32: the handle's feature_type is 'waveform'; call tasr_waveform_f32");
  if (B == 0) return TASR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  TASR_CUDA(cudaGetDevice(&dev));
  if (dev != f->device) return fail(TASR_ERR_BAD_ARG, "tasr_logmel_f32: featurizer was created on device %d, current device is %d", f->device, dev);
  if (f->generic) {     // any other frame geometry: the general kernel (writes every padding row; no single-pass mode)
    if (peak_out) return fail(TASR_ERR_UNSUPPORTED, "tasr_logmel_f32_single_pass: built for the 400/160/512/80 geometry only");
    return tasr_logmel_generic_launch(f, wav, len, peak, B, row_stride, out, T_max, n_frames, st);
  }

  const int tiles_per_row = (T_max + kTileFrames - 1) / kTileFrames;
  const long long total = (long long)tiles_per_row * B;
  if (total > 0x7fffffffLL) return fail(TASR_ERR_UNSUPPORTED, "tasr_logmel_f32: too many tiles");
  static bool attr_set[64] = {false};
  const size_t smem = sizeof(Smem);
  if (dev < 64 && !attr_set[dev]) {
    TASR_CUDA(cudaFuncSetAttribute(logmel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TASR_CUDA(cudaFuncSetAttribute(logmel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[dev] = true;
  }
  LogmelArgs a;
  a.wav = wav; a.len = len; a.peak = peak; a.out = out; a.n_frames = n_frames; a.peak_out = peak_out;
  a.hwin = f->d_hwin; a.tw256 = f->d_tw256; a.tw512 = f->d_tw512; a.band_w = f->d_band_w; a.bands = f->d_bands;
  a.row_stride = row_stride; a.B = B; a.T_max = T_max; a.tiles_per_row = tiles_per_row;
  a.normalize = f->p.normalize_signal ? 1 : 0;
  a.pad_end = f->p.pad_end ? 1 : 0;
  a.mode = (f->p.feature_type == TASR_FEAT_SPECTROGRAM) ? 1 : 0;
  a.pad_fill_rows = pad_fill_rows;
  { const char* e = getenv("TASR_LOGMEL_TMA"); a.tma = (e && atoi(e) == 0) ? 0 : 1; }
  a.preemph = f->p.preemphasis; a.floor_ = f->p.output_floor; a.log_scale = f->log_scale;
  if (peak_out) {
    a.floor_ = 0.0f;   // the floor is applied by the reader, after the gain (lg2(0) = -inf survives the addition)
    TASR_CUDA(cudaMemsetAsync(peak_out, 0, (size_t)B * sizeof(float), st));
  }
  {
    // default: second FFT stage on the tensor cores (logmel_tc.cu) when the handle has the compiled-in filterbank
    const int rc_tc = tasr_logmel_tc_launch(f, a, st);
    if (rc_tc >= 0) {
      if (rc_tc != TASR_OK) return rc_tc;
      if (f->p.feature_type == TASR_FEAT_MFCC || f->p.normalize_zscore || f->p.normalize_min_max)
        return tasr_feature_post_launch(f, out, n_frames, B, T_max, st);
      return TASR_OK;
    }
  }
  // Persistent grid: two CTAs per SM; never more CTAs than work items (valid tiles + padding chunks <= total + B).
  const long long cap = total + B;
  static const int ctas_per_sm = [] { const char* e = getenv("TASR_LOGMEL_CTAS_PER_SM"); const int v = e ? atoi(e) : 2; return v >= 1 && v <= 2 ? v : 2; }();
  const long long full = (long long)ctas_per_sm * sm_count();
  const int grid = (int)((cap < full) ? (cap > 0 ? cap : 1) : full);
  if (f->mel_fixed) {
    const MelFixedW* mw = reinterpret_cast<const MelFixedW*>(f->mel_fixed_w);
    logmel_kernel<true><<<grid, kThreads, smem, st>>>(a, *mw);
  } else {
    static const MelFixedW zero_w = {};
    logmel_kernel<false><<<grid, kThreads, smem, st>>>(a, zero_w);
  }
  TASR_LAUNCH_CHECK("logmel_kernel");
  if (f->p.feature_type == TASR_FEAT_MFCC || f->p.normalize_zscore || f->p.normalize_min_max)
    return tasr_feature_post_launch(f, out, n_frames, B, T_max, st);
  return TASR_OK;
}

extern "C" int tasr_logmel_f32(const TasrFeaturizer* f, const float* wav, const int32_t* len,
                               const float* peak, int32_t B, int64_t row_stride, float* out,
                               int32_t T_max, int32_t* n_frames, tasr_stream_t stream) {
  return logmel_launch(f, wav, len, peak, B, row_stride, out, T_max, n_frames, -1, stream);
}

extern "C" int tasr_logmel_f32_lean(const TasrFeaturizer* f, const float* wav, const int32_t* len,
                                    const float* peak, int32_t B, int64_t row_stride, float* out,
                                    int32_t T_max, int32_t* n_frames, int32_t pad_fill_rows, tasr_stream_t stream) {
  if (pad_fill_rows < 0) return fail(TASR_ERR_BAD_ARG, "tasr_logmel_f32_lean: pad_fill_rows must be >= 0 (use tasr_logmel_f32 to write every row)");
  if (f && (f->p.feature_type == TASR_FEAT_MFCC || f->p.normalize_zscore || f->p.normalize_min_max))
    return fail(TASR_ERR_UNSUPPORTED, "tasr_logmel_f32_lean: mfcc / per-frame normalisation post-process the whole tensor; use tasr_logmel_f32");
  return logmel_launch(f, wav, len, peak, B, row_stride, out, T_max, n_frames, pad_fill_rows, stream);
}

extern "C" int tasr_logmel_f32_single_pass(const TasrFeaturizer* f, const float* wav, const int32_t* len, int32_t B,
                                           int64_t row_stride, float* out, int32_t T_max, int32_t* n_frames,
                                           int32_t pad_fill_rows, float* peak_out, TasrDeferredGain* gain_host,
                                           tasr_stream_t stream) {
  if (!f || !peak_out || !gain_host) return fail(TASR_ERR_BAD_ARG, "tasr_logmel_f32_single_pass: null argument");
  if (!f->p.normalize_signal)
    return fail(TASR_ERR_BAD_ARG, "tasr_logmel_f32_single_pass: the handle does not normalise the signal; use tasr_logmel_f32");
  if (f->p.pad_end || f->p.feature_type == TASR_FEAT_MFCC || f->p.normalize_zscore || f->p.normalize_min_max)
    return fail(TASR_ERR_UNSUPPORTED, "tasr_logmel_f32_single_pass: pad_end / mfcc / per-frame normalisation need the two-pass tasr_logmel_f32");
  gain_host->peak = peak_out;
  gain_host->log_scale_x2 = 2.0f * f->log_scale;
  gain_host->log_floor = f->p.log_base_e ? logf(f->p.output_floor) : log10f(f->p.output_floor);
  return logmel_launch(f, wav, len, nullptr, B, row_stride, out, T_max, n_frames, pad_fill_rows, stream, peak_out);
}

// Host side of the fixed-geometry check (called by tasr_featurizer_create): fills wr/wf from the dense
// [257,80] matrix when its sparsity structure is the compiled-in one, else returns false.
bool tasr_mel_fixed_from_dense(const float* mel_w_host, float* wr_wf_512) {
  MelFixedW w = {};
  if (kMelSegStart[0] != 1 || kMelSegStart[81] != 256) return false;
  for (int k = 0; k < kBins; ++k) {
    int j0 = -1, j1 = -1;
    for (int m = 0; m < kMel; ++m)
      if (mel_w_host[k * kMel + m] != 0.0f) { if (j0 < 0) j0 = m; j1 = m; }
    if (k == 0 || k == 256) { if (j0 >= 0) return false; continue; }
    if (j0 < 0 || j1 - j0 > 1) return false;
    const int seg = j0 + 1;                                  // 1..80
    if (!(k >= kMelSegStart[seg] && k < kMelSegStart[seg + 1])) return false;
    w.wf[k] = mel_w_host[k * kMel + j0];
    w.wr[k] = (seg < kMel) ? mel_w_host[k * kMel + seg] : 0.0f;
  }
  for (int k = 0; k < 256; ++k) { wr_wf_512[k] = w.wr[k]; wr_wf_512[256 + k] = w.wf[k]; }
  return true;
}
