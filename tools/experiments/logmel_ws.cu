// EXPERIMENT (not built): warp-specialised log-mel kernel — one 24-warp CTA per SM, 16 FFT warps running back
// to back, 4 mel warps and 4 I/O warps on other tiles, mbarrier hand-offs with two buffers each.  Correct (all GPU
// parity tests pass with TASR_LOGMEL_WS=1 when it is wired into tasr_logmel_f32) but SLOWER than logmel.cu on B200 /
// config 3: 181 us against 154 us.  Why: with 24 resident warps the register budget is 85 per thread, so the
// per-lane window halves and transpose twiddles (56 registers in logmel.cu) have to come from shared memory — 28
// extra LDS.64 per frame-lane on top of ~77 shared-memory operations, in the one phase that already saturates the
// shared-memory pipe; a first version with 8 FP32x2-packed FFT warps (two frames per register pair) took 190 us
// because the FFT needs sixteen warps' worth of latency hiding.  To try it: copy next to logmel.cu, add it to
// build.py SOURCES and call tasr_logmel_ws_launch from tasr_logmel_f32 (see git history, commit "logmel_ws").
// Fused waveform -> log-mel kernel, warp-specialised variant (sm_100a).  Same arithmetic contract as logmel.cu
// (tests hold both to the same tolerances); what changes is the schedule and the FP instruction form.
//
// logmel.cu runs two 8-warp CTAs per SM whose warps walk the phases of a 32-frame tile together (stage ->
// barrier -> FFT -> barrier -> mel -> barrier -> store).  Ablations show where its time goes: the FFT phase is
// bound by the shared-memory pipe (~10 KB of shared traffic per frame), the other phases leave that pipe idle,
// and nothing overlaps inside a CTA.  Here ONE 24-warp CTA per SM keeps three roles busy on different tiles at
// the same time, handing tiles over through mbarriers (two buffers per hand-off):
//
//   warps 20-23  I/O: stage the next tile's samples (gain, pre-emphasis in the reference's float32 op order)
//                into the packed sample buffer, write the previous tile's 32x80 outputs with coalesced 128-bit
//                stores, prefetch the tile after next into L2, zero-fill the collate padding in between;
//   warps 0-15   FFT: 16 lanes per frame, two frames per warp, one whole tile per pass: 16x16 four-step 256-point
//                complex FFT through a per-frame shared transpose tile, real-FFT split by warp shuffle, |X|^2
//                rows into the power buffer — the same sixteen FFT warps per SM as logmel.cu has in total, but
//                running back to back instead of sharing their time with the other phases;
//   warps 16-19  mel: lane <-> frame, the unrolled fixed-geometry projection with constant-bank weights (two
//                of the eight bin groups per warp), log, into the output staging tile.
//
// Work distribution as in logmel.cu: valid tiles enumerated through a prefix sum over the utterances and dealt
// round-robin to the CTAs; every role walks the same per-CTA tile list.
#include "logmel_common.cuh"

using namespace tasr;
using namespace tasr_lm;

namespace {

constexpr int kFftWarps = 16;
constexpr int kMelWarps = 4;
constexpr int kIoWarps = 4;
constexpr int kThreadsWs = (kFftWarps + kMelWarps + kIoWarps) * 32;   // 512
constexpr int kIoThreads = kIoWarps * 32;
constexpr int kWavSmem = 5376;                         // (32-1)*160+400 = 5360, +16 floats of window tail
constexpr int kIoSlots = 12;                            // float4 groups per I/O thread (1344 = 10.5 * 128), two batches of six
constexpr int kScrStride = 17;                         // float2 units; odd -> conflict-free transposed reads
constexpr int kScrPerFrame = 16 * kScrStride;          // float2 per frame
constexpr int kPStride = kBins + 4;                    // 261, odd: lane <-> frame reads hit 32 banks
constexpr int kMaxUtt = 1024;                          // utterances indexed in shared memory
constexpr int kListCap = 512;                          // tiles per CTA

struct __align__(16) SmemWs {
  float wav[2][kWavSmem];                  // staged samples (gain, pre-emphasis applied), double buffered (I/O -> FFT)
  float2 scr[kFftWarps * 2 * kScrPerFrame];// per frame: 16x17 transpose tile
  float P[2][kTileFrames * kPStride];      // power spectra, double buffered (FFT -> mel)
  float stage[2][kTileFrames * kOutStride];// log-mel outputs, double buffered (mel -> I/O)
  float2 twl[16 * 16];                     // twl[k2][t] = W256^(t*k2)
  float2 hwin2[256];                       // 0.5*Hann zero padded to 512, as (w[2m], w[2m+1])
  float2 tw512[136];                       // W512^k, k = 0..128
  int32_t vcum[kMaxUtt + 1];               // exclusive prefix of valid tiles per utterance
  int32_t pcum[kMaxUtt + 1];               // exclusive prefix of 128-row padding chunks per utterance
  int2 list[kListCap];                     // this CTA's tiles: (utterance, tile index | valid frames << 24)
  int2 info[kListCap];                     //                   (samples in the utterance, gain as float bits)
  unsigned long long bars[16];
};

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
// Forward 16-point DFT, radix 4x4, in registers.  Input natural order; X[4c+d] ends up at v[c+4d] (X16).
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f, H = 0.70710678118654752440f;
#pragma unroll
  for (int a = 0; a < 4; ++a) fft4(v[a], v[a + 4], v[a + 8], v[a + 12]);
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

// ---- mbarrier helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}

template <int W>
__device__ __forceinline__ void mel_two_groups(const float* Prow, const MelFixedW& w, float* srow, float floor_, float scale) {
  mel_fixed_group<2 * W>(Prow, w, srow, floor_, scale);
  mel_fixed_group<2 * W + 1>(Prow, w, srow, floor_, scale);
}

__global__ void __launch_bounds__(kThreadsWs, 1)
logmel_ws_kernel(const __grid_constant__ LogmelArgs a, const __grid_constant__ MelFixedW mw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemWs& S = *reinterpret_cast<SmemWs*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = (int)gridDim.x, me = (int)blockIdx.x;
  const uint32_t bar0 = smem_u32(S.bars);
  // barriers (index b = buffer): samples full/empty, power full/empty, outputs full/empty
  auto bar_wf = [&](int b) { return bar0 + 8u * (uint32_t)(0 + b); };
  auto bar_we = [&](int b) { return bar0 + 8u * (uint32_t)(2 + b); };
  auto bar_pf = [&](int b) { return bar0 + 8u * (uint32_t)(4 + b); };
  auto bar_pe = [&](int b) { return bar0 + 8u * (uint32_t)(6 + b); };
  auto bar_sf = [&](int b) { return bar0 + 8u * (uint32_t)(8 + b); };
  auto bar_se = [&](int b) { return bar0 + 8u * (uint32_t)(10 + b); };

  // ---- prologue (all warps): n_frames, tables, barriers, work lists ------------------------------------
  for (int b = me * kThreadsWs + tid; b < a.B; b += G * kThreadsWs) a.n_frames[b] = frames_of(a.len[b], a);
  for (int i = tid; i < 256; i += kThreadsWs) {
    S.twl[i] = a.tw256[((i & 15) * (i >> 4)) & 255];
    S.hwin2[i] = *reinterpret_cast<const float2*>(a.hwin + 2 * i);
  }
  for (int i = tid; i <= 128; i += kThreadsWs) S.tw512[i] = a.tw512[i];
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_wf(b), kIoWarps);  mbar_init(bar_we(b), kFftWarps);
      mbar_init(bar_pf(b), kFftWarps); mbar_init(bar_pe(b), kMelWarps);
      mbar_init(bar_sf(b), kMelWarps); mbar_init(bar_se(b), kIoWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int u = tid; u < a.B; u += kThreadsWs) {
    const int vt = (frames_of(a.len[u], a) + kTileFrames - 1) / kTileFrames;
    const int pad_rows = a.T_max - vt * kTileFrames;
    S.vcum[u + 1] = vt;
    S.pcum[u + 1] = pad_rows > 0 ? (pad_rows + kPadChunkRows - 1) / kPadChunkRows : 0;
  }
  __syncthreads();
  if (warp == 0) {   // inclusive scans (B <= 1024: 32 steps of a 32-wide scan)
    int cv = 0, cp = 0;
    for (int base = 0; base < a.B; base += 32) {
      const int u = base + lane;
      int v = (u < a.B) ? S.vcum[u + 1] : 0, p = (u < a.B) ? S.pcum[u + 1] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v2 = __shfl_up_sync(0xffffffffu, v, d), p2 = __shfl_up_sync(0xffffffffu, p, d);
        if (lane >= d) { v += v2; p += p2; }
      }
      if (u < a.B) { S.vcum[u + 1] = cv + v; S.pcum[u + 1] = cp + p; }
      cv += __shfl_sync(0xffffffffu, v, 31);
      cp += __shfl_sync(0xffffffffu, p, 31);
    }
    if (lane == 0) { S.vcum[0] = 0; S.pcum[0] = 0; }
  }
  __syncthreads();
  const int vtot = S.vcum[a.B], ptot = S.pcum[a.B];
  const int n_t = (vtot > me) ? (vtot - me - 1) / G + 1 : 0;     // this CTA's tiles (<= kListCap, checked by the host)
  const int n_p = (ptot > me) ? (ptot - me - 1) / G + 1 : 0;     // this CTA's padding chunks
  auto find = [&](const int32_t* cum, int x) -> int {   // largest u in [0,B) with cum[u] <= x
    int lo = 0, hi = a.B;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (cum[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
  };
  for (int k = tid; k < n_t; k += kThreadsWs) {
    const int j = me + k * G;
    const int u = find(S.vcum, j);
    const int tile = j - S.vcum[u];
    const int nv = min(kTileFrames, frames_of(a.len[u], a) - tile * kTileFrames);   // 1..32 valid frames in the tile
    S.list[k] = make_int2(u, tile | (nv << 24));
    float g = 1.0f;
    if (a.normalize) g = __fdiv_rn(1.0f, __fadd_rn(a.peak[u], 1e-9f));  // src/speech_featurizer.py:70
    S.info[k] = make_int2(a.len[u], __float_as_int(g));
  }
  __syncthreads();

  if (warp < kFftWarps) {
    // =========================== FFT warps =======================================================
    // Sixteen warps: warp w transforms frames w (lanes 0-15) and w+16 (lanes 16-31) of every tile, 16 lanes per
    // frame.  The window halves and the transpose twiddles come from shared tables (they would cost 56 registers,
    // and 24 resident warps leave 85 per thread).
    const int t = lane & 15, half = lane >> 4;
    const int fr = warp + 16 * half;                      // this half-warp's frame inside the tile
    float2* scr = S.scr + (warp * 2 + half) * kScrPerFrame;
    const int partner = (lane & 16) | ((16 - t) & 15);
    const float2* twp = S.tw512 + t;                      // W512^(t+16j) at twp[16j]; lane t=0 uses W512^128 for j=0
    const int tw0 = (t == 0) ? 128 : 0;
    for (int k = 0; k < n_t; ++k) {
      const int b = k & 1, n = k >> 1;
      const int nvalid = S.list[k].y >> 24;
      const bool active = (warp < nvalid);                // warp-uniform; frame w+16 may be padding: its samples are
                                                          // stale but finite and its P row is never stored
      mbar_wait(bar_wf(b), n & 1);
      float2 v[16];
      if (active) {
        const float* frp = S.wav[b] + fr * kFrameStep + 2 * t;
#pragma unroll
        for (int m2 = 0; m2 < 13; ++m2) {
          const float2 sv = *reinterpret_cast<const float2*>(frp + 32 * m2);
          const float2 hw = S.hwin2[t + 16 * m2];
          v[m2] = make_float2(sv.x * hw.x, sv.y * hw.y);
        }
        v[13] = v[14] = v[15] = make_float2(0.f, 0.f);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_we(b));              // this warp's samples are in registers
      if (active) {
        fft16(v);
        scr[t] = X16(v, 0);
#pragma unroll
        for (int k2 = 1; k2 < 16; ++k2) scr[k2 * kScrStride + t] = cmul(X16(v, k2), S.twl[k2 * 16 + t]);   // W256^(t*k2)
        __syncwarp();
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) v[n1] = scr[t * kScrStride + n1];
        fft16(v);                                         // X16(v,k1) = Z[t + 16*k1] (half scaled)
      }
      __syncwarp();                                       // (scratch rows read: the next tile may overwrite them)
      if (n > 0) mbar_wait(bar_pe(b), (n - 1) & 1);       // the mel warps have consumed this power buffer
      if (active) {
        float* Pa = S.P[b] + fr * kPStride + t;           // P[k],     k = t + 16j
        float* Pb = S.P[b] + fr * kPStride + 256 - t;     // P[256-k]
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
          const float2 wk = (j == 0) ? twp[tw0] : twp[16 * j]; // W512^k
          const float2 tt = cmul(make_float2(di, -dr), wk);    // W512^k * (-i*D)
          const float ar = er + tt.x, ai = ei + tt.y;          // X[k]
          const float br = er - tt.x, bi = ei - tt.y;          // conj(X[256-k])
          const int ka = (j == 0) ? ((t == 0) ? 128 : 0) : 16 * j;   // lane 0, j = 0: k = 128 (both stores hit P[128])
          Pa[ka] = ar * ar + ai * ai;
          Pb[-ka] = br * br + bi * bi;
        }
        if (t == 0) {
          const float2 z0 = X16(v, 0);
          const float p = 2.0f * (z0.x + z0.y), q = 2.0f * (z0.x - z0.y);
          Pa[0] = p * p;
          Pb[0] = q * q;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pf(b));
    }
  } else if (warp < kFftWarps + kMelWarps) {
    // =========================== mel warps ========================================================
    const int mwp = warp - kFftWarps;
    for (int k = 0; k < n_t; ++k) {
      const int b = k & 1, n = k >> 1;
      mbar_wait(bar_pf(b), n & 1);
      if (n > 0) mbar_wait(bar_se(b), (n - 1) & 1);       // the I/O warps have written this staging tile out
      const float* Prow = S.P[b] + lane * kPStride;       // lane <-> frame
      float* srow = S.stage[b] + lane * kOutStride;
      if (a.mode == 1) {   // "spectrogram": log power of the first 80 FFT bins
        for (int kk = mwp; kk < kMel; kk += kMelWarps) srow[kk] = lg2_normal(fmaxf(Prow[kk], a.floor_)) * a.log_scale;
      } else {
        switch (mwp) {
          case 0: mel_two_groups<0>(Prow, mw, srow, a.floor_, a.log_scale); break;
          case 1: mel_two_groups<1>(Prow, mw, srow, a.floor_, a.log_scale); break;
          case 2: mel_two_groups<2>(Prow, mw, srow, a.floor_, a.log_scale); break;
          default: mel_two_groups<3>(Prow, mw, srow, a.floor_, a.log_scale); break;
        }
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_pe(b)); mbar_arrive(bar_sf(b)); }
    }
  } else {
    // =========================== I/O warps =========================================================
    const int it = tid - (kFftWarps + kMelWarps) * 32;    // 0..127
    int pdone = 0;
    auto pad_upto = [&](int upto) {                       // zero-fill this CTA's padding chunks [pdone, upto)
      for (; pdone < upto; ++pdone) {
        const int j = me + pdone * G;
        const int u = find(S.pcum, j);
        const int vt = S.vcum[u + 1] - S.vcum[u];
        const int r0 = vt * kTileFrames + (j - S.pcum[u]) * kPadChunkRows;
        const int rows = min(kPadChunkRows, a.T_max - r0);
        float* dst = a.out + ((size_t)u * a.T_max + r0) * kMel;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = it; i < rows * (kMel / 4); i += kIoThreads) st_global_v4(dst + 4 * i, z);
      }
    };
    auto prefetch_tile = [&](int k) {                     // tile k's samples -> L2 (5360 samples = 168 lines)
      if (k >= n_t) return;
      const int2 item = S.list[k];
      const float* row = a.wav + (size_t)item.x * a.row_stride;
      for (int l = it; l < 168; l += kIoThreads) {
        const int ns = (item.y & 0xffffff) * kTileFrames * kFrameStep + l * 32;
        if (ns < a.len[item.x]) prefetch_l2(row + ns);
      }
    };
    auto store_tile = [&](int k) {                        // outputs of tile k: staging tile -> global, coalesced
      const int b = k & 1, n = k >> 1;
      const int2 item = S.list[k];
      const int f0 = (item.y & 0xffffff) * kTileFrames;
      const int rows = min(kTileFrames, a.T_max - f0);
      const int nvalid = item.y >> 24;
      float* orow = a.out + ((size_t)item.x * a.T_max + f0) * kMel;
      mbar_wait(bar_sf(b), n & 1);
      const float* stage = S.stage[b];
      for (int i = it; i < rows * (kMel / 4); i += kIoThreads) {
        const int r = i / (kMel / 4), m4 = (i - r * (kMel / 4)) * 4;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);       // rows beyond n_frames[b] inside the tile: the collate's 0.0
        if (r < nvalid) {
          const float* sp = stage + r * kOutStride + m4;
          o = make_float4(sp[0], sp[1], sp[2], sp[3]);
        }
        st_global_v4(orow + 4 * i, o);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_se(b));
    };
    prefetch_tile(0);
    prefetch_tile(1);
    for (int k = 0; k < n_t; ++k) {
      const int b = k & 1, n = k >> 1;
      const int2 item = S.list[k];
      const int nlen = S.info[k].x;
      const int f0 = (item.y & 0xffffff) * kTileFrames;
      const int nvalid = item.y >> 24;
      prefetch_tile(k + 2);
      // ---- stage tile k: gain, pre-emphasis (reference float32 op order) ----------------------------------
      const float* row = a.wav + (size_t)item.x * a.row_stride;
      const int s0 = f0 * kFrameStep;
      const int count = (nvalid - 1) * kFrameStep + kFrameLen;  // multiple of 4; s0+count <= n unless pad_end
      const int lim = a.pad_end ? min(count, nlen - s0) : count;
      const float g = __int_as_float(S.info[k].y);
      const float c = a.preemph;
      if (n > 0) mbar_wait(bar_we(b), (n - 1) & 1);       // the FFT warps have taken tile k-2 out of this buffer
#pragma unroll 1
      for (int u0 = 0; u0 < kIoSlots; u0 += 6) {          // six 128-bit + six 32-bit loads in flight per thread
        float4 x[6];
        float xp[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int idx = 4 * (it + (u0 + q) * kIoThreads);      // sample index inside the tile
          x[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          xp[q] = 0.0f;
          if (idx < lim) {
            x[q] = *reinterpret_cast<const float4*>(row + s0 + idx);   // (rows are padded to 4 samples: in bounds)
            if (s0 + idx > 0) xp[q] = row[s0 + idx - 1];
          }
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int idx = 4 * (it + (u0 + q) * kIoThreads);
          if (idx >= kWavSmem || idx >= count + 16) continue;    // [count, count+16) must be finite zeros (window tail)
          float4 v = x[q];
          v.x = __fmul_rn(v.x, g); v.y = __fmul_rn(v.y, g); v.z = __fmul_rn(v.z, g); v.w = __fmul_rn(v.w, g);  // :71
          float4 y = v;
          if (c > 0.0f) {  // :75-79  y[0]=x[0]; y[n]=x[n]-c*x[n-1], product and difference rounded separately
            const float vp = __fmul_rn(xp[q], g);
            y.x = (s0 + idx > 0) ? __fsub_rn(v.x, __fmul_rn(c, vp)) : v.x;
            y.y = __fsub_rn(v.y, __fmul_rn(c, v.x));
            y.z = __fsub_rn(v.z, __fmul_rn(c, v.y));
            y.w = __fsub_rn(v.w, __fmul_rn(c, v.z));
          }
          if (a.pad_end) {   // the zero padding is appended AFTER pre-emphasis
            if (idx + 0 >= lim) y.x = 0.0f;
            if (idx + 1 >= lim) y.y = 0.0f;
            if (idx + 2 >= lim) y.z = 0.0f;
            if (idx + 3 >= lim) y.w = 0.0f;
          }
          *reinterpret_cast<float4*>(S.wav[b] + idx) = y;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_wf(b));
      // ---- outputs of the tile before, and a share of the padding --------------------------------------
      if (k >= 1) store_tile(k - 1);
      pad_upto((int)(((long long)n_p * (k + 1)) / n_t));
    }
    if (n_t >= 1) store_tile(n_t - 1);
    pad_upto(n_p);
  }
}

}  // namespace

int tasr_logmel_ws_launch(const TasrFeaturizer* f, const LogmelArgs& a, cudaStream_t st) {
  if (!f->mel_fixed || a.B > kMaxUtt) return -1;
  const int grid_full = sm_count();
  const long long total = (long long)a.tiles_per_row * a.B;
  if ((total + grid_full - 1) / grid_full > kListCap) return -1;
  const size_t smem = sizeof(SmemWs);
  TASR_CUDA(cudaFuncSetAttribute(logmel_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long cap = total + a.B;
  const int grid = (int)((cap < grid_full) ? (cap > 0 ? cap : 1) : grid_full);
  const MelFixedW* mw = reinterpret_cast<const MelFixedW*>(f->mel_fixed_w);
  logmel_ws_kernel<<<grid, kThreadsWs, smem, st>>>(a, *mw);
  TASR_LAUNCH_CHECK("logmel_ws_kernel");
  return TASR_OK;
}
