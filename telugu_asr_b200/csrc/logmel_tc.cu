// Fused waveform -> log-mel kernel with the second FFT stage on the tensor cores (sm_100a, tcgen05 + TMEM).
//
// Same contract as logmel_kernel (logmel.cu): src/speech_featurizer.py:136-161 per utterance (normalize_signal ->
// preemphasis_signal -> tf.signal.stft(400/160, periodic Hann, rFFT-512) -> |X|^2 -> HTK mel matmul -> log) plus the
// zero-padded collate of src/dataset.py:236-252.  It is the default whenever the handle has the config/model.yaml
// filterbank structure, feature_type log_mel_spectrogram and pad_end=False; everything else stays on logmel_kernel.
//
// Why: logmel_kernel does the whole 512-point real FFT on the CUDA cores (~490 warp instructions per frame, 16 resident
// warps per SM) and runs at 0.20 of the HBM roof.  Here the transform is split 512 = 16 x 32 (n = n1 + 32 n2,
// k = 16 k1 + p):
//
//   stage 1, CUDA cores   V[n1,p]  = sum_{n2<13} u[n1 + 32 n2] W16^(n2 p),  p = 0..8        thirty-two real-input FFT-16
//                         Vt[n1,p] = V[n1,p] W512^(n1 p)                                     per frame, packed FP32x2
//   stage 2, tcgen05      F_p[k1'] = sum_{n1<32} Vt[n1,p] W32^(n1 k1'),  k1' = 0..31        one GEMM: rows = (p, frame),
//                         K = 64 = (n1, re/im), N = 64 = (k1', re/im), the SAME DFT-32 matrix for every row
//   bins                  X[16 k1' + p] = F_p[k1'] (k1' < 16),  X[512 - 16 k1' - p] = conj F_p[k1'] (k1' >= 16)
//
// Both GEMM operands are split into FP16 high and low parts (hi = a rounded to 11 bits, lo = fp16(a - hi)) and three
// products are issued, Ahi Bhi + Alo Bhi + Ahi Blo, accumulated in FP32 in TMEM: 22 significant bits per operand.  A
// power-of-two scale chosen per 32-frame tile from the tile's max |x| keeps the FP16 parts in the normal range for
// any input level; it is removed exactly after the mel projection.  tools/fft_tc_proto.py is the numpy model of these
// numerics (8e-6 from the float64 oracle on the primary distribution, the same as the float32 oracle's own band).
//
// Work unit = a tile of 28 consecutive frames of one utterance: 9 x 28 = 252 GEMM rows = two M=128 UMMA tiles.
// One persistent CTA per SM, 20 warps with roles (register budget re-dealt with setmaxnreg):
//   warp  0      loader: one cp.async.bulk (TMA bulk copy) per tile brings its raw samples (4 + 4720 floats, contiguous in
//                the utterance) into a 4-deep shared ring, up to four tiles ahead; in single-pass mode the warp also takes
//                max |x| of the few samples behind an utterance's last frame;
//   warps 2-3    scanners: max |x| of a landed tile (-> the tile's power-of-two scale, and the utterance peak in
//                single-pass mode), then hand the tile to the producers;
//   warp  1      MMA issuer: per M-tile and K=16 slice  Ahi x [Bhi | Blo] (N=128)  and  Alo x Bhi (N=64, into the second
//                half), i.e. the main product and the two correction products land in separate TMEM columns and are
//                added in FP32 by the consumers; accumulators double-buffered (2 x 256 of the 512 columns);
//   warps 4-11   producers: thread = (n1 pair, two consecutive frames): gain and pre-emphasis in the reference's float32
//                op order (the sample before a pair comes from the neighbouring lane), tile scale, window,
//                FFT-16, twiddle, hi/lo split, swizzled UMMA A rows; the pair's window and twiddles stay in registers;
//   warps 12-19  consumers: TMEM -> |.|^2 -> P[frame][bin] -> banded mel projection (mel_geometry.inc) -> log ->
//                coalesced stores; they also write the collate padding rows.
// Hand-offs are mbarriers only: raw ring (TMA complete_tx / producer arrivals), the single A stage (producer arrivals /
// tcgen05.commit), accumulators (tcgen05.commit / consumer arrivals).
#include "logmel_common.cuh"
#include "sepconv_common.cuh"
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

using namespace tasr;
using namespace tasr_lm;
using namespace tasr_sep;

namespace {

constexpr int kTcThreads = 640;
constexpr int kWarpLoad = 0, kWarpMmaTc = 1;
constexpr int kProdWarp0 = 4, kProdWarps = 8;
constexpr int kConsWarp0 = 12, kConsWarps = 8;
constexpr int kConsThreads = kConsWarps * 32;

constexpr int kTcFrames = 28;                                            // frames per tile: 9 x 28 = 252 rows
constexpr int kTcRows = 9 * kTcFrames;
// A stage (one tile): UMMA K-major SWIZZLE_128B blocks of 128-byte rows (64 FP16 = 32 n1 x (re, im)); GEMM row
// r = 28 p + f lives in M-tile r / 128:  [M-tile 0: hi 16 KB | lo 16 KB][M-tile 1: hi 16 KB | lo 16 KB].
constexpr int kStageBytes = 65536;
constexpr int kOffLo = 16384, kOffMt = 32768;
constexpr int kPStrideTc = 261;                                          // P[32][261] power rows (28 used; lanes 28..31 of the mel phase
constexpr int kPBytes = 32 * kPStrideTc * 4;                             //   work on the spare rows), [32][81] output tile
constexpr int kOutBytes = 32 * kOutStride * 4;
constexpr int kRawStages = 4;
constexpr int kRawFloats = 4 + 4736;                                     // 4 samples before the tile, 27*160+400 = 4720 of it, 16 zeros
constexpr int kListCapTc = 768;                                          // tiles per CTA per launch
constexpr int kMaxUttTc = 1024;                                          // utterances per launch
constexpr int kAccCols = 256;                                            // 2 M-tiles x (64 main + 64 correction) columns
constexpr int kTmemColsTc = 512;

struct TcLayout {
  uint32_t a, raw, b, p, out, hwin, vcum, pcum, list, lmax, lgain, bars, tmem_slot, total;
};
__host__ __device__ inline TcLayout tc_layout() {
  TcLayout L;
  uint32_t o = 0;
  L.a = o; o += kStageBytes;
  L.b = o; o += 16384;
  L.raw = o; o += kRawStages * kRawFloats * 4;
  L.p = o; o += kPBytes;
  L.out = o; o += kOutBytes;
  L.hwin = o; o += 416 * 4;
  L.vcum = o; o += (kMaxUttTc + 4) * 4;
  L.pcum = o; o += (kMaxUttTc + 4) * 4;
  L.list = o; o += kListCapTc * 8;
  L.lmax = o; o += kListCapTc * 4;
  L.lgain = o; o += kListCapTc * 4;
  L.bars = o; o += 20 * 8;
  L.tmem_slot = o; o += 16;
  L.total = o;
  return L;
}

struct TcArgs {
  LogmelArgs a;
  const unsigned char* dft32;   // 16 KB: shared-memory images of the DFT-32 matrix, FP16 high part then low part
  long long* trace;             // development aid (tools/tc_trace.py): per-role (tag, globaltimer) log of CTA 0, or null
  int32_t ablate;               // development aid (TASR_TC_ABLATE): 1 no MMAs, 2 no stage-1 arithmetic, 4 no consumer arithmetic, 8 no sample loads
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ long long gtime_tc() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// role r logs (tag, time) pairs; 256 pairs per role
#define TC_TRACE(role, tag)                                                                   \
  do {                                                                                        \
    if (ta.trace != nullptr && blockIdx.x == 0 && lane == 0 && trace_n < 255) {               \
      ta.trace[(role) * 512 + 2 * trace_n] = (tag);                                           \
      ta.trace[(role) * 512 + 2 * trace_n + 1] = gtime_tc();                                  \
      ++trace_n;                                                                              \
    }                                                                                         \
  } while (0)
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint32_t cvt_f16x2(float hi_half, float lo_half) {   // {upper 16 bits, lower 16 bits}
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half));
  return r;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Instruction descriptor: D=F32, A=B=F16, both K-major, N, M.
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cons_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsThreads) : "memory"); }

// Power-of-two scale of a tile from its max |x| (bit pattern) and the utterance gain: with mg = max|x| * g in
// [2^E, 2^(E+1)), the samples are multiplied by 2^(10-E) (|x g| < 2048: every stage-1 output stays below 2^15 and FP16's
// normal range covers 29 binades below that).  s2 = 2 * 2^(10-E) because the window table holds 0.5 * Hann; inv = 2^(E-10).
__device__ __forceinline__ void tile_scales(unsigned mbits, float g, float& s2, float& inv) {
  const float mg = __fmul_rn(__uint_as_float(mbits), g);
  int E = (int)((__float_as_uint(mg) >> 23) & 0xffu) - 127;
  E = max(-90, min(100, E));
  s2 = __uint_as_float((unsigned)(127 + 11 - E) << 23);
  inv = __uint_as_float((unsigned)(127 + E - 10) << 23);
}

// hi/lo FP16 split: hi = the value rounded to 11 significant bits (half-ulp added to the bit pattern, then truncated:
// exactly representable in FP16), lo = fp16(a - hi) with |a - hi| <= 2^-11 |a|: 22+ significant bits together.
__device__ __forceinline__ float hi11(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ u64 hi11x2(u64 v) {
  float a, b;
  unpack2(v, a, b);
  return pack2(hi11(a), hi11(b));
}
// split of a packed complex pair (two n1) and its 8-byte stores into the A stage
__device__ __forceinline__ void split_store(u64 re, u64 im, unsigned char* hi_ptr) {
  const u64 re_h = hi11x2(re), im_h = hi11x2(im);
  const u64 re_l = sub2(re, re_h), im_l = sub2(im, im_h);
  float rha, rhb, iha, ihb, rla, rlb, ila, ilb;
  unpack2(re_h, rha, rhb); unpack2(im_h, iha, ihb);
  unpack2(re_l, rla, rlb); unpack2(im_l, ila, ilb);
  *reinterpret_cast<uint2*>(hi_ptr) = make_uint2(cvt_f16x2(iha, rha), cvt_f16x2(ihb, rhb));
  *reinterpret_cast<uint2*>(hi_ptr + kOffLo) = make_uint2(cvt_f16x2(ila, rla), cvt_f16x2(ilb, rlb));
}
__device__ __forceinline__ void split_store_real(u64 re, unsigned char* hi_ptr) {
  const u64 re_h = hi11x2(re);
  const u64 re_l = sub2(re, re_h);
  float rha, rhb, rla, rlb;
  unpack2(re_h, rha, rhb);
  unpack2(re_l, rla, rlb);
  *reinterpret_cast<uint2*>(hi_ptr) = make_uint2(cvt_f16x2(0.0f, rha), cvt_f16x2(0.0f, rhb));
  *reinterpret_cast<uint2*>(hi_ptr + kOffLo) = make_uint2(cvt_f16x2(0.0f, rla), cvt_f16x2(0.0f, rlb));
}
// byte offset of GEMM row r, 16-byte chunk j8, 8-byte half pe inside the stage's hi blocks
__device__ __forceinline__ int a_row_off(int r, int j8, int pe) {
  return (r >> 7) * kOffMt + (r & 127) * 128 + ((j8 ^ (r & 7)) << 4) + pe * 8;
}

// Mel projection of one frame (lane) for the mel bins of group W from its power row: the same walk as tasr_lm::MelSeg
// (every power bin feeds the rising side of bin m and the falling side of bin m-1, ascending k), but the sums stay in
// registers until the whole group is done, so that the loads of the row are not fenced by the result stores.
template <int M, int M0, int M1>
struct MelSegTc {
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
    if constexpr (M < M1) MelSegTc<M + 1, M0, M1>::run(Prow, w, res, acc_cur);
  }
};
// tile scale removed (acc * inv * inv, exact: inv is a power of two) before the floor and the log
template <int W>
__device__ __forceinline__ void mel_group_tc(const float* Prow, const MelFixedW& w, float* srow, float floor_, float scale, float inv) {
  constexpr int M0 = kMelGrp[W], M1 = kMelGrp[W + 1];
  float res[M1 - M0];
  MelSegTc<M0, M0, M1>::run(Prow, w, res, 0.0f);
#pragma unroll
  for (int i = 0; i < M1 - M0; ++i) srow[M0 + i] = lg2_normal(fmaxf((res[i] * inv) * inv, floor_)) * scale;
}

__global__ void __launch_bounds__(kTcThreads, 1)
logmel_tc_kernel(const __grid_constant__ TcArgs ta, const __grid_constant__ MelFixedW mw) {
  const LogmelArgs& a = ta.a;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_u = smem_u32(smem_raw);
  unsigned char* sm = smem_raw + ((1024u - (raw_u & 1023u)) & 1023u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const TcLayout L = tc_layout();

  unsigned char* sA = sm + L.a;
  float* s_raw = reinterpret_cast<float*>(sm + L.raw);
  float* s_hwin = reinterpret_cast<float*>(sm + L.hwin);
  int32_t* vcum = reinterpret_cast<int32_t*>(sm + L.vcum);
  int32_t* pcum = reinterpret_cast<int32_t*>(sm + L.pcum);
  int2* list = reinterpret_cast<int2*>(sm + L.list);
  unsigned* lmax = reinterpret_cast<unsigned*>(sm + L.lmax);
  float* lgain = reinterpret_cast<float*>(sm + L.lgain);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem_slot);
  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sm + L.b), bar_u = smem_u32(sm + L.bars);
  auto bar_wfull = [&](int s) { return bar_u + 8u * (uint32_t)s; };           // raw samples landed (TMA complete_tx)
  auto bar_wempty = [&](int s) { return bar_u + 8u * (uint32_t)(4 + s); };    // raw samples consumed (one per producer warp)
  const uint32_t bar_afull = bar_u + 8u * 8u;                                  // A stage written (one per producer warp)
  const uint32_t bar_afree = bar_u + 8u * 9u;                                  // A stage read by the tensor core (tcgen05.commit)
  auto bar_accf = [&](int s) { return bar_u + 8u * (uint32_t)(10 + s); };     // accumulators complete (tcgen05.commit)
  auto bar_acce = [&](int s) { return bar_u + 8u * (uint32_t)(12 + s); };     // accumulators read (one per consumer warp)
  auto bar_wready = [&](int s) { return bar_u + 8u * (uint32_t)(14 + s); };   // raw samples scanned for max |x| (one per scanner warp)

  // ---- prologue --------------------------------------------------------------------------------------------------
  for (int b = blockIdx.x * kTcThreads + tid; b < a.B; b += gridDim.x * kTcThreads) a.n_frames[b] = frames_of(a.len[b], a);
  if (warp == kWarpMmaTc) tmem_alloc(smem_u32(tmem_slot), kTmemColsTc);
  if (tid == 0) {
    for (int s = 0; s < kRawStages; ++s) {
      mbar_init(bar_wfull(s), 1);
      mbar_init(bar_wready(s), 2);
      mbar_init(bar_wempty(s), kProdWarps);
    }
    mbar_init(bar_afull, kProdWarps);
    mbar_init(bar_afree, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_accf(s), 1);
      mbar_init(bar_acce(s), kConsWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 416; i += kTcThreads) s_hwin[i] = a.hwin[i];
  {
    const uint4* src = reinterpret_cast<const uint4*>(ta.dft32);
    uint4* dst = reinterpret_cast<uint4*>(sm + L.b);
    for (int i = tid; i < 1024; i += kTcThreads) dst[i] = __ldg(src + i);
  }
  for (int i = tid; i < kListCapTc; i += kTcThreads) lmax[i] = 0u;
  for (int u = tid; u < a.B; u += kTcThreads) {
    const int Tu = frames_of(a.len[u], a);
    const int vt = (Tu + kTcFrames - 1) / kTcFrames;
    const int pad_rows = pad_limit(Tu, a) - vt * kTcFrames;
    vcum[u + 1] = vt;
    pcum[u + 1] = pad_rows > 0 ? (pad_rows + kPadChunkRows - 1) / kPadChunkRows : 0;
  }
  __syncthreads();
  if (warp == 0) {   // inclusive scans (B <= 1024)
    int cv = 0, cp = 0;
    for (int base = 0; base < a.B; base += 32) {
      const int u = base + lane;
      int v = (u < a.B) ? vcum[u + 1] : 0, p = (u < a.B) ? pcum[u + 1] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v2 = __shfl_up_sync(0xffffffffu, v, d), p2 = __shfl_up_sync(0xffffffffu, p, d);
        if (lane >= d) { v += v2; p += p2; }
      }
      if (u < a.B) { vcum[u + 1] = cv + v; pcum[u + 1] = cp + p; }
      cv += __shfl_sync(0xffffffffu, v, 31);
      cp += __shfl_sync(0xffffffffu, p, 31);
    }
    if (lane == 0) { vcum[0] = 0; pcum[0] = 0; }
  }
  __syncthreads();
  const int total_v = vcum[a.B], total_p = pcum[a.B];
  const int G = (int)gridDim.x, me = (int)blockIdx.x;
  const int n_v = (total_v > me) ? (total_v - me - 1) / G + 1 : 0;   // <= kListCapTc (checked by the host)
  const int n_p = (total_p > me) ? (total_p - me - 1) / G + 1 : 0;
  const bool single_pass = (a.peak_out != nullptr);
  auto find = [&](const int32_t* cum, int x) -> int {   // largest u in [0,B) with cum[u] <= x
    int lo = 0, hi = a.B;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (cum[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
  };
  for (int k = tid; k < n_v; k += kTcThreads) {
    const int j = me + k * G;
    const int u = find(vcum, j);
    list[k] = make_int2((u << 16) | (j - vcum[u]), a.len[u]);
    // the utterance gain rides in the list (no global load in the tile loops); src/speech_featurizer.py:70
    lgain[k] = (a.normalize && !single_pass) ? __fdiv_rn(1.0f, __fadd_rn(a.peak[u], 1e-9f)) : 1.0f;
  }
  fence_async_smem();     // the DFT-32 images were written through the generic proxy; the tensor core reads them
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  int trace_n = 0;
  if (ta.trace != nullptr && blockIdx.x == 0 && tid == 0) { ta.trace[5 * 512] = n_v; ta.trace[5 * 512 + 1] = gtime_tc(); }

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == kWarpLoad) {
      // =========================== loader (TMA bulk copies) ===================================================
      // raw stage: [0,4) the four samples before the tile (zeros at the start of an utterance: y[0] = x[0] - c*0),
      // [4, 4+count) the tile, then 16 zeros (the last frame's zero-weight window tail must read finite values).
      for (int k = 0; k < n_v; ++k) {
        const int sw = k % kRawStages;
        const int2 item = list[k];
        const int u = item.x >> 16, tf = item.x & 0xffff, n = item.y;
        const int Tb = frames_of(n, a);
        const int f0 = tf * kTcFrames;
        const int nvalid = min(kTcFrames, Tb - f0);
        const int s0 = f0 * kFrameStep;
        const int count = (nvalid - 1) * kFrameStep + kFrameLen;       // multiple of 4; s0 + count <= n
        const float* row = a.wav + (size_t)u * a.row_stride;
        if (k >= kRawStages) mbar_wait(bar_wempty(sw), ((k / kRawStages) - 1) & 1);
        TC_TRACE(0, 1000 + k);
        float* rw = s_raw + sw * kRawFloats;
        if (lane < 4) *reinterpret_cast<float4*>(rw + 4 + count + 4 * lane) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane == 4 && s0 == 0) *reinterpret_cast<float4*>(rw) = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        if (lane == 0) {
          const int pre = (s0 > 0) ? 4 : 0;
          const uint32_t bytes = (ta.ablate & 8) ? 16u : (uint32_t)(count + pre) * 4u;
          mbar_expect_tx(bar_wfull(sw), bytes);
          bulk_g2s(smem_u32(rw + 4 - pre), row + s0 - pre, bytes, bar_wfull(sw));
        }
        if (single_pass && f0 + nvalid >= Tb) {
          // the utterance's last tile: max |x| of the samples no frame covers, [s0+count, n) (fewer than 160)
          unsigned m = 0u;
          for (int i = s0 + count + lane; i < n; i += 32) m = max(m, abs_bits(row[i]));
          m = __reduce_max_sync(0xffffffffu, m);
          if (lane == 0 && m != 0u) atomicMax(reinterpret_cast<unsigned*>(a.peak_out) + u, m);
        }
      }
    } else if (warp >= 2) {
      // =========================== scanners ====================================================================
      const int sc = warp - 2;                       // each scanner takes every other float4 of the tile
      for (int k = 0; k < n_v; ++k) {
        const int sw = k % kRawStages;
        const int2 item = list[k];
        const int u = item.x >> 16, tf = item.x & 0xffff;
        const int nvalid = min(kTcFrames, frames_of(item.y, a) - tf * kTcFrames);
        const int count4 = ((nvalid - 1) * kFrameStep + kFrameLen) >> 2;
        mbar_wait(bar_wfull(sw), (k / kRawStages) & 1);
        const float4* rw4 = reinterpret_cast<const float4*>(s_raw + sw * kRawFloats + 4);
        unsigned m = 0u;
        if (!(ta.ablate & 8))
          for (int i = 2 * lane + sc; i < count4; i += 64) {
            const float4 x = rw4[i];
            m = max(max(m, max(abs_bits(x.x), abs_bits(x.y))), max(abs_bits(x.z), abs_bits(x.w)));
          }
        m = __reduce_max_sync(0xffffffffu, m);
        if (lane == 0) {
          if (m != 0u) {
            atomicMax(&lmax[k], m);
            if (single_pass) atomicMax(reinterpret_cast<unsigned*>(a.peak_out) + u, m);
          }
          mbar_arrive(bar_wready(sw));
        }
      }
    } else if (warp == kWarpMmaTc && lane == 0) {
      // =========================== MMA issuer ==================================================================
      const uint32_t idesc128 = umma_idesc_f16(128, 128), idesc64 = umma_idesc_f16(128, 64);
      const uint64_t db = umma_desc_sw128(sB_u);      // rows 0..63 = Bhi, rows 64..127 = Blo (the two images are contiguous)
      for (int k = 0; k < n_v; ++k) {
        const int acc = k & 1;
        mbar_wait(bar_afull, k & 1);
        if (k >= 2) mbar_wait(bar_acce(acc), ((k >> 1) - 1) & 1);
        tc_fence_after();
        TC_TRACE(1, 1000 + k);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (ta.ablate & 1) break;
          const uint64_t da_hi = umma_desc_sw128(sA_u + mt * kOffMt), da_lo = umma_desc_sw128(sA_u + mt * kOffMt + kOffLo);
          const uint32_t d = tmem + (uint32_t)(acc * kAccCols + mt * 128);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            umma_f16(d, da_hi + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc128, j != 0 ? 1u : 0u);   // [Ahi Bhi | Ahi Blo]
            umma_f16(d + 64u, da_lo + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc64, 1u);             // + Alo Bhi
          }
        }
        umma_commit(bar_afree);
        umma_commit(bar_accf(acc));
        TC_TRACE(1, 2000 + k);
      }
    }
  } else if (warp < kConsWarp0) {
    // =========================== producers =======================================================================
    // thread = (n1 pair c, frame slot): lanes 0-15 / 16-31 of a warp are the sixteen n1 pairs (4 j8 + 2 pe, +1) of two
    // frame slots; a slot is two consecutive frames, whose sample windows overlap by 8 of 13 polyphase values.  The
    // pair's window and twiddles live in registers for the whole kernel: no table loads in the loop.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
    const int ptid = tid - kProdWarp0 * 32;
    const int c16 = lane & 15, j8 = c16 >> 1, pe = c16 & 1;
    const int fa = 2 * ((ptid >> 5) * 2 + (lane >> 4));        // first frame of this thread's slot (0, 2, .., 30; 28 and 30 idle)
    constexpr float kH = 0.70710678118654752440f, kC1 = 0.92387953251128675613f, kS1 = 0.38268343236508977173f;
    const u64 H2 = pack2(kH, kH), NH2 = pack2(-kH, -kH), C12 = pack2(kC1, kC1), S12 = pack2(kS1, kS1), Z2 = pack2(0.f, 0.f);
    const float cpre = a.preemph;
    const u64 CP2 = pack2(cpre, cpre);
    u64 wreg[13], twc[8], tws[8];
#pragma unroll
    for (int n2 = 0; n2 < 13; ++n2) wreg[n2] = *reinterpret_cast<const u64*>(s_hwin + 4 * j8 + 2 * pe + 32 * n2);
#pragma unroll
    for (int p = 1; p <= 8; ++p) {      // W512^(n1 p) = C - i S for the pair's two n1
      const int n1 = 4 * j8 + 2 * pe;
      const float2 w0 = a.tw512[n1 * p], w1 = a.tw512[(n1 + 1) * p];
      twc[p - 1] = pack2(w0.x, w1.x);
      tws[p - 1] = pack2(-w0.y, -w1.y);
    }
    for (int k = 0; k < n_v; ++k) {
      const int sw = k % kRawStages;
      const int2 item = list[k];
      const int tf = item.x & 0xffff;
      const int nvalid = min(kTcFrames, frames_of(item.y, a) - tf * kTcFrames);
      const float g = lgain[k];
      const u64 G2 = pack2(g, g);
      mbar_wait(bar_wready(sw), (k / kRawStages) & 1);
      if (warp == kProdWarp0) TC_TRACE(2, 1000 + k);
      const bool act_a = (fa < nvalid) && !(ta.ablate & 2), act_b = (fa + 1 < nvalid) && !(ta.ablate & 2);
      float s2f, invf;
      tile_scales(lmax[k], g, s2f, invf);
      const u64 S2 = pack2(s2f, s2f);
      bool waited = false;
#pragma unroll 1
      for (int fi = 0; fi < 2; ++fi) {
        const int f = fa + fi;
        const bool active = (fi == 0) ? act_a : act_b;       // (the shuffles below need the whole warp: idle lanes compute along)
        if (!__any_sync(0xffffffffu, active)) break;
        // ---- raw samples of this n1 pair -> gain, pre-emphasis (src/speech_featurizer.py:70-79), tile scale, window -----
        u64 uu[13];
        {
          const float* xp = s_raw + sw * kRawFloats + 4 + f * kFrameStep + 4 * j8 + 2 * pe;
          const float x_before = xp[-1];             // the sample before the pair's first one (used by pair 0 at n2 = 0)
          float v15_prev = 0.0f;                      // pair 15's gained second sample of the previous n2
#pragma unroll
          for (int n2 = 0; n2 < 13; ++n2) {
            const u64 v2 = mul2(*reinterpret_cast<const u64*>(xp + 32 * n2), G2);             // :71  x * gain
            float va, vb;
            unpack2(v2, va, vb);
            // the gained sample before va: the neighbouring pair's second sample; pair 0 takes pair 15's of the previous n2
            float vp = __shfl_up_sync(0xffffffffu, vb, 1, 16);
            const float v15 = __shfl_sync(0xffffffffu, vb, 15, 16);
            if (c16 == 0) vp = (n2 == 0) ? __fmul_rn(x_before, g) : v15_prev;
            v15_prev = v15;
            const u64 y2 = (cpre > 0.0f) ? sub2(v2, mul2(CP2, pack2(vp, va))) : v2;          // :77-79  y[n] = x[n] - c x[n-1]
            uu[n2] = mul2(mul2(y2, S2), wreg[n2]);
          }
        }
        // ---- real-input FFT-16 of (u0..u12, 0, 0, 0): X0, X8 real, X1..X7 complex -------------------------
        const u64 e0 = add2(uu[0], uu[8]), e1 = add2(uu[1], uu[9]), e2 = add2(uu[2], uu[10]), e3 = add2(uu[3], uu[11]),
                  e4 = add2(uu[4], uu[12]), e5 = uu[5], e6 = uu[6], e7 = uu[7];
        const u64 o0 = sub2(uu[0], uu[8]), o1 = sub2(uu[1], uu[9]), o2 = sub2(uu[2], uu[10]), o3 = sub2(uu[3], uu[11]),
                  o4 = sub2(uu[4], uu[12]), o4n = sub2(uu[12], uu[4]), o5 = uu[5], o5n = sub2(Z2, uu[5]), o6 = uu[6], o7 = uu[7];
        u64 xr[9], xi[9];
        {
          const u64 ee0 = add2(e0, e4), ee1 = add2(e1, e5), ee2 = add2(e2, e6), ee3 = add2(e3, e7);
          const u64 eo0 = sub2(e0, e4), eo1 = sub2(e1, e5), eo2 = sub2(e2, e6), eo2n = sub2(e6, e2), eo3 = sub2(e3, e7);
          const u64 aa = add2(ee0, ee2), bb = add2(ee1, ee3);
          xr[0] = add2(aa, bb);
          xr[8] = sub2(aa, bb);
          xr[4] = sub2(ee0, ee2);
          xi[4] = sub2(ee3, ee1);
          const u64 sd = sub2(eo1, eo3), sm_ = add2(eo1, eo3);
          xr[2] = fma2(sd, H2, eo0);
          xr[6] = fma2(sd, NH2, eo0);
          xi[2] = fma2(sm_, NH2, eo2n);
          xi[6] = fma2(sm_, NH2, eo2);
        }
        {
          u64 sd = sub2(o2, o6), sm_ = add2(o2, o6);
          const u64 A0r = fma2(sd, H2, o0), A1r = fma2(sd, NH2, o0), A0i = fma2(sm_, NH2, o4n), A1i = fma2(sm_, NH2, o4);
          sd = sub2(o3, o7); sm_ = add2(o3, o7);
          const u64 B0r = fma2(sd, H2, o1), B1r = fma2(sd, NH2, o1), B0i = fma2(sm_, NH2, o5n), B1i = fma2(sm_, NH2, o5);
          const u64 P1 = fma2(S12, B0i, mul2(C12, B0r)), Q1 = sub2(mul2(C12, B0i), mul2(S12, B0r));
          const u64 P3 = fma2(C12, B1i, mul2(S12, B1r)), Q3 = sub2(mul2(S12, B1i), mul2(C12, B1r));
          xr[1] = add2(A0r, P1); xi[1] = add2(A0i, Q1);
          xr[7] = sub2(A0r, P1); xi[7] = sub2(Q1, A0i);
          xr[3] = add2(A1r, P3); xi[3] = add2(A1i, Q3);
          xr[5] = sub2(A1r, P3); xi[5] = sub2(Q3, A1i);
        }
        // the single A stage is free again once the previous tile's MMAs have read it
        if (fi == 0 && k >= 1) { mbar_wait(bar_afree, (k - 1) & 1); waited = true; }
        // ---- twiddle, split, store: GEMM row 28 p + f ---------------------------------------------------
        if (!active) continue;
        split_store_real(xr[0], sA + a_row_off(f, j8, pe));
#pragma unroll
        for (int p = 1; p <= 8; ++p) {
          const u64 C = twc[p - 1], S = tws[p - 1];
          u64 vr, vi;
          if (p < 8) {
            vr = fma2(xi[p], S, mul2(xr[p], C));
            vi = sub2(mul2(xi[p], C), mul2(xr[p], S));
          } else {
            vr = mul2(xr[8], C);
            vi = sub2(Z2, mul2(xr[8], S));
          }
          split_store(vr, vi, sA + a_row_off(kTcFrames * p + f, j8, pe));
        }
      }
      // A warp with no active frame stores nothing, but it must not run ahead either: mbarrier arrivals are not tagged
      // with a phase, so its arrival for this tile may only be made once the previous tile's phase has completed.
      if (!waited && k >= 1) mbar_wait(bar_afree, (k - 1) & 1);
      fence_async_smem();                  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_afull);
        mbar_arrive(bar_wempty(sw));
      }
      if (warp == kProdWarp0) TC_TRACE(2, 3000 + k);
    }
  } else {
    // =========================== consumers =======================================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int e = warp - kConsWarp0;
    const int q = e & 3, mt = e >> 2;            // TMEM lane quadrant (= warp id % 4), M-tile
    const int grow = mt * 128 + q * 32 + lane;   // this thread's GEMM row = 28 p + f
    const int p = grow / kTcFrames, fr = grow - p * kTcFrames;
    const int ctid = tid - kConsWarp0 * 32;
    float* P = reinterpret_cast<float*>(sm + L.p);
    float* ostg = reinterpret_cast<float*>(sm + L.out);
    int pdone = 0;
    auto do_fill = [&](int upto) {               // collate padding: rows beyond the last valid tile, 128-row chunks
      for (; pdone < upto; ++pdone) {
        const int j = me + pdone * G;
        const int u = find(pcum, j);
        const int vt = vcum[u + 1] - vcum[u];
        const int r0 = vt * kTcFrames + (j - pcum[u]) * kPadChunkRows;
        const int rows = min(kPadChunkRows, pad_limit(frames_of(a.len[u], a), a) - r0);
        float* dst = a.out + ((size_t)u * a.T_max + r0) * kMel;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = ctid; i < rows * (kMel / 4); i += kConsThreads) st_global_v4(dst + 4 * i, z);
      }
    };
    for (int k = 0; k < n_v; ++k) {
      const int acc = k & 1;
      const int2 item = list[k];
      const int u = item.x >> 16, tf = item.x & 0xffff;
      const int Tb = frames_of(item.y, a);
      const int f0 = tf * kTcFrames;
      const int nvalid = min(kTcFrames, Tb - f0);
      const int rows = min(kTcFrames, a.T_max - f0);
      const float g = lgain[k];
      // a NaN peak (a NaN sample somewhere in the utterance) makes every feature of the utterance NaN, as in the reference
      const float floor_b = (g != g) ? g : a.floor_;
      mbar_wait(bar_accf(acc), (k >> 1) & 1);
      tc_fence_after();
      if (e == 0) TC_TRACE(3, 1000 + k);
      float s2f, invf;
      tile_scales(lmax[k], g, s2f, invf);
      if (!(ta.ablate & 4)) {
        // F_p[k1'] = main + correction columns; X[16 k1' + p] (k1' < 16) and conj X[256 - 16 (k1' - 16) - p] (k1' >= 16, 0 < p < 8)
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccCols + mt * 128);
        float* Pw = P + fr * kPStrideTc;
        const bool row_ok = (grow < kTcRows);
        const bool second = row_ok && p >= 1 && p <= 7;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          if (cc >= 2 && !__any_sync(0xffffffffu, second)) break;
          uint32_t rm[16], rc[16];
          tmem_ld16_nowait(taddr + 16u * cc, rm);
          tmem_ld16_nowait(taddr + 64u + 16u * cc, rc);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float re = __uint_as_float(rm[2 * i]) + __uint_as_float(rc[2 * i]);
            const float im = __uint_as_float(rm[2 * i + 1]) + __uint_as_float(rc[2 * i + 1]);
            const float pw = fmaf(re, re, im * im);
            const int k1 = 8 * cc + i;
            if (cc < 2) { if (row_ok) Pw[16 * k1 + p] = pw; }
            else if (second) Pw[256 - 16 * (k1 - 16) - p] = pw;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce(acc));     // this warp has read its part of the accumulators
      cons_bar();
      if (e == 0) TC_TRACE(3, 2000 + k);
      if (!(ta.ablate & 4)) {
        const float* Prow = P + lane * kPStrideTc;     // lane = frame (lanes 28..31 work on the spare rows)
        float* srow = ostg + lane * kOutStride;
        switch (e) {
          case 0: mel_group_tc<0>(Prow, mw, srow, floor_b, a.log_scale, invf); break;
          case 1: mel_group_tc<1>(Prow, mw, srow, floor_b, a.log_scale, invf); break;
          case 2: mel_group_tc<2>(Prow, mw, srow, floor_b, a.log_scale, invf); break;
          case 3: mel_group_tc<3>(Prow, mw, srow, floor_b, a.log_scale, invf); break;
          case 4: mel_group_tc<4>(Prow, mw, srow, floor_b, a.log_scale, invf); break;
          case 5: mel_group_tc<5>(Prow, mw, srow, floor_b, a.log_scale, invf); break;
          case 6: mel_group_tc<6>(Prow, mw, srow, floor_b, a.log_scale, invf); break;
          default: mel_group_tc<7>(Prow, mw, srow, floor_b, a.log_scale, invf); break;
        }
      }
      cons_bar();
      if (e == 0) TC_TRACE(3, 3000 + k);
      if (!(ta.ablate & 4)) {
        // coalesced store; rows beyond n_frames[b] inside this tile are the collate's 0.0
        float* orow = a.out + ((size_t)u * a.T_max + f0) * kMel;
        for (int i = ctid; i < rows * (kMel / 4); i += kConsThreads) {
          const int r = i / (kMel / 4), m4 = (i - r * (kMel / 4)) * 4;
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < nvalid) {
            const float* sp = ostg + r * kOutStride + m4;
            o = make_float4(sp[0], sp[1], sp[2], sp[3]);
          }
          st_global_v4(orow + 4 * i, o);
        }
      }
      if (e == 0) TC_TRACE(3, 4000 + k);
      do_fill((int)(((long long)n_p * (k + 1)) / n_v));
    }
    do_fill(n_p);
    if (e == 0) TC_TRACE(3, 9000);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMmaTc) tmem_dealloc(tmem, kTmemColsTc);
}

// FP16 bits of x rounded to nearest even (host).
uint16_t f16_bits(float x) {
  const __half h = __float2half_rn(x);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
float f16_value(uint16_t b) {
  __half h;
  memcpy(&h, &b, 2);
  return __half2float(h);
}

}  // namespace

// Shared-memory images of the DFT-32 matrix as the UMMA B operand: [N = 64 output columns (k1', re/im)] rows of
// K = 64 FP16 ((n1, re/im), 128 bytes, SWIZZLE_128B), high part then low part.
//   D[row, 2 k1']     = sum_n1  a cos(th) + b sin(th)     (a + i b = Vt[n1], th = 2 pi n1 k1' / 32)
//   D[row, 2 k1' + 1] = sum_n1 -a sin(th) + b cos(th)
void tasr_logmel_tc_build_dft32(unsigned char* img16k) {
  const double two_pi = 6.283185307179586476925286766559;
  double cs[32], sn[32];
  for (int j = 0; j < 32; ++j) { cs[j] = cos(two_pi * j / 32.0); sn[j] = sin(two_pi * j / 32.0); }
  cs[8] = cs[24] = 0.0; sn[0] = sn[16] = 0.0;
  for (int c = 0; c < 64; ++c) {
    const int k1 = c >> 1, imag_out = c & 1;
    for (int kk = 0; kk < 64; ++kk) {
      const int n1 = kk >> 1, imag_in = kk & 1;
      const int j = (n1 * k1) & 31;
      double v;
      if (!imag_out) v = imag_in ? sn[j] : cs[j];
      else v = imag_in ? cs[j] : -sn[j];
      const uint16_t hi = f16_bits((float)v);
      const uint16_t lo = f16_bits((float)(v - (double)f16_value(hi)));
      const size_t off = (size_t)c * 128 + ((((size_t)kk >> 3) ^ ((size_t)c & 7)) << 4) + ((size_t)kk & 7) * 2;
      memcpy(img16k + off, &hi, 2);
      memcpy(img16k + 8192 + off, &lo, 2);
    }
  }
}

static long long* g_tc_trace = nullptr;
// Development aid: a device buffer of 6*512 int64 that CTA 0 of the next launches logs (tag, ns) pairs into.
extern "C" void tasr_debug_tc_trace(long long* dev_buf) { g_tc_trace = dev_buf; }

// Launches logmel_tc_kernel over the batch (sub-batches of at most 1024 utterances / kListCapTc tiles per CTA).
// Returns a negative value (and launches nothing) when the configuration is outside the kernel's scope.
int tasr_logmel_tc_launch(const TasrFeaturizer* f, const tasr_lm::LogmelArgs& a_all, cudaStream_t st) {
  if (!f->mel_fixed || !f->d_dft32 || a_all.mode != 0 || a_all.pad_end) return -1;
  static const int enabled = [] { const char* e = getenv("TASR_LOGMEL_TC"); return e ? atoi(e) : 0; }();
  if (!enabled) return -1;
  const TcLayout L = tc_layout();
  const size_t smem = (size_t)L.total + 1024;
  static bool attr_set[64] = {false};
  if (f->device < 64 && !attr_set[f->device]) {
    TASR_CUDA(cudaFuncSetAttribute(logmel_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[f->device] = true;
  }
  const int sms = sm_count();
  const int tiles_per_row = (a_all.T_max + kTcFrames - 1) / kTcFrames;
  if (tiles_per_row > 0xffff) return -1;
  long long bmax = (long long)kListCapTc * sms / (tiles_per_row > 0 ? tiles_per_row : 1);
  if (bmax < 1) return -1;
  const int bsub = (int)(bmax < kMaxUttTc ? bmax : kMaxUttTc);
  const MelFixedW* mw = reinterpret_cast<const MelFixedW*>(f->mel_fixed_w);
  for (int b0 = 0; b0 < a_all.B; b0 += bsub) {
    TcArgs ta;
    ta.a = a_all;
    ta.dft32 = f->d_dft32;
    static const int ablate = [] { const char* e = getenv("TASR_TC_ABLATE"); return e ? atoi(e) : 0; }();
    ta.ablate = ablate;
    ta.trace = g_tc_trace;
    ta.a.B = (a_all.B - b0 < bsub) ? a_all.B - b0 : bsub;
    ta.a.wav = a_all.wav + (size_t)b0 * a_all.row_stride;
    ta.a.len = a_all.len + b0;
    if (a_all.peak) ta.a.peak = a_all.peak + b0;
    if (a_all.peak_out) ta.a.peak_out = a_all.peak_out + b0;
    ta.a.out = a_all.out + (size_t)b0 * a_all.T_max * kMel;
    ta.a.n_frames = a_all.n_frames + b0;
    const long long cap = (long long)tiles_per_row * ta.a.B + ta.a.B;
    const int grid = (int)(cap < sms ? (cap > 0 ? cap : 1) : sms);
    logmel_tc_kernel<<<grid, kTcThreads, smem, st>>>(ta, *mw);
    TASR_LAUNCH_CHECK("logmel_tc_kernel");
  }
  return TASR_OK;
}
