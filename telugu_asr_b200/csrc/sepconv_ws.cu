// SeparableConv1D forward, persistent warp-specialised TF32 kernel (sm_100a).  The default for every layer whose
// shape is inside its limits (the launcher falls back to the per-tile kernel of sepconv_tf32.cu otherwise);
// TASR_SEPCONV_WS=0 at plan creation selects the per-tile kernel everywhere.
//
// Same arithmetic, operand layouts and summation order as sepconv_tf32.cu (its output is bit-identical,
// tested); what changes is the schedule.  sepconv_tf32_kernel runs one CTA per 128-frame tile and walks
// its channel chunks in lock step (depthwise -> barrier -> MMA issue).  Here ONE CTA per SM lives for the
// whole layer and its warps take roles:
//
//   x loader    (one warp) TMA tiled loads (cp.async.bulk.tensor.3d, tensor map over (C_in, T_in, B)) of each
//               (tile, chunk) input window — 264 rows x 32 channels, out-of-range rows/channels zero-filled by
//               the hardware — into a 3-deep shared ring, two chunks ahead of the arithmetic;
//   depthwise   DWG groups of 8 warps: lane <-> channel, a run of 16 output frames per warp from rows
//               [32w, 32w+39) of the staged window (conflict-free 128-byte rows), result rounded to TF32 into
//               a 2-deep A ring (UMMA K-major SWIZZLE_128B), mbarrier hand-off per stage — no CTA-wide barrier.
//               With DWG = 2 the groups take alternate chunks (group i owns A stage i): the window loads of one
//               chunk run under the FMAs / stores / hand-off of the other;
//   B loader    (one warp) cp.async.bulk (TMA bulk copy) of the packed pw^T chunk into a 2-deep B ring;
//   MMA issuer  (one warp) tcgen05.mma kind::tf32, M=128, N=NT, accumulators double-buffered in TMEM
//               (2 x NT columns of the 512), tcgen05.commit frees ring stages and publishes accumulators;
//   epilogue    EPW (8 or 16) warps: tcgen05.ld 32x32b.x16 (thread <-> output frame) -> +bias -> activation (SFU, all
//               chains of a column block in flight) -> the thread's 64-byte row into a SWIZZLE_64B staging tile ->
//               one TMA tensor store (cp.async.bulk.tensor.3d, 16 channels x 32 frames) per block by lane 0;
//               overlapped with the next tile's main loop;
//   fill        (one warp) streams the constant padding rows of the ragged mode (tiles that lie entirely in the
//               collate padding) with 512-byte coalesced stores.
//
// Work distribution: the tiles that need computing are enumerated through a block-wide prefix sum over the
// utterances (ragged: ceil(cf/256) tiles per utterance, cf = ceil(n_frames / 2^layer)) and dealt round-robin
// to the CTAs, so every SM gets the same number +-1; the padding tiles are dealt the same way.
//
// What paces it (globaltimer traces of CTA 0, tools/ws_trace.py, B200, config 3): with one depthwise group a 32-channel
// chunk takes ~1.05 us — 0.1 us barrier wake-up, 0.35 us window -> registers (the eight warps run in lock step, so their
// 312 shared-memory wavefronts come as one burst), 0.15 us FMAs, 0.16 us TF32 stores, 0.12 us fence + hand-off, 0.2 us
// tap loads — and layers 2 / 3 (6 / 12 chunks per tile) are bound by that chain; layer 1 (3 chunks) is bound by the
// epilogue, ~3.5 us per 128 x 192 tile on eight warps however it stores (transpose + coalesced stores before, TMA now).
#include "sepconv_common.cuh"
#include <cuda.h>
#include <stdlib.h>

using namespace tasr;
using namespace tasr_sep;

namespace {

constexpr int kDwWarps = 8;                     // warps per depthwise group
constexpr int kRunWs = kMT / kDwWarps;          // 16 output frames per depthwise warp
constexpr int kWinWs = 2 * (kRunWs - 1) + 9;    // 39 input rows per run
// depthwise -> MMA ring: STA stages of 16 KB (template parameter: 2, or 3 for the layers bound by the depthwise -> MMA -> commit
// round trip, which then stage their epilogue tiles single-buffered: the 16 KB come from there)
constexpr int kStagesB = 2;   // pw^T chunk ring (NT*128 B each)
constexpr int kStagesX = 3;   // input-tile ring: 264 rows x 32 channels fetched by TMA two chunks ahead of the arithmetic
constexpr int kXRows = 264;   // 2*(128-1)+9 = 263 rows per 128-frame tile, fetched as boxes of 256 + 8 rows
constexpr int kXBytes = kXRows * kKC * 4;
constexpr int kMaxUtt = 512;      // utterances indexed in shared memory
constexpr int kListCap = 256;    // work items per CTA
constexpr int kTmemColsWs = 512;
constexpr int kStgTile = 32 * 64;       // one staging tile: 32 frames x 16 channels, the SWIZZLE_64B image a TMA store reads
__host__ __device__ constexpr int ws_stg_bytes(int sta) { return (sta == 3 ? 8 : 16) * kStgTile; }   // 16 tiles (8 warps x 2 or 16 x 1), or 8 x 1
__host__ __device__ constexpr int ws_threads(int dwg, int epw) { return (dwg * kDwWarps + epw + 4) * 32; }

// TMA tiled load of a 3-D box (coordinates: channel, row, utterance) onto an mbarrier; out-of-range rows and
// channels arrive as zeros.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// TMA tiled store of a 16-channel x 32-frame box from shared memory (frames beyond T_out are clipped by the hardware), committed
// as one bulk async-group of the issuing thread.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

struct WsLayout {
  uint32_t a, b, xs, stg, bias, cum_c, cum_f, utt_cf, utt_gc, list_c, list_f, bars, tmem_slot, total;
};
__host__ __device__ inline WsLayout ws_layout(int NT, int C_out, int sta) {
  WsLayout L;
  uint32_t o = 0;
  L.a = o; o += sta * kABytes;
  L.b = o; o += kStagesB * (uint32_t)NT * 128u;
  L.xs = o; o += kStagesX * kXBytes;
  L.stg = o; o += ws_stg_bytes(sta);         // (a, b and xs are multiples of 1024 bytes: the tiles are 2048-byte aligned)
  L.bias = o; o += (uint32_t)((C_out + 3) & ~3) * 4u;
  L.cum_c = o; o += (kMaxUtt + 1) * 4;
  L.cum_f = o; o += (kMaxUtt + 1) * 4;
  L.utt_cf = o; o += kMaxUtt * 4;
  L.utt_gc = o; o += kMaxUtt * 4;
  L.list_c = o; o += kListCap * 4;
  L.list_f = o; o += kListCap * 4;
  o = (o + 7u) & ~7u;
  L.bars = o; o += 32 * 8;
  L.tmem_slot = o; o += 16;
  L.total = o;
  return L;
}

struct WsArgs {
  CUtensorMap tm_hi;   // x as (C_in, T_in, B) float32, box 32 channels x 64 rows (up to four per window)
  CUtensorMap tm_lo;   // same tensor, box 32 channels x 8 rows (rows 256..263 of a tile's window)
  CUtensorMap tm_y;    // y as (C_out, T_out, B) float32, box 16 channels x 32 rows, SWIZZLE_64B: the epilogue's TMA stores
  SepArgs s;
  int32_t B, n_tiles, n_split;
  long long* trace;   // development aid (tools/ws_trace.py): per-role (tag, globaltimer) log of CTA 0, or null
};

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define WS_TRACE(role, tag)                                                            \
  do {                                                                                 \
    if (wa.trace != nullptr && blockIdx.x == 0 && lane == 0 && trace_n < 250) {        \
      wa.trace[(role) * 512 + 2 * trace_n] = (tag);                                    \
      wa.trace[(role) * 512 + 2 * trace_n + 1] = gtime();                              \
      ++trace_n;                                                                       \
    }                                                                                  \
  } while (0)

template <int CIN, int ACT, int DWG, int EPW, int STA>
__global__ void __launch_bounds__(ws_threads(DWG, EPW), 1) sepconv_ws_kernel(const __grid_constant__ WsArgs wa) {
  constexpr int kStagesA = STA;
  constexpr int kWsThreads = ws_threads(DWG, EPW);
  constexpr int kDwAll = DWG * kDwWarps;
  constexpr bool kRedeal = kWsThreads > 640;       // 896 threads start at 72 registers: re-dealt per role below
  const SepArgs& a = wa.s;
  const int C_in = CIN ? CIN : a.C_in;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NT = a.NT;
  const uint32_t bBytes = (uint32_t)NT * 128u;
  const WsLayout L = ws_layout(NT, a.C_out, STA);
  const long long t_entry = (wa.trace != nullptr) ? gtime() : 0;
  constexpr int kWarpB = kDwAll + EPW, kWarpX = kWarpB + 1, kWarpMma = kWarpB + 2;   // kWarpB + 3: padding fill
  if (warp == kWarpX && lane == 0) {
    prefetch_tmap(&wa.tm_hi);
    prefetch_tmap(&wa.tm_lo);
    prefetch_tmap(&wa.tm_y);
  }

  unsigned char* sA = sm + L.a;
  float* sBias = reinterpret_cast<float*>(sm + L.bias);
  int32_t* cum_c = reinterpret_cast<int32_t*>(sm + L.cum_c);
  int32_t* cum_f = reinterpret_cast<int32_t*>(sm + L.cum_f);
  int32_t* utt_cf = reinterpret_cast<int32_t*>(sm + L.utt_cf);   // cf of each utterance (rows 2t >= cf are padding)
  float* utt_gc = reinterpret_cast<float*>(sm + L.utt_gc);       // deferred input gain of each utterance (2 log g), see SepArgs::in_peak
  int32_t* list_c = reinterpret_cast<int32_t*>(sm + L.list_c);
  int32_t* list_f = reinterpret_cast<int32_t*>(sm + L.list_f);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem_slot);
  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sm + L.b), bar_u = smem_u32(sm + L.bars);
  // barriers: A full (one arrival per warp of the producing group) / A empty (commit), B full (tx) / B empty (commit),
  //           accumulator full (commit) / accumulator empty (one arrival per epilogue warp), x full (tx) / x empty.
  // x full comes in TWO barriers per stage, used by alternate fills of the stage: with two depthwise groups consecutive
  // fills of a stage are consumed by different groups, and TMA loads may complete out of order — a group that waits for
  // fill n of a stage while fill n-1 (the other group's) is still in flight would see the parity of fill n-2 on a single
  // barrier and run through.  On its own barrier a waiter is never more than one phase ahead.
  auto bar_afull = [&](int s) { return bar_u + 8u * (uint32_t)s; };
  auto bar_aempty = [&](int s) { return bar_u + 8u * (uint32_t)(kStagesA + s); };
  auto bar_bfull = [&](int s) { return bar_u + 8u * (uint32_t)(2 * kStagesA + s); };
  auto bar_bempty = [&](int s) { return bar_u + 8u * (uint32_t)(2 * kStagesA + kStagesB + s); };
  auto bar_accf = [&](int i) { return bar_u + 8u * (uint32_t)(2 * kStagesA + 2 * kStagesB + i); };
  auto bar_acce = [&](int i) { return bar_u + 8u * (uint32_t)(2 * kStagesA + 2 * kStagesB + 2 + i); };
  auto bar_xempty = [&](int s) { return bar_u + 8u * (uint32_t)(2 * kStagesA + 2 * kStagesB + 4 + s); };
  auto bar_xfull = [&](int s, int n) { return bar_u + 8u * (uint32_t)(2 * kStagesA + 2 * kStagesB + 4 + kStagesX + 2 * s + (n & 1)); };
  const uint32_t sX_u = smem_u32(sm + L.xs);

  // ---- prologue: TMEM, barriers, bias, work lists ---------------------------------------------------
  if (warp == kWarpMma) tmem_alloc(smem_u32(tmem_slot), kTmemColsWs);
  if (tid == 0) {
    for (int s = 0; s < kStagesA; ++s) {
      mbar_init(bar_afull(s), kDwWarps);
      mbar_init(bar_aempty(s), 1);
    }
    for (int s = 0; s < kStagesB; ++s) {
      mbar_init(bar_bfull(s), 1);
      mbar_init(bar_bempty(s), 1);
    }
    for (int s = 0; s < kStagesX; ++s) {
      mbar_init(bar_xfull(s, 0), 1);
      mbar_init(bar_xfull(s, 1), 1);
      mbar_init(bar_xempty(s), kDwWarps);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_accf(i), 1);
      mbar_init(bar_acce(i), EPW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < a.C_out; i += kWsThreads) sBias[i] = a.bias[i];
  // per utterance: tiles that must be computed (receptive field reaches real data) and tiles that are padding;
  // cf and the deferred gain are kept in shared memory for the roles
  for (int u = tid; u < wa.B; u += kWsThreads) {
    int ct = wa.n_tiles, cf = 0x7fffffff;
    if (a.len0 != nullptr) {
      cf = (max(a.len0[u], 0) + (1 << a.shift) - 1) >> a.shift;
      ct = min(wa.n_tiles, (cf + 2 * kMT - 1) / (2 * kMT));      // tiles t0 with 2*t0 < cf
    }
    int ft = wa.n_tiles - ct;
    if (a.len0 != nullptr && a.fill_rows >= 0) {   // lean: only padding tiles that start within fill_rows of the data
      const int lim = ((cf + 1) >> 1) + a.fill_rows;            // tiles with t0 < lim are written
      ft = max(0, min(wa.n_tiles, (lim + kMT - 1) / kMT) - ct);
    }
    float gc = 0.0f;
    if (a.in_peak != nullptr) {            // g = 1/(peak+1e-9) as the reference rounds it (src/speech_featurizer.py:70)
      float lg;
      const float g = __fdiv_rn(1.0f, __fadd_rn(a.in_peak[u], 1e-9f));
      asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(g));
      gc = a.in_scale2 * lg;
    }
    utt_cf[u] = cf;
    utt_gc[u] = gc;
    cum_c[u + 1] = ct * wa.n_split;
    cum_f[u + 1] = ft;
  }
  __syncthreads();
  if (warp == 0) {   // inclusive scans of the two count arrays (B <= 512: 16 steps of a 32-wide scan)
    int carry_c = 0, carry_f = 0;
    for (int base = 0; base < wa.B; base += 32) {
      const int u = base + lane;
      int vc = (u < wa.B) ? cum_c[u + 1] : 0, vf = (u < wa.B) ? cum_f[u + 1] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int c2 = __shfl_up_sync(0xffffffffu, vc, d), f2 = __shfl_up_sync(0xffffffffu, vf, d);
        if (lane >= d) { vc += c2; vf += f2; }
      }
      if (u < wa.B) { cum_c[u + 1] = carry_c + vc; cum_f[u + 1] = carry_f + vf; }
      carry_c += __shfl_sync(0xffffffffu, vc, 31);
      carry_f += __shfl_sync(0xffffffffu, vf, 31);
    }
    if (lane == 0) { cum_c[0] = 0; cum_f[0] = 0; }
  }
  __syncthreads();
  const int total_c = cum_c[wa.B], total_f = cum_f[wa.B];
  const int G = (int)gridDim.x, me = (int)blockIdx.x;
  const int n_c = (total_c > me) ? (total_c - me - 1) / G + 1 : 0;   // <= kListCap (checked by the host)
  const int n_f = (total_f > me) ? (total_f - me - 1) / G + 1 : 0;
  auto find = [&](const int32_t* cum, int x) -> int {   // largest u in [0,B) with cum[u] <= x
    int lo = 0, hi = wa.B;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (cum[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
  };
  for (int k = tid; k < n_c; k += kWsThreads) {
    const int j = me + k * G;
    const int u = find(cum_c, j);
    const int r = j - cum_c[u];
    const int tile = r / wa.n_split, nh = r - tile * wa.n_split;
    list_c[k] = (u << 16) | (tile << 4) | nh;
  }
  for (int k = tid; k < n_f; k += kWsThreads) {
    const int j = me + k * G;
    const int u = find(cum_f, j);
    const int ct = (cum_c[u + 1] - cum_c[u]) / wa.n_split;
    list_f[k] = (u << 16) | (ct + (j - cum_f[u]));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (wa.trace != nullptr && blockIdx.x == 0 && tid == 0) { wa.trace[5 * 512] = 1; wa.trace[5 * 512 + 1] = gtime(); wa.trace[5 * 512 + 2] = n_c; wa.trace[5 * 512 + 3] = n_f; wa.trace[5 * 512 + 4] = t_entry; }
  const uint32_t tmem = *tmem_slot;
  const int n_chunks = a.n_chunks;
  const int n_g = n_c * n_chunks;          // (tile, chunk) steps of this CTA
  int trace_n = 0;

  if (warp < kDwAll) {
    // =========================== depthwise producers ===========================================
    // The input window of every (tile, chunk) — 264 rows x 32 channels — is brought into the shared x ring by
    // the TMA warp two chunks ahead; each warp reduces its run of 16 output frames from rows [32w, 32w+39)
    // of it (lane <-> channel, conflict-free rows of 128 B) into the A ring.  Group `grp` takes steps grp, grp + DWG, ...
    if (kRedeal) {
      if (DWG == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
      else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    }
    const int grp = warp / kDwWarps, wg = warp - grp * kDwWarps;
    const float* xs_f = reinterpret_cast<const float*>(sm + L.xs);
    auto load_w = [&](float (&w)[9], int kc) {     // depthwise taps of chunk kc for this lane's channel
      const int c0 = kc * kKC;
      const bool cok = lane < min(kKC, C_in - c0);
#pragma unroll
      for (int kk = 0; kk < 9; ++kk) w[kk] = cok ? __ldg(a.dw + kk * C_in + c0 + lane) : 0.0f;
    };
    float w[9];
    if (grp < n_g) load_w(w, grp % n_chunks);
    for (int g = grp; g < n_g; g += DWG) {
      const int k = g / n_chunks, kc = g - k * n_chunks;
      const int item = list_c[k];
      const int r0 = 2 * (((item >> 4) & 0xfff) * kMT + wg * kRunWs);   // first input row of this warp's run
      const int cf = utt_cf[item >> 16];
      const bool run_needed = (r0 < cf);   // a run that starts in the padding only yields the constant row
      const int sx = g % kStagesX, nx = g / kStagesX;
      const int s = g % kStagesA, n = g / kStagesA;
      if (warp == 0) WS_TRACE(0, 140 + kc);
      mbar_wait(bar_xfull(sx, nx), (nx >> 1) & 1);
      if (warp == 0) WS_TRACE(0, 160 + kc);
      float v[kWinWs];
      if (run_needed) {
        const float* xw = xs_f + (size_t)sx * (kXBytes / 4) + (size_t)(2 * wg * kRunWs) * kKC + lane;
#pragma unroll
        for (int i = 0; i < kWinWs; ++i) v[i] = xw[i * kKC];
        if (a.in_peak != nullptr) {      // rows of the data get the deferred gain and the floor; padding rows stay 0.0
          const float gc = utt_gc[item >> 16];
          const float fl = (gc == gc) ? a.in_floor : gc;   // NaN peak -> NaN rows (fmaxf alone would drop the NaN)
          if (r0 + kWinWs <= cf) {
#pragma unroll
            for (int i = 0; i < kWinWs; ++i) v[i] = fmaxf(v[i] + gc, fl);
          } else {
#pragma unroll
            for (int i = 0; i < kWinWs; ++i) v[i] = (r0 + i < cf) ? fmaxf(v[i] + gc, fl) : v[i];
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_xempty(sx));            // this warp has its window in registers
      if (n > 0) mbar_wait(bar_aempty(s), (n - 1) & 1);
      if (warp == 0) WS_TRACE(0, 100 + kc);
      if (run_needed) {
        unsigned char* As = sA + s * kABytes;
        float acc[kRunWs];
        depthwise_run16(v, w, acc);
#pragma unroll
        for (int j = 0; j < kRunWs; ++j) {
          const int row = wg * kRunWs + j;
          const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)lane >> 2) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)lane & 3u) * 4u;
          *reinterpret_cast<uint32_t*>(As + off) = to_tf32(acc[j]);
        }
      }
      fence_async_smem();                  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_afull(s));
      if (warp == 0) WS_TRACE(0, 200 + kc);
      // taps of this group's next step: the loads fly during its x wait and window loads (the window registers are dead here)
      if (g + DWG < n_g) load_w(w, (g + DWG) % n_chunks);
    }
  } else if (warp < kDwAll + EPW) {
    // =========================== epilogue ========================================================
    // Per 16-column block: tcgen05.ld (thread <-> output frame) -> + bias -> activation with every chain of the block in
    // flight (act_block16) -> the thread's 64-byte row into a staging tile in the SWIZZLE_64B pattern (conflict-free 128-bit
    // stores) -> one TMA tensor store of the 16 x 32 box by lane 0.  Frames beyond T_out are clipped by the tensor map,
    // frames in the collate padding take the layer's constant row.
    if (kRedeal && EPW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    constexpr int kTiles = ws_stg_bytes(STA) / kStgTile / EPW;   // staging tiles per warp: 2 (double-buffered) or 1
    static_assert(kTiles >= 1, "no staging tile for this role / stage combination");
    constexpr int kSlices = EPW / 4;          // warps per TMEM lane quadrant
    const int e = warp - kDwAll;
    const int q = e & 3, slice = e >> 2;      // TMEM lane quadrant (= warp id % 4), column-block phase
    unsigned char* stg = sm + L.stg + e * (kTiles * kStgTile);
    const uint32_t stg_u = smem_u32(stg);
    const int nblocks = NT >> 4;
    int issued = 0;                           // TMA stores issued by lane 0 of this warp
    for (int k = 0; k < n_c; ++k) {
      const int item = list_c[k];
      const int b = item >> 16, t0 = ((item >> 4) & 0xfff) * kMT, n0 = (item & 0xf) * NT;
      const int cf = utt_cf[b];
      const int acc = k & 1;
      const bool is_pad = 2 * (t0 + q * 32 + lane) >= cf;    // this thread's output frame lies in the collate padding
      if (e == 0 || e == EPW - 1) WS_TRACE(e == 0 ? 3 : 4, 300);
      mbar_wait(bar_accf(acc), (k >> 1) & 1);
      tc_fence_after();
      if (e == 0 || e == EPW - 1) WS_TRACE(e == 0 ? 3 : 4, 301);
      for (int g0 = slice; g0 < nblocks; g0 += kSlices * kTiles) {
        // kTiles column blocks per round (one per staging tile): one proxy fence and one hand-off to lane 0 for all of them
        if (lane == 0 && issued > 0) bulk_wait_read<0>();   // the previous round's stores have read the staging tiles
        __syncwarp();
#pragma unroll
        for (int j = 0; j < kTiles; ++j) {
          const int g = g0 + j * kSlices;
          if (g < nblocks) {
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT + g * 16), r);
            u64 z[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 bv = *reinterpret_cast<const float4*>(sBias + n0 + g * 16 + 4 * i);
              z[2 * i] = add2(pack2(__uint_as_float(r[4 * i + 0]), __uint_as_float(r[4 * i + 1])), pack2(bv.x, bv.y));
              z[2 * i + 1] = add2(pack2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), pack2(bv.z, bv.w));
            }
            act_block16<ACT>(z);
            unsigned char* tile = stg + j * kStgTile;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 o;
              unpack2(z[2 * i], o.x, o.y);
              unpack2(z[2 * i + 1], o.z, o.w);
              if (is_pad) o = __ldg(reinterpret_cast<const float4*>(a.pad_out + n0 + g * 16 + 4 * i));
              *reinterpret_cast<float4*>(tile + lane * 64 + (((uint32_t)i ^ (((uint32_t)lane >> 1) & 3u)) << 4)) = o;
            }
          }
        }
        fence_async_smem();                  // generic-proxy writes -> visible to the TMA engine (async proxy)
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < kTiles; ++j) {
            const int g = g0 + j * kSlices;
            if (g < nblocks) tma_store_3d(&wa.tm_y, n0 + g * 16, t0 + q * 32, b, stg_u + j * kStgTile);
          }
        }
        ++issued;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce(acc));   // this warp has read its part of the accumulator
      if (e == 0 || e == EPW - 1) WS_TRACE(e == 0 ? 3 : 4, 302);
    }
    if (lane == 0 && issued > 0) bulk_wait_all();   // the stores have completed before the CTA exits
    if (e == 0 || e == EPW - 1) WS_TRACE(e == 0 ? 3 : 4, 304);
  } else {
    if (kRedeal) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp == kWarpB) {
      // =========================== B loader ======================================================
      if (lane == 0) {
        int g = 0;
        for (int k = 0; k < n_c; ++k) {
          const int nh = list_c[k] & 0xf;
          const float* bsrc = a.bpack + (size_t)nh * n_chunks * NT * kKC;
          for (int kc = 0; kc < n_chunks; ++kc, ++g) {
            const int s = g % kStagesB, n = g / kStagesB;
            if (n > 0) mbar_wait(bar_bempty(s), (n - 1) & 1);
            mbar_expect_tx(bar_bfull(s), bBytes);
            bulk_g2s(sB_u + s * bBytes, bsrc + (size_t)kc * NT * kKC, bBytes, bar_bfull(s));
          }
        }
      }
    } else if (warp == kWarpX) {
      // =========================== x loader (TMA) ================================================
      if (lane == 0) {
        int g = 0;
        for (int k = 0; k < n_c; ++k) {
          const int item = list_c[k];
          const int b = item >> 16, row0 = 2 * ((item >> 4) & 0xfff) * kMT + a.row_off;   // (negative / past-the-end rows arrive as zeros)
          const int cf = utt_cf[b];
          const int need = (cf >= 0x7fffff00) ? kXRows : min(kXRows, cf + 8 - row0);   // >= 9: the tile has a needed output
          const int n64 = min(4, (need + 63) >> 6);
          const bool tail = need > 256;
          for (int kc = 0; kc < n_chunks; ++kc, ++g) {
            const int s = g % kStagesX, n = g / kStagesX;
            if (n > 0) mbar_wait(bar_xempty(s), (n - 1) & 1);
            WS_TRACE(1, 800 + kc);
            // Only the rows a needed output can read are fetched: [row0, cf + 8) in 64-row boxes (+ the 8-row tail box).  The
            // rest of the stage keeps stale values that only reach frames in the collate padding (replaced by the epilogue).
            mbar_expect_tx(bar_xfull(s, n), (uint32_t)(n64 * 64 + (tail ? 8 : 0)) * kKC * 4u);
            const uint32_t dst = sX_u + (uint32_t)s * kXBytes;
            for (int i = 0; i < n64; ++i) tma_load_3d(dst + (uint32_t)i * 64u * kKC * 4u, &wa.tm_hi, kc * kKC, row0 + 64 * i, b, bar_xfull(s, n));
            if (tail) tma_load_3d(dst + 256u * kKC * 4u, &wa.tm_lo, kc * kKC, row0 + 256, b, bar_xfull(s, n));
          }
        }
      }
    } else if (warp == kWarpMma) {
      // =========================== MMA issuer ====================================================
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_tf32(kMT, NT);
        int g = 0;
        for (int k = 0; k < n_c; ++k) {
          const int acc = k & 1;
          if (k >= 2) {                        // the epilogue has drained this accumulator (tile k-2)
            mbar_wait(bar_acce(acc), ((k >> 1) - 1) & 1);
            tc_fence_after();
          }
          const uint32_t d_tmem = tmem + (uint32_t)(acc * NT);
          for (int kc = 0; kc < n_chunks; ++kc, ++g) {
            const int sa = g % kStagesA, na = g / kStagesA, sb = g % kStagesB, nb = g / kStagesB;
            mbar_wait(bar_afull(sa), na & 1);
            WS_TRACE(2, 400 + kc);
            mbar_wait(bar_bfull(sb), nb & 1);
            tc_fence_after();
            WS_TRACE(2, 500 + kc);
            const uint64_t da = umma_desc_sw128(sA_u + sa * kABytes);
            const uint64_t db = umma_desc_sw128(sB_u + sb * bBytes);
            const int ksteps = min(kKC, C_in - kc * kKC) >> 3;
            for (int kk = 0; kk < ksteps; ++kk)
              umma_tf32(d_tmem, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc, (kc | kk) != 0 ? 1u : 0u);
            umma_commit(bar_aempty(sa));
            umma_commit(bar_bempty(sb));
          }
          umma_commit(bar_accf(acc));
          WS_TRACE(2, 600);
        }
      }
    } else {
      // =========================== padding fill ==================================================
      // Tiles that lie entirely in the collate padding are whole rows of the layer's constant padding row: one warp streams
      // them with 512-byte coalesced stores (lane <-> float4 slot of the flattened tile; when C_out / 4 divides 96 the slot's
      // column repeats every three steps, so the lane keeps its three float4 of the row in registers) while the others compute.
      const int q4 = a.C_out >> 2;
      const float4* pr = reinterpret_cast<const float4*>(a.pad_out);
      if (n_f > 0 && 96 % q4 == 0) {
        float4 pv[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) pv[i] = __ldg(pr + (lane + 32 * i) % q4);
        for (int k = 0; k < n_f; ++k) {
          const int item = list_f[k];
          const int b = item >> 16, t0 = (item & 0xffff) * kMT;
          const int n4 = min(kMT, a.T_out - t0) * q4;
          float4* dst = reinterpret_cast<float4*>(a.y + ((size_t)b * a.T_out + t0) * a.C_out);
          int i = lane;
          for (; i + 64 < n4; i += 96) {
            dst[i] = pv[0];
            dst[i + 32] = pv[1];
            dst[i + 64] = pv[2];
          }
          if (i < n4) dst[i] = pv[0];
          if (i + 32 < n4) dst[i + 32] = pv[1];
        }
      } else {
        for (int k = 0; k < n_f; ++k) {
          const int item = list_f[k];
          const int b = item >> 16, t0 = (item & 0xffff) * kMT;
          const int rows = min(kMT, a.T_out - t0);
          float* dst = a.y + ((size_t)b * a.T_out + t0) * a.C_out;
          for (int c4 = lane; c4 < q4; c4 += 32) {
            const float4 pv = __ldg(pr + c4);
            for (int r = 0; r < rows; ++r) *reinterpret_cast<float4*>(dst + (size_t)r * a.C_out + 4 * c4) = pv;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wa.trace != nullptr && blockIdx.x == 0 && tid == 0) wa.trace[5 * 512 + 5] = gtime();
  if (warp == kWarpMma) tmem_dealloc(tmem, kTmemColsWs);
}

typedef void (*WsKernel)(const WsArgs);
template <int CIN, int DWG, int EPW, int STA>
WsKernel pick_act_ws(int act) {
  switch (act) {
    case TASR_ACT_TANH: return sepconv_ws_kernel<CIN, TASR_ACT_TANH, DWG, EPW, STA>;
    case TASR_ACT_GELU_ERF: return sepconv_ws_kernel<CIN, TASR_ACT_GELU_ERF, DWG, EPW, STA>;
    default: break;
  }
  if (DWG != 1 || EPW != 8) return nullptr;      // the other activations / widths only exist with the base role counts
  switch (act) {
    case TASR_ACT_RELU: return sepconv_ws_kernel<CIN, TASR_ACT_RELU, 1, 8, 2>;
    default: return sepconv_ws_kernel<CIN, TASR_ACT_NONE, 1, 8, 2>;
  }
}
template <int DWG, int EPW, int STA>
WsKernel pick_kernel_ws(int c_in, int act) {
  switch (c_in) {
    case 80: return pick_act_ws<80, DWG, EPW, STA>(act);
    case 192: return pick_act_ws<192, DWG, EPW, STA>(act);
    case 384: return pick_act_ws<384, DWG, EPW, STA>(act);
    default: return (DWG == 1 && EPW == 8) ? pick_act_ws<0, 1, 8, 2>(act) : nullptr;
  }
}

}  // namespace

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static long long* g_ws_trace = nullptr;
// Development aid: a device buffer of 6*512 int64 that CTA 0 of the next launches logs (tag, ns) pairs into.
extern "C" void tasr_debug_ws_trace(long long* dev_buf) { g_ws_trace = dev_buf; }

// Returns TASR_OK after launching, or a negative value when this shape is not handled by the persistent
// kernel (the caller then uses sepconv_tf32_kernel).
int tasr_sepconv_ws_launch(const TasrSepConvPlan* p, const SepArgs& sa_all, int32_t B_all, cudaStream_t st) {
  if (B_all > kMaxUtt) {
    // The kernel indexes at most kMaxUtt utterances in shared memory: a larger batch runs as consecutive sub-batches
    // (utterances are independent; each sub-launch is a full persistent grid) — when a sub-batch is worth a persistent
    // launch (measured: 30 s x 1024 8.4 -> 9.1 M audio-s/s, but 1 s x 1024 3.1 -> 2.8 M: short tensors stay per-tile).
    const long long items = (long long)kMaxUtt * ((sa_all.T_out + kMT - 1) / kMT) * p->n_split;
    if (items < 8LL * sm_count()) return -1;
    for (int32_t b0 = 0; b0 < B_all; b0 += kMaxUtt) {
      SepArgs sub = sa_all;
      sub.x = sa_all.x + (size_t)b0 * sa_all.T_in * p->L.c_in;
      sub.y = sa_all.y + (size_t)b0 * sa_all.T_out * p->L.c_out;
      if (sa_all.len0) sub.len0 = sa_all.len0 + b0;
      if (sa_all.in_peak) sub.in_peak = sa_all.in_peak + b0;
      const int rc = tasr_sepconv_ws_launch(p, sub, B_all - b0 < kMaxUtt ? B_all - b0 : kMaxUtt, st);
      if (rc != TASR_OK) return (b0 == 0) ? rc : (rc < 0 ? fail(TASR_ERR_CUDA, "sepconv_ws: sub-batch launch refused after earlier sub-batches ran") : rc);
    }
    return TASR_OK;
  }
  const SepArgs& sa = sa_all;
  const int32_t B = B_all;
  const int n_tiles = (sa.T_out + kMT - 1) / kMT;
  if (n_tiles > 0xfff || p->n_split > 15 || 2 * p->NT > kTmemColsWs || (p->NT & 15) || (p->L.c_out & 15)) return -1;
  const int grid = sm_count();
  const long long dense = (long long)B * n_tiles * p->n_split;
  if ((dense + grid - 1) / grid > kListCap) return -1;
  // Role counts: two depthwise groups on alternate chunks wherever that instantiation exists (layers 2 / 3 are bound by the
  // depthwise chain; layer 1 is not slower with it), with a third A stage for the layers of more than three chunks per tile
  // (layer 1 is bound by its epilogue and keeps the double-buffered staging tiles instead); the plan's
  // TASR_WS_ROLES = 18 | 28 | 116 overrides (development aid).
  int roles = p->ws_roles;
  const int sta = (roles == 28 && p->n_chunks > 3) ? 3 : 2;
  WsKernel kern = roles == 116 ? pick_kernel_ws<1, 16, 2>(p->L.c_in, p->L.activation)
                : roles == 28 ? (sta == 3 ? pick_kernel_ws<2, 8, 3>(p->L.c_in, p->L.activation) : pick_kernel_ws<2, 8, 2>(p->L.c_in, p->L.activation))
                              : nullptr;
  if (kern == nullptr) { roles = 18; kern = pick_kernel_ws<1, 8, 2>(p->L.c_in, p->L.activation); }
  const int threads = roles == 18 ? ws_threads(1, 8) : ws_threads(2, 8);
  const WsLayout L = ws_layout(p->NT, p->L.c_out, roles == 28 ? sta : 2);
  const size_t smem = (size_t)L.total + 1024;
  if (smem > 227 * 1024) return -1;
  TASR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  WsArgs wa;
  {
    static PFN_encodeTiled encode = nullptr;
    if (!encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      TASR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      if (!fn || qres != cudaDriverEntryPointSuccess) return -1;
      encode = reinterpret_cast<PFN_encodeTiled>(fn);
    }
    const cuuint64_t dims[3] = {(cuuint64_t)p->L.c_in, (cuuint64_t)sa.T_in, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)p->L.c_in * 4, (cuuint64_t)sa.T_in * p->L.c_in * 4};
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint32_t box_hi[3] = {(cuuint32_t)kKC, 64, 1}, box_lo[3] = {(cuuint32_t)kKC, 8, 1};
    CUresult r1 = encode(&wa.tm_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(sa.x), dims, strides, box_hi, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&wa.tm_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(sa.x), dims, strides, box_lo, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return -1;   // e.g. strides not multiples of 16 bytes: per-tile kernel instead
    const cuuint64_t ydims[3] = {(cuuint64_t)p->L.c_out, (cuuint64_t)sa.T_out, (cuuint64_t)B};
    const cuuint64_t ystrides[2] = {(cuuint64_t)p->L.c_out * 4, (cuuint64_t)sa.T_out * p->L.c_out * 4};
    const cuuint32_t box_y[3] = {16, 32, 1};
    CUresult r3 = encode(&wa.tm_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, sa.y, ydims, ystrides, box_y, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r3 != CUDA_SUCCESS) return -1;
  }
  wa.s = sa; wa.B = B; wa.n_tiles = n_tiles; wa.n_split = p->n_split;
  wa.trace = g_ws_trace;
  const int g = (int)(dense < grid ? (dense > 0 ? dense : 1) : grid);
  kern<<<g, threads, smem, st>>>(wa);
  TASR_LAUNCH_CHECK("sepconv_ws_kernel");
  return TASR_OK;
}
