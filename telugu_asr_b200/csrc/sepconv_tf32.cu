// SeparableConv1D forward on the 5th-generation tensor cores (TASR_MATH_TF32), sm_100a only.
//
// Replaces one tf.keras.layers.SeparableConv1D call (src/models/moonshine/encoder.py:31-40,60):
// depthwise cross-correlation (k=9, stride 2, VALID, no bias) -> 1x1 pointwise -> bias_add ->
// activation.  One kernel per layer:
//
//   tile      128 output frames of one utterance x NT output channels (NT = c_out / n_split <= 256);
//   depthwise CUDA cores, straight from global memory: lane <-> channel (128-byte coalesced rows),
//             each warp a run of 16 output frames with a 39-row register sliding window; the result
//             is rounded to TF32 (cvt.rna) and written to shared memory as the A operand in the
//             UMMA K-major SWIZZLE_128B layout (32 channels = one 128-byte row per frame);
//   pointwise tcgen05.mma.cta_group::1.kind::tf32, M=128, N=NT, K=8 per instruction, FP32
//             accumulators in tensor memory (TMEM).  The B operand (pw^T, TF32-rounded) is packed
//             once per layer by the plan as ready-made shared-memory images and fetched per
//             32-channel chunk with cp.async.bulk (TMA bulk copy) onto an mbarrier;
//   pipeline  two A/B stages: the MMAs of chunk i run asynchronously while all warps compute the
//             depthwise result of chunk i+1; tcgen05.commit frees a stage; two CTAs per SM overlap
//             one CTA's epilogue with the other's main loop;
//   epilogue  tcgen05.ld (32 lanes x 32 columns per warp) -> + bias -> activation -> per-warp
//             shared-memory transpose -> 128-bit coalesced stores, rows t >= t_out masked.
//
// Like the reference, the convolution runs over the whole zero-padded tensor (no masking between
// layers).  Error model: both GEMM operands carry TF32 rounding (2^-11 relative), accumulation is
// FP32 — measured <= 3e-4 of max|y| per layer against the float64 oracle (budget 1e-3).
#include "sepconv_common.cuh"
#include <stdlib.h>

using namespace tasr;

using namespace tasr_sep;

namespace {

// CIN > 0 bakes the row stride into load immediates (the reference shapes 80/192/384 and the
// model_dim=288 default 288/576); CIN == 0 reads it from the arguments.
template <int CIN, int ACT>
__global__ void __launch_bounds__(kThreads, 2) sepconv_tf32_kernel(const SepArgs a) {
  const int C_in = CIN ? CIN : a.C_in;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NT = a.NT;
  const uint32_t bBytes = (uint32_t)NT * 128u;

  unsigned char* sA = sm;                          // 2 x 16 KiB
  unsigned char* sB = sm + 2 * kABytes;            // 2 x NT*128
  float* sBias = reinterpret_cast<float*>(sB + 2 * bBytes);   // 256 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 256);  // [0,1] B full, [2,3] stage free, [4] accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB), bar_u = smem_u32(bars);

  const int b = blockIdx.z, nh = blockIdx.y, t0 = blockIdx.x * kMT, n0 = nh * NT;

  // Rows t with 2t >= cf have a receptive field made of the input's padding row only.
  int cf = 0x7fffffff;
  if (a.len0 != nullptr) {               // CTA-uniform: leaves before any barrier / TMEM allocation
    cf = (max(a.len0[b], 0) + (1 << a.shift) - 1) >> a.shift;
    if (2 * t0 >= cf) {
      if (a.fill_rows >= 0 && t0 >= ((cf + 1) >> 1) + a.fill_rows) return;   // lean: nobody reads this tile
      const int rows = min(kMT, a.T_out - t0), q4 = NT >> 2;   // q4 <= 64 float4 per row: one warp per row
      float* dst = a.y + ((size_t)b * a.T_out + t0) * a.C_out + n0;
      const float4* pr = reinterpret_cast<const float4*>(a.pad_out + n0);
      const float4 p0 = (lane < q4) ? __ldg(pr + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 p1 = (lane + 32 < q4) ? __ldg(pr + lane + 32) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = warp; r < rows; r += kThreads / 32) {
        float4* row = reinterpret_cast<float4*>(dst + (size_t)r * a.C_out);
        if (lane < q4) row[lane] = p0;
        if (lane + 32 < q4) row[lane + 32] = p1;
      }
      return;
    }
  }

  // deferred input gain of the single-pass featurizer (CTA-uniform): c = 2*log(g), g = 1/(peak+1e-9) as the reference
  // rounds it (src/speech_featurizer.py:70)
  const bool fix = (a.in_peak != nullptr);
  float gc = 0.0f;
  if (fix) {
    float lg;
    const float g = __fdiv_rn(1.0f, __fadd_rn(a.in_peak[b], 1e-9f));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(g));
    gc = a.in_scale2 * lg;
  }

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  if (tid == 32) {
    for (int i = 0; i < 5; ++i) mbar_init(bar_u + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < NT; i += kThreads) sBias[i] = a.bias[n0 + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const float* bsrc = a.bpack + (size_t)nh * a.n_chunks * NT * kKC;
  const uint32_t idesc = umma_idesc_tf32(kMT, NT);
  const int tw = t0 + warp * kRun;       // first output frame of this warp's run
  const int r0 = 2 * tw + a.row_off;     // first input row of the run (row_off = -pad_left for padding='same', else 0)
  const float* xrow = a.x + ((long long)b * a.T_in + r0) * C_in + lane;
  const bool run_inside = (r0 >= 0 && r0 + kWin <= a.T_in);   // warp-uniform: no row of this run is outside the input
  // Ragged mode: a run that starts in the padding produces only the constant row; its depthwise work is
  // skipped (its A rows stay stale — every output row depends on its own A row only) and the epilogue
  // writes the constant row for those frames.
  const bool run_needed = (r0 < cf);

  // B (pw^T) chunks are fetched one chunk AHEAD of their use: chunks 0 and 1 up front, chunk kc+1 in the shadow
  // of chunk kc's input loads (its stage is free once the MMAs of chunk kc-1 completed), so that thread 0 — and
  // with it the whole CTA at the next barrier — does not sit waiting for the copy it has only just issued.
  auto fetch_b = [&](int kc) {
    const int sb = kc & 1;
    mbar_expect_tx(bar_u + 8 * sb, bBytes);
    bulk_g2s(sB_u + sb * bBytes, bsrc + (size_t)kc * NT * kKC, bBytes, bar_u + 8 * sb);
  };
  if (tid == 0) {
    fetch_b(0);
    if (a.n_chunks > 1) fetch_b(1);
  }
  for (int kc = 0; kc < a.n_chunks; ++kc) {
    const int s = kc & 1, use = kc >> 1;
    if (kc >= 2) {                       // stage s is free once the MMAs of chunk kc-2 completed
      mbar_wait(bar_u + 8 * (2 + s), (use - 1) & 1);
      tc_fence_after();
    }
    // ---- depthwise for channels [c0, c0+32) -> A[s] ----------------------------------------
    const int c0 = kc * kKC;
    const int kvalid = min(kKC, C_in - c0);
    const bool cok = lane < kvalid;
    float v[kWin];
    float w[9];
    if (run_needed) {
    if (run_inside && kvalid == kKC) {   // common case: unpredicated loads at immediate offsets
      const float* xp = xrow + c0;
#pragma unroll
      for (int i = 0; i < kWin; ++i) v[i] = __ldg(xp + i * C_in);
#pragma unroll
      for (int k = 0; k < 9; ++k) w[k] = __ldg(a.dw + k * C_in + c0 + lane);
    } else {
#pragma unroll
      for (int i = 0; i < kWin; ++i)
        v[i] = (cok && r0 + i >= 0 && r0 + i < a.T_in) ? __ldg(xrow + c0 + (long long)i * C_in) : 0.0f;
#pragma unroll
      for (int k = 0; k < 9; ++k) w[k] = cok ? __ldg(a.dw + k * C_in + c0 + lane) : 0.0f;
    }
    if (fix) {                            // rows of the data get the gain and the floor; padding rows stay 0.0
      const float fl = (gc == gc) ? a.in_floor : gc;   // NaN peak -> NaN rows (fmaxf alone would drop the NaN)
      if (r0 + kWin <= cf) {
#pragma unroll
        for (int i = 0; i < kWin; ++i) v[i] = fmaxf(v[i] + gc, fl);
      } else {
#pragma unroll
        for (int i = 0; i < kWin; ++i) v[i] = (r0 + i < cf) ? fmaxf(v[i] + gc, fl) : v[i];
      }
    }
    }
    if (tid == 0 && kc >= 1 && kc + 1 < a.n_chunks) {   // (the loads above are in flight while this waits)
      mbar_wait(bar_u + 8 * (2 + (s ^ 1)), ((kc - 1) >> 1) & 1);   // MMAs of chunk kc-1 done: its B stage is free
      fetch_b(kc + 1);
    }
    if (run_needed) {
    unsigned char* As = sA + s * kABytes;
    float acc[kRun];
    depthwise_run16(v, w, acc);
#pragma unroll
    for (int j = 0; j < kRun; ++j) {
      const int row = warp * kRun + j;
      const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)lane >> 2) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)lane & 3u) * 4u;
      *reinterpret_cast<uint32_t*>(As + off) = to_tf32(acc[j]);
    }
    }
    fence_async_smem();                  // generic-proxy writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      mbar_wait(bar_u + 8 * s, use & 1); // B chunk has landed
      tc_fence_after();
      const uint64_t da = umma_desc_sw128(sA_u + s * kABytes);
      const uint64_t db = umma_desc_sw128(sB_u + s * bBytes);
      const int ksteps = kvalid >> 3;
      for (int k = 0; k < ksteps; ++k)
        umma_tf32(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
      umma_commit(bar_u + 8 * (2 + s));
      if (kc == a.n_chunks - 1) umma_commit(bar_u + 8 * 4);
    }
  }

  // ---- epilogue: TMEM -> bias + activation -> transpose in shared memory -> coalesced stores --
  mbar_wait(bar_u + 8 * 4, 0);
  tc_fence_after();
  {
    const int q = warp & 3, half = warp >> 2;
    float* stg = reinterpret_cast<float*>(sm) + warp * (32 * kStgStride);   // aliases A/B (all MMAs done)
    const int ngroups = NT >> 5;
    for (int g = half; g < ngroups; g += 2) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 32), r);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 bv = *reinterpret_cast<const float4*>(sBias + g * 32 + 4 * i);
        float4 o;
        act_apply2<ACT>(__uint_as_float(r[4 * i + 0]), __uint_as_float(r[4 * i + 1]), bv.x, bv.y, o.x, o.y);
        act_apply2<ACT>(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]), bv.z, bv.w, o.z, o.w);
        *reinterpret_cast<float4*>(stg + lane * kStgStride + 4 * i) = o;
      }
      __syncwarp();
{
          const int c4 = (lane & 7) * 4, r_lo = lane >> 3, tb = t0 + q * 32;
          float* yb = a.y + ((size_t)b * a.T_out + tb + r_lo) * a.C_out + n0 + g * 32 + c4;
          const float* sp = stg + r_lo * kStgStride + c4;
          if (2 * (tb + 31) < cf && tb + 31 < a.T_out) {      // warp-uniform: all 32 rows are data rows inside the tensor
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(yb + (size_t)(4 * i) * a.C_out) = *reinterpret_cast<const float4*>(sp + 4 * i * kStgStride);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int t = tb + r_lo + 4 * i;
              float4 o = *reinterpret_cast<const float4*>(sp + 4 * i * kStgStride);
              if (2 * t >= cf) o = __ldg(reinterpret_cast<const float4*>(a.pad_out + n0 + g * 32 + c4));
              if (t < a.T_out) *reinterpret_cast<float4*>(yb + (size_t)(4 * i) * a.C_out) = o;
            }
          }
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// pw [c_in, c_out] -> per (slice, chunk) shared-memory image of B = pw^T: row n (128 bytes) holds
// channels c0..c0+31 of output channel n, 16-byte groups XOR-swizzled by (n & 7), TF32-rounded.
__global__ void pack_pw_kernel(const float* __restrict__ pw, int C_in, int C_out, int NT, int n_chunks,
                               float* __restrict__ out) {
  const size_t total = (size_t)(C_out / NT) * n_chunks * NT * kKC;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cl = (int)(i % kKC);
    const int n = (int)((i / kKC) % NT);
    const int kc = (int)((i / ((size_t)kKC * NT)) % n_chunks);
    const int nh = (int)(i / ((size_t)kKC * NT * n_chunks));
    const int c = kc * kKC + cl;
    const float v = (c < C_in) ? pw[(size_t)c * C_out + nh * NT + n] : 0.0f;
    const size_t base = ((size_t)nh * n_chunks + kc) * NT * kKC;
    const int phys = n * kKC + ((((cl >> 2) ^ (n & 7)) << 2) | (cl & 3));
    out[base + phys] = __uint_as_float(to_tf32(v));
  }
}

typedef void (*SepKernel)(const SepArgs);
template <int CIN>
SepKernel pick_act(int act) {
  switch (act) {
    case TASR_ACT_TANH: return sepconv_tf32_kernel<CIN, TASR_ACT_TANH>;
    case TASR_ACT_GELU_ERF: return sepconv_tf32_kernel<CIN, TASR_ACT_GELU_ERF>;
    case TASR_ACT_RELU: return sepconv_tf32_kernel<CIN, TASR_ACT_RELU>;
    default: return sepconv_tf32_kernel<CIN, TASR_ACT_NONE>;
  }
}
SepKernel pick_kernel(int c_in, int act) {
  switch (c_in) {
    case 80: return pick_act<80>(act);
    case 192: return pick_act<192>(act);
    case 384: return pick_act<384>(act);
    case 288: return pick_act<288>(act);
    case 576: return pick_act<576>(act);
    default: return pick_act<0>(act);
  }
}

size_t smem_bytes(int NT) { return 1024 + 2 * kABytes + 2 * (size_t)NT * 128 + 256 * 4 + 128; }

}  // namespace

namespace tasr {
int validate_sepconv(const char* who, const void* x, int32_t B, int32_t T_in, const TasrSepConvLayer* L,
                     const void* y, int32_t T_out);
}

extern "C" int tasr_sepconv_plan_create(const TasrSepConvLayer* L, TasrSepConvPlan** out, tasr_stream_t stream) {
  if (!L || !out) return fail(TASR_ERR_BAD_ARG, "tasr_sepconv_plan_create: null argument");
  *out = nullptr;
  if (!L->dw || !L->pw || !L->bias) return fail(TASR_ERR_BAD_ARG, "tasr_sepconv_plan_create: null weight pointer");
  if (L->kernel != 9 || L->stride != 2)
    return fail(TASR_ERR_UNSUPPORTED, "tasr_sepconv_plan_create: kernels are built for kernel=9, stride=2; got k=%d s=%d",
                L->kernel, L->stride);
  if (L->activation < TASR_ACT_NONE || L->activation > TASR_ACT_RELU)
    return fail(TASR_ERR_BAD_ARG, "tasr_sepconv_plan_create: unknown activation %d", L->activation);
  if (L->c_in < 8 || (L->c_in & 7))
    return fail(TASR_ERR_UNSUPPORTED, "tasr_sepconv_plan_create: c_in=%d must be a positive multiple of 8 (UMMA K for TF32)", L->c_in);
  int n_split = 0;
  for (int s = 1; s <= 16; ++s) {
    if (L->c_out % s) continue;
    const int nt = L->c_out / s;
    if (nt <= 256 && nt >= 32 && (nt & 31) == 0) { n_split = s; break; }
  }
  if (!n_split)
    return fail(TASR_ERR_UNSUPPORTED, "tasr_sepconv_plan_create: c_out=%d cannot be cut into equal slices of <= 256 channels that are multiples of 32", L->c_out);
  TasrSepConvPlan* p = new TasrSepConvPlan();
  p->L = *L;
  p->n_split = n_split;
  p->NT = L->c_out / n_split;
  p->n_chunks = (L->c_in + kKC - 1) / kKC;
  p->d_bpack = nullptr;
  p->d_pad_in = nullptr;
  p->d_pad_out = nullptr;
  p->pad_ready = 0;
  {  // the persistent warp-specialised kernel of sepconv_ws.cu (identical bits) is the default: 49 / 70 / 57 us against
     // 56 / 90 / 69 us for the per-tile kernel on config 3; launches outside its limits (batch > 512, ...) fall back
     // to the per-tile kernel by themselves.  TASR_SEPCONV_WS=0 / 1: never / always try it.
    const char* e = getenv("TASR_SEPCONV_WS");
    p->use_ws = e ? (e[0] == '1' ? 1 : 0) : 1;
    const char* r = getenv("TASR_WS_ROLES");
    const int roles = r ? atoi(r) : 0;
    p->ws_roles = (roles == 18 || roles == 116) ? roles : 28;
  }
  int rc = check_cuda(cudaGetDevice(&p->device), "cudaGetDevice");
  if (rc == TASR_OK) rc = check_cuda(cudaMalloc(&p->d_pad_in, (size_t)9 * L->c_in * sizeof(float)), "cudaMalloc pad rows");
  if (rc == TASR_OK) rc = check_cuda(cudaMalloc(&p->d_pad_out, (size_t)L->c_out * sizeof(float)), "cudaMalloc pad row");
  const size_t n = (size_t)n_split * p->n_chunks * p->NT * kKC;
  if (rc == TASR_OK) rc = check_cuda(cudaMalloc(&p->d_bpack, n * sizeof(float)), "cudaMalloc packed pointwise weights");
  if (rc == TASR_OK) {
    pack_pw_kernel<<<(unsigned)((n + 255) / 256 > 1024 ? 1024 : (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        L->pw, L->c_in, L->c_out, p->NT, p->n_chunks, p->d_bpack);
    count_launch();
    rc = check_cuda(cudaGetLastError(), "pack_pw_kernel");
  }
  if (rc == TASR_OK) {
    p->kernel = reinterpret_cast<void*>(pick_kernel(L->c_in, L->activation));
    rc = check_cuda(cudaFuncSetAttribute(p->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(256)),
                    "cudaFuncSetAttribute(sepconv_tf32_kernel)");
  }
  if (rc != TASR_OK) { tasr_sepconv_plan_destroy(p); return rc; }
  *out = p;
  return TASR_OK;
}

extern "C" int tasr_sepconv_plan_destroy(TasrSepConvPlan* p) {
  if (!p) return TASR_OK;
  cudaFree(p->d_bpack);
  cudaFree(p->d_pad_in);
  cudaFree(p->d_pad_out);
  delete p;
  return TASR_OK;
}

static int launch_tf32(const char* who, const TasrSepConvPlan* p, const float* x, const int32_t* len0, int32_t shift,
                       int32_t B, int32_t T_in, float* y, int32_t T_out, tasr_stream_t stream, int32_t fill_rows = -1,
                       const TasrDeferredGain* gain = nullptr) {
  if (!p) return fail(TASR_ERR_BAD_ARG, "%s: null plan", who);
  int rc = validate_sepconv(who, x, B, T_in, &p->L, y, T_out);
  if (rc != TASR_OK) return rc;
  int dev = 0;
  TASR_CUDA(cudaGetDevice(&dev));
  if (dev != p->device) return fail(TASR_ERR_BAD_ARG, "%s: plan was created on device %d, current device is %d", who, p->device, dev);
  if (B == 0 || T_out == 0) return TASR_OK;
  SepArgs a;
  a.x = x; a.dw = p->L.dw; a.bpack = p->d_bpack; a.bias = p->L.bias; a.y = y;
  a.T_in = T_in; a.T_out = T_out; a.C_in = p->L.c_in; a.C_out = p->L.c_out; a.NT = p->NT;
  a.n_chunks = p->n_chunks; a.act = p->L.activation;
  a.len0 = len0; a.pad_out = p->d_pad_out; a.shift = shift; a.fill_rows = fill_rows;
  a.in_peak = gain ? gain->peak : nullptr;
  a.in_scale2 = gain ? gain->log_scale_x2 : 0.0f;
  a.in_floor = gain ? gain->log_floor : 0.0f;
  a.row_off = 0;
  if (p->L.same) {
    // padding='same' (the reference constructor's default, encoder.py:24): dense mode only — the ragged bookkeeping
    // (constant padding rows, tiles skipped by length) is written for 'valid' receptive fields
    if (len0 != nullptr)
      return fail(TASR_ERR_UNSUPPORTED, "%s: padding='same' layers run in dense mode (tasr_sepconv1d_tf32), not ragged", who);
    a.row_off = -tasr_same_pad_left(T_in, p->L.kernel, p->L.stride);
  }
  if (p->use_ws) {
    const int wrc = tasr_sepconv_ws_launch(p, a, B, (cudaStream_t)stream);
    if (wrc >= 0) return wrc;            // launched (TASR_OK) or failed with an error code
  }
  dim3 grid((T_out + kMT - 1) / kMT, p->n_split, B);
  reinterpret_cast<SepKernel>(p->kernel)<<<grid, kThreads, smem_bytes(p->NT), (cudaStream_t)stream>>>(a);
  TASR_LAUNCH_CHECK("sepconv_tf32_kernel");
  return TASR_OK;
}

extern "C" int tasr_sepconv1d_tf32(const TasrSepConvPlan* p, const float* x, int32_t B, int32_t T_in,
                                   float* y, int32_t T_out, tasr_stream_t stream) {
  return launch_tf32("tasr_sepconv1d_tf32", p, x, nullptr, 0, B, T_in, y, T_out, stream);
}

extern "C" int tasr_sepconv_plan_set_pad_row(TasrSepConvPlan* p, const float* pad_row_in, tasr_stream_t stream) {
  if (!p) return fail(TASR_ERR_BAD_ARG, "tasr_sepconv_plan_set_pad_row: null plan");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t row = (size_t)p->L.c_in * sizeof(float);
  if (pad_row_in) {
    for (int k = 0; k < 9; ++k)
      TASR_CUDA(cudaMemcpyAsync(p->d_pad_in + (size_t)k * p->L.c_in, pad_row_in, row, cudaMemcpyDeviceToDevice, st));
  } else {
    TASR_CUDA(cudaMemsetAsync(p->d_pad_in, 0, 9 * row, st));
  }
  // The padding row of the output is whatever THIS kernel computes for a receptive field made of the
  // input's padding row (one output frame from nine identical rows), so filled and computed rows are
  // bit-identical and the constant propagates exactly through the stack.
  int rc = launch_tf32("tasr_sepconv_plan_set_pad_row", p, p->d_pad_in, nullptr, 0, 1, 9, p->d_pad_out, 1, stream);
  if (rc == TASR_OK) p->pad_ready = 1;
  return rc;
}

extern "C" const float* tasr_sepconv_plan_pad_row(const TasrSepConvPlan* p) { return (p && p->pad_ready) ? p->d_pad_out : nullptr; }

extern "C" int tasr_sepconv1d_tf32_ragged(const TasrSepConvPlan* p, const float* x, const int32_t* len0, int32_t shift,
                                          int32_t B, int32_t T_in, float* y, int32_t T_out, tasr_stream_t stream) {
  if (!p || !len0) return fail(TASR_ERR_BAD_ARG, "tasr_sepconv1d_tf32_ragged: null argument");
  if (shift < 0 || shift > 30) return fail(TASR_ERR_BAD_ARG, "tasr_sepconv1d_tf32_ragged: shift must be in [0,30]");
  if (!p->pad_ready)
    return fail(TASR_ERR_BAD_ARG, "tasr_sepconv1d_tf32_ragged: call tasr_sepconv_plan_set_pad_row first (the input's padding row is unknown)");
  return launch_tf32("tasr_sepconv1d_tf32_ragged", p, x, len0, shift, B, T_in, y, T_out, stream);
}

extern "C" int32_t tasr_sepconv_ragged_margin(void) { return kWin; }

extern "C" int tasr_sepconv1d_tf32_ragged_lean(const TasrSepConvPlan* p, const float* x, const int32_t* len0, int32_t shift,
                                               int32_t B, int32_t T_in, float* y, int32_t T_out, int32_t fill_rows,
                                               const TasrDeferredGain* gain, tasr_stream_t stream) {
  if (!p || !len0) return fail(TASR_ERR_BAD_ARG, "tasr_sepconv1d_tf32_ragged_lean: null argument");
  if (shift < 0 || shift > 29) return fail(TASR_ERR_BAD_ARG, "tasr_sepconv1d_tf32_ragged_lean: shift must be in [0,29]");
  if (gain && (shift != 0 || !gain->peak))
    return fail(TASR_ERR_BAD_ARG, "tasr_sepconv1d_tf32_ragged_lean: a deferred input gain belongs to the first layer (shift 0) and needs a peak pointer");
  if (!p->pad_ready)
    return fail(TASR_ERR_BAD_ARG, "tasr_sepconv1d_tf32_ragged_lean: call tasr_sepconv_plan_set_pad_row first (the input's padding row is unknown)");
  return launch_tf32("tasr_sepconv1d_tf32_ragged_lean", p, x, len0, shift, B, T_in, y, T_out, stream, fill_rows < 0 ? -1 : fill_rows, gain);
}
