// First encoder block of the Moonshine-style encoder, sm_100a only.
//
// Replaces EncoderBlock.call (src/models/moonshine/encoder.py:151-154): MHSAModule (src/models/layers/attention.py:519-602:
// bias-free q/k/v projections, RoPE on q and k (positional_encoding.py:20-93), q / sqrt(head_dim), Keras masked softmax with
// the padding mask on queries and keys, @ v, output projection, x + ., LayerNorm) followed by FFNModule
// (src/models/layers/mlp.py:9-60: dense1 + exact-erf GELU, dense2, + input, LayerNorm).  Inference only (dropout = identity).
//
//   linear_tf32_kernel   the four dense layers on the tensor cores: tile = 128 tokens x 192 outputs, the activations rounded
//                        to TF32 (cvt.rna) on their way into the UMMA K-major SWIZZLE_128B A tile, weights pre-packed per
//                        32-input chunk as shared-memory images fetched by cp.async.bulk, tcgen05.mma kind::tf32 with FP32
//                        accumulators in TMEM, two A/B stages.  Epilogues out of TMEM:
//                          QKV      RoPE (interleaved pairs, position = token index) on the q and k slices, q scaled;
//                          GELU     + bias, exact-erf GELU;
//                          RES_LN   (+ bias) + residual, LayerNormalization over the 192 outputs of the token (two passes
//                                   over the accumulator row: moments, then normalise).
//   attention_kernel     one CTA per (utterance, head): K and V of the head in shared memory, a thread per query row with
//                        an online softmax over the utterance's valid keys (FP32 CUDA cores; 6 % of the block's FLOPs);
//                        padded query rows get the uniform average of v over all T keys, which is what the reference's
//                        "+ (1 - mask) * -1e9" leaves of a fully masked row in float32.
#include "sepconv_common.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

using namespace tasr;
using namespace tasr_sep;

struct TasrEncoderBlockPlan {
  TasrEncoderBlockWeights W;    // LayerNorm / bias pointers are borrowed from the caller (must outlive the plan)
  int device;
  int d, hd, F;                 // model dim, heads * head_dim, FFN width
  float* d_pack_qkv;            // [3][d/32][192*32]
  float* d_pack_o;              // [1][hd/32][192*32]
  float* d_pack_1;              // [F/192][d/32][192*32]
  float* d_pack_2;              // [1][F/32][192*32]
  float2* d_rope;               // [rope_T][16] (cos, sin)
  int rope_T;
};

namespace {

constexpr int kNT = 192;
enum { EPI_QKV = 0, EPI_BIAS_GELU = 1, EPI_RES_LN = 2 };

struct LinArgs {
  const float* x;        // [M, K]
  const float* bpack;    // [n_split][K/32][192*32]
  const float* bias;     // [N] or null
  float* y;              // [M, N]   (EPI_QKV: q [M,192]; y2 = k, y3 = v)
  float* y2;
  float* y3;
  const float* res;      // [M, 192] residual (EPI_RES_LN)
  const float* gamma;
  const float* beta;
  const float2* rope;    // [T][16]
  int32_t M, K, N, n_chunks, T;
  float eps, qscale;
};

template <int EPI>
__global__ void __launch_bounds__(kThreads, 2) linear_tf32_kernel(const LinArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t bBytes = (uint32_t)kNT * 128u;

  unsigned char* sA = sm;                          // 2 x 16 KiB
  unsigned char* sB = sm + 2 * kABytes;            // 2 x 24 KiB
  float* sVec = reinterpret_cast<float*>(sB + 2 * bBytes);    // bias | gamma | beta (3 x 192)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sVec + 3 * kNT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB), bar_u = smem_u32(bars);

  const int nh = blockIdx.y, m0 = blockIdx.x * kMT, n0 = nh * kNT;

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  if (tid == 32) {
    for (int i = 0; i < 5; ++i) mbar_init(bar_u + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < kNT; i += kThreads) {
    sVec[i] = a.bias ? a.bias[n0 + i] : 0.0f;
    if (EPI == EPI_RES_LN) { sVec[kNT + i] = a.gamma[i]; sVec[2 * kNT + i] = a.beta[i]; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const float* bsrc = a.bpack + (size_t)nh * a.n_chunks * kNT * kKC;
  const uint32_t idesc = umma_idesc_tf32(kMT, kNT);
  auto fetch_b = [&](int kc) {
    const int sb = kc & 1;
    mbar_expect_tx(bar_u + 8 * sb, bBytes);
    bulk_g2s(sB_u + sb * bBytes, bsrc + (size_t)kc * kNT * kKC, bBytes, bar_u + 8 * sb);
  };
  if (tid == 0) {
    fetch_b(0);
    if (a.n_chunks > 1) fetch_b(1);
  }
  const int rw = m0 + warp * kRun;                       // this warp's 16 token rows
  const float* xrow = a.x + (size_t)rw * a.K + lane;
  for (int kc = 0; kc < a.n_chunks; ++kc) {
    const int s = kc & 1, use = kc >> 1;
    if (kc >= 2) {                       // stage s is free once the MMAs of chunk kc-2 completed
      mbar_wait(bar_u + 8 * (2 + s), (use - 1) & 1);
      tc_fence_after();
    }
    float v[kRun];
#pragma unroll
    for (int j = 0; j < kRun; ++j) v[j] = (rw + j < a.M) ? __ldg(xrow + (size_t)j * a.K + kc * kKC) : 0.0f;
    if (tid == 0 && kc >= 1 && kc + 1 < a.n_chunks) {
      mbar_wait(bar_u + 8 * (2 + (s ^ 1)), ((kc - 1) >> 1) & 1);   // MMAs of chunk kc-1 done: its B stage is free
      fetch_b(kc + 1);
    }
    unsigned char* As = sA + s * kABytes;
#pragma unroll
    for (int j = 0; j < kRun; ++j) {
      const int row = warp * kRun + j;
      const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)lane >> 2) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)lane & 3u) * 4u;
      *reinterpret_cast<uint32_t*>(As + off) = to_tf32(v[j]);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(bar_u + 8 * s, use & 1);
      tc_fence_after();
      const uint64_t da = umma_desc_sw128(sA_u + s * kABytes);
      const uint64_t db = umma_desc_sw128(sB_u + s * bBytes);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_tf32(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
      umma_commit(bar_u + 8 * (2 + s));
      if (kc == a.n_chunks - 1) umma_commit(bar_u + 8 * 4);
    }
  }

  // ---- epilogue ------------------------------------------------------------------------------------------------
  mbar_wait(bar_u + 8 * 4, 0);
  tc_fence_after();
  {
    const int q = warp & 3, half = warp >> 2;
    float* stg = reinterpret_cast<float*>(sm) + warp * (32 * kStgStride);   // aliases A/B (all MMAs done)
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int m = m0 + q * 32 + lane;                  // this thread's token row
    float mean = 0.0f, rstd = 0.0f;
    if (EPI == EPI_RES_LN) {
      // moments of z = acc + bias + residual over the 192 outputs of the row (tf.nn.moments: mean, then mean((z - mean)^2))
      const float* rr = a.res + (size_t)min(m, a.M - 1) * kNT;
      float sum = 0.0f;
      for (int g = 0; g < kNT / 32; ++g) {
        uint32_t r[32];
        tmem_ld32(trow + (uint32_t)(g * 32), r);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 rv = __ldg(reinterpret_cast<const float4*>(rr + g * 32 + 4 * i));
          sum += (__uint_as_float(r[4 * i]) + sVec[g * 32 + 4 * i] + rv.x) + (__uint_as_float(r[4 * i + 1]) + sVec[g * 32 + 4 * i + 1] + rv.y) +
                 (__uint_as_float(r[4 * i + 2]) + sVec[g * 32 + 4 * i + 2] + rv.z) + (__uint_as_float(r[4 * i + 3]) + sVec[g * 32 + 4 * i + 3] + rv.w);
        }
      }
      mean = sum * (1.0f / kNT);
      float var = 0.0f;
      for (int g = 0; g < kNT / 32; ++g) {
        uint32_t r[32];
        tmem_ld32(trow + (uint32_t)(g * 32), r);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 rv = __ldg(reinterpret_cast<const float4*>(rr + g * 32 + 4 * i));
          const float z0 = __uint_as_float(r[4 * i]) + sVec[g * 32 + 4 * i] + rv.x - mean;
          const float z1 = __uint_as_float(r[4 * i + 1]) + sVec[g * 32 + 4 * i + 1] + rv.y - mean;
          const float z2 = __uint_as_float(r[4 * i + 2]) + sVec[g * 32 + 4 * i + 2] + rv.z - mean;
          const float z3 = __uint_as_float(r[4 * i + 3]) + sVec[g * 32 + 4 * i + 3] + rv.w - mean;
          var = fmaf(z0, z0, fmaf(z1, z1, fmaf(z2, z2, fmaf(z3, z3, var))));
        }
      }
      rstd = rsqrtf(var * (1.0f / kNT) + a.eps);
    }
    for (int g = half; g < kNT / 32; g += 2) {
      uint32_t r[32];
      tmem_ld32(trow + (uint32_t)(g * 32), r);
      __syncwarp();
      if (EPI == EPI_QKV) {
        // one 32-column group = one head (head_dim 32 = rot_dim): out[2i] = x[2i] c - x[2i+1] s, out[2i+1] = x[2i+1] c + x[2i] s
        const bool rot = (nh < 2);
        const float sc = (nh == 0) ? a.qscale : 1.0f;
        const float2* cs = a.rope + (size_t)(min(m, a.M - 1) % a.T) * 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 o = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
          if (rot) {
            const float2 c0 = __ldg(cs + 2 * i), c1 = __ldg(cs + 2 * i + 1);
            const float a0 = o.x * c0.x - o.y * c0.y, b0 = o.y * c0.x + o.x * c0.y;
            const float a1 = o.z * c1.x - o.w * c1.y, b1 = o.w * c1.x + o.z * c1.y;
            o = make_float4(a0 * sc, b0 * sc, a1 * sc, b1 * sc);
          }
          *reinterpret_cast<float4*>(stg + lane * kStgStride + 4 * i) = o;
        }
      } else if (EPI == EPI_BIAS_GELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bv = *reinterpret_cast<const float4*>(sVec + g * 32 + 4 * i);
          float4 o;
          act_apply2<TASR_ACT_GELU_ERF>(__uint_as_float(r[4 * i + 0]), __uint_as_float(r[4 * i + 1]), bv.x, bv.y, o.x, o.y);
          act_apply2<TASR_ACT_GELU_ERF>(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]), bv.z, bv.w, o.z, o.w);
          *reinterpret_cast<float4*>(stg + lane * kStgStride + 4 * i) = o;
        }
      } else {
        const float* rr = a.res + (size_t)min(m, a.M - 1) * kNT;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 rv = __ldg(reinterpret_cast<const float4*>(rr + g * 32 + 4 * i));
          const float4 bv = *reinterpret_cast<const float4*>(sVec + g * 32 + 4 * i);
          const float4 gm = *reinterpret_cast<const float4*>(sVec + kNT + g * 32 + 4 * i);
          const float4 bt = *reinterpret_cast<const float4*>(sVec + 2 * kNT + g * 32 + 4 * i);
          float4 o;
          o.x = fmaf((__uint_as_float(r[4 * i]) + bv.x + rv.x - mean) * rstd, gm.x, bt.x);
          o.y = fmaf((__uint_as_float(r[4 * i + 1]) + bv.y + rv.y - mean) * rstd, gm.y, bt.y);
          o.z = fmaf((__uint_as_float(r[4 * i + 2]) + bv.z + rv.z - mean) * rstd, gm.z, bt.z);
          o.w = fmaf((__uint_as_float(r[4 * i + 3]) + bv.w + rv.w - mean) * rstd, gm.w, bt.w);
          *reinterpret_cast<float4*>(stg + lane * kStgStride + 4 * i) = o;
        }
      }
      __syncwarp();
      {
        float* ybase = (EPI == EPI_QKV) ? (nh == 0 ? a.y : nh == 1 ? a.y2 : a.y3) : a.y;
        const int ldy = (EPI == EPI_QKV) ? kNT : a.N;
        const int ncol = (EPI == EPI_QKV) ? 0 : n0;
        const int c4 = (lane & 7) * 4, r_lo = lane >> 3, mb = m0 + q * 32;
        float* yb = ybase + (size_t)(mb + r_lo) * ldy + ncol + g * 32 + c4;
        const float* sp = stg + r_lo * kStgStride + c4;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (mb + r_lo + 4 * i < a.M)
            *reinterpret_cast<float4*>(yb + (size_t)(4 * i) * ldy) = *reinterpret_cast<const float4*>(sp + 4 * i * kStgStride);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// W [K, N] -> per (slice of 192 outputs, chunk of 32 inputs) shared-memory image of B = W^T: row n (128 bytes) holds
// inputs c0..c0+31 of output n, 16-byte groups XOR-swizzled by (n & 7), TF32-rounded.  (W = [W_a | W_b | ...] when several
// matrices are concatenated along N: `srcs`.)
__global__ void pack_linear_kernel(const float* w0, const float* w1, const float* w2, int K, int N_each, int n_split, int n_chunks,
                                   float* __restrict__ out) {
  const size_t total = (size_t)n_split * n_chunks * kNT * kKC;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cl = (int)(i % kKC);
    const int n = (int)((i / kKC) % kNT);
    const int kc = (int)((i / ((size_t)kKC * kNT)) % n_chunks);
    const int nh = (int)(i / ((size_t)kKC * kNT * n_chunks));
    const int c = kc * kKC + cl;
    const int ncol = nh * kNT + n;                     // column of the concatenated matrix
    const int which = ncol / N_each, col = ncol - which * N_each;
    const float* w = which == 0 ? w0 : which == 1 ? w1 : w2;
    const float v = (c < K) ? w[(size_t)c * N_each + col] : 0.0f;
    const size_t base = ((size_t)nh * n_chunks + kc) * kNT * kKC;
    const int phys = n * kKC + ((((cl >> 2) ^ (n & 7)) << 2) | (cl & 3));
    out[base + phys] = __uint_as_float(to_tf32(v));
  }
}

// One CTA per (utterance, head): thread = query row, K / V of the head in shared memory (rows of 32 floats, read as
// warp-wide broadcasts), online softmax over blocks of 8 keys.
struct AttArgs {
  const float* q;        // [B, T, H*32]  (RoPE applied, scaled)
  const float* k;        // [B, T, H*32]  (RoPE applied)
  const float* v;        // [B, T, H*32]
  const int32_t* len;    // [B] valid tokens per utterance, or null (all T)
  float* out;            // [B, T, H*32]
  int32_t B, T, H, causal;
};
constexpr int kAttThreads = 192;   // T3 = 181 (15 s) in one round of query rows, 368 (30 s) in two
__global__ void __launch_bounds__(kAttThreads) attention_kernel(const AttArgs a) {
  extern __shared__ __align__(16) unsigned char att_smem[];
  float* sK = reinterpret_cast<float*>(att_smem);
  float* sV = sK + (size_t)a.T * 32;
  float* sMean = sV + (size_t)a.T * 32;
  const int b = blockIdx.x / a.H, h = blockIdx.x - b * a.H, tid = threadIdx.x;
  const int ld = a.H * 32;
  const int L = a.len ? max(0, min(a.len[b], a.T)) : a.T;
  const float* kb = a.k + (size_t)b * a.T * ld + h * 32;
  const float* vb = a.v + (size_t)b * a.T * ld + h * 32;
  for (int i = tid; i < a.T * 8; i += kAttThreads) {
    const int j = i >> 3, c = (i & 7) * 4;
    *reinterpret_cast<float4*>(sK + j * 32 + c) = __ldg(reinterpret_cast<const float4*>(kb + (size_t)j * ld + c));
    *reinterpret_cast<float4*>(sV + j * 32 + c) = __ldg(reinterpret_cast<const float4*>(vb + (size_t)j * ld + c));
  }
  __syncthreads();
  if (L < a.T && tid < 32) {               // padded query rows: uniform weights over all T keys
    float s = 0.0f;
    for (int j = 0; j < a.T; ++j) s += sV[j * 32 + tid];
    sMean[tid] = s / (float)a.T;
  }
  __syncthreads();
  for (int t = tid; t < a.T; t += kAttThreads) {
    float* op = a.out + ((size_t)b * a.T + t) * ld + h * 32;
    if (t >= L) {
#pragma unroll
      for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(op + c) = *reinterpret_cast<const float4*>(sMean + c);
      continue;
    }
    float q[32], acc[32];
    const float* qp = a.q + ((size_t)b * a.T + t) * ld + h * 32;
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
      const float4 v4 = __ldg(reinterpret_cast<const float4*>(qp + c));
      q[c] = v4.x; q[c + 1] = v4.y; q[c + 2] = v4.z; q[c + 3] = v4.w;
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.0f;
    float mx = -INFINITY, den = 0.0f;
    const int jmax = a.causal ? min(L, t + 1) : L;
    for (int j0 = 0; j0 < jmax; j0 += 8) {
      float s[8];
      float bm = mx;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        float d = 0.0f;
        if (j < jmax) {
          const float4* kr = reinterpret_cast<const float4*>(sK + j * 32);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 kv = kr[c];
            d = fmaf(q[4 * c], kv.x, fmaf(q[4 * c + 1], kv.y, fmaf(q[4 * c + 2], kv.z, fmaf(q[4 * c + 3], kv.w, d))));
          }
        } else {
          d = -INFINITY;
        }
        s[u] = d;
        bm = fmaxf(bm, d);
      }
      const float corr = __expf(mx - bm);            // exp(-inf) = 0 on the first block
      den *= corr;
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] *= corr;
      mx = bm;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        if (j < jmax) {
          const float p = __expf(s[u] - mx);
          den += p;
          const float4* vr = reinterpret_cast<const float4*>(sV + j * 32);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 vv = vr[c];
            acc[4 * c] = fmaf(p, vv.x, acc[4 * c]);
            acc[4 * c + 1] = fmaf(p, vv.y, acc[4 * c + 1]);
            acc[4 * c + 2] = fmaf(p, vv.z, acc[4 * c + 2]);
            acc[4 * c + 3] = fmaf(p, vv.w, acc[4 * c + 3]);
          }
        }
      }
    }
    const float inv = 1.0f / den;
#pragma unroll
    for (int c = 0; c < 32; c += 4)
      *reinterpret_cast<float4*>(op + c) = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
  }
}

size_t lin_smem() { return 1024 + 2 * kABytes + 2 * (size_t)kNT * 128 + 3 * kNT * 4 + 128; }

template <int EPI>
int launch_linear(const LinArgs& a, int n_split, cudaStream_t st) {
  static bool attr_set[64] = {false};
  int dev = 0;
  TASR_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    TASR_CUDA(cudaFuncSetAttribute(linear_tf32_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin_smem()));
    attr_set[dev] = true;
  }
  dim3 grid((unsigned)((a.M + kMT - 1) / kMT), (unsigned)n_split);
  linear_tf32_kernel<EPI><<<grid, kThreads, lin_smem(), st>>>(a);
  TASR_LAUNCH_CHECK("linear_tf32_kernel");
  return TASR_OK;
}

int pack(const float* w0, const float* w1, const float* w2, int K, int N_each, int n_mats, float** out, cudaStream_t st) {
  const int n_split = N_each * n_mats / kNT, n_chunks = (K + kKC - 1) / kKC;
  const size_t n = (size_t)n_split * n_chunks * kNT * kKC;
  TASR_CUDA(cudaMalloc(out, n * sizeof(float)));
  pack_linear_kernel<<<(unsigned)((n + 255) / 256 > 1024 ? 1024 : (n + 255) / 256), 256, 0, st>>>(w0, w1, w2, K, N_each, n_split, n_chunks, *out);
  return check_cuda(cudaGetLastError(), "pack_linear_kernel");
}

}  // namespace

extern "C" int tasr_encoder_block_plan_create(const TasrEncoderBlockWeights* W, TasrEncoderBlockPlan** out, tasr_stream_t stream) {
  if (!W || !out) return fail(TASR_ERR_BAD_ARG, "tasr_encoder_block_plan_create: null argument");
  *out = nullptr;
  if (!W->wq || !W->wk || !W->wv || !W->wo || !W->ln1_gamma || !W->ln1_beta || !W->w1 || !W->b1 || !W->w2 || !W->b2 || !W->ln2_gamma || !W->ln2_beta)
    return fail(TASR_ERR_BAD_ARG, "tasr_encoder_block_plan_create: null weight pointer");
  if (W->head_dim != 32 || W->d_model != kNT || W->num_heads * W->head_dim != kNT || W->fc_factor < 1 || W->fc_factor > 8)
    return fail(TASR_ERR_UNSUPPORTED,
                "tasr_encoder_block_plan_create: kernels are specialised for d_model = num_heads * head_dim = 192, head_dim = 32 "
                "(config/model.yaml: 6 x 32; rot_dim = max(head_dim // 2, 32) = head_dim), fc_factor 1..8; got %d / %d x %d / %d",
                W->d_model, W->num_heads, W->head_dim, W->fc_factor);
  TasrEncoderBlockPlan* p = new TasrEncoderBlockPlan();
  memset(p, 0, sizeof(*p));
  p->W = *W;
  p->d = W->d_model; p->hd = W->num_heads * W->head_dim; p->F = W->d_model * W->fc_factor;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = check_cuda(cudaGetDevice(&p->device), "cudaGetDevice");
  if (rc == TASR_OK) rc = pack(W->wq, W->wk, W->wv, p->d, p->hd, 3, &p->d_pack_qkv, st);
  if (rc == TASR_OK) rc = pack(W->wo, nullptr, nullptr, p->hd, p->d, 1, &p->d_pack_o, st);
  if (rc == TASR_OK) rc = pack(W->w1, nullptr, nullptr, p->d, p->F, 1, &p->d_pack_1, st);
  if (rc == TASR_OK) rc = pack(W->w2, nullptr, nullptr, p->F, p->d, 1, &p->d_pack_2, st);
  if (rc == TASR_OK) rc = check_cuda(cudaStreamSynchronize(st), "pack weights");
  if (rc != TASR_OK) { tasr_encoder_block_plan_destroy(p); return rc; }
  *out = p;
  return TASR_OK;
}

extern "C" int tasr_encoder_block_plan_destroy(TasrEncoderBlockPlan* p) {
  if (!p) return TASR_OK;
  cudaFree(p->d_pack_qkv);
  cudaFree(p->d_pack_o);
  cudaFree(p->d_pack_1);
  cudaFree(p->d_pack_2);
  cudaFree(p->d_rope);
  delete p;
  return TASR_OK;
}

extern "C" int64_t tasr_encoder_block_workspace_floats(const TasrEncoderBlockPlan* p, int32_t B, int32_t T) {
  if (!p || B < 0 || T < 0) return -1;
  return (int64_t)B * T * (5 * (int64_t)kNT + p->F);       // q, k, v, attention output, h1, FFN hidden
}

// RoPE table for positions 0..T-1 with the reference's float32 arithmetic (positional_encoding.py:13-17, 48-52, 79-80)
static int ensure_rope(TasrEncoderBlockPlan* p, int T) {
  if (p->rope_T >= T) return TASR_OK;
  int cap = p->rope_T > 0 ? p->rope_T : 512;
  while (cap < T) cap *= 2;
  std::vector<float2> tab((size_t)cap * 16);
  for (int i = 0; i < 16; ++i) {
    const float inv_freq = 1.0f / powf(10000.0f, (float)(2 * i) / 32.0f);
    for (int t = 0; t < cap; ++t) {
      const float f = (float)t * inv_freq;
      tab[(size_t)t * 16 + i] = make_float2(cosf(f), sinf(f));
    }
  }
  float2* d = nullptr;
  TASR_CUDA(cudaMalloc(&d, tab.size() * sizeof(float2)));
  TASR_CUDA(cudaMemcpy(d, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice));
  cudaFree(p->d_rope);
  p->d_rope = d;
  p->rope_T = cap;
  return TASR_OK;
}

extern "C" int tasr_encoder_block_prepare(TasrEncoderBlockPlan* p, int32_t T_max) {
  if (!p || T_max < 0) return fail(TASR_ERR_BAD_ARG, "tasr_encoder_block_prepare: bad argument");
  return ensure_rope(p, T_max);
}

extern "C" int tasr_encoder_block_f32(TasrEncoderBlockPlan* p, const float* x, const int32_t* len, int32_t B, int32_t T,
                                      int32_t use_causal_mask, float* workspace, float* out, tasr_stream_t stream) {
  if (!p || !x || !workspace || !out) return fail(TASR_ERR_BAD_ARG, "tasr_encoder_block_f32: null argument");
  if (B < 0 || T < 0) return fail(TASR_ERR_BAD_ARG, "tasr_encoder_block_f32: negative size");
  if (!aligned16(x) || !aligned16(workspace) || !aligned16(out))
    return fail(TASR_ERR_MISALIGNED, "tasr_encoder_block_f32: x / workspace / out must be 16-byte aligned");
  if (B == 0 || T == 0) return TASR_OK;
  int dev = 0;
  TASR_CUDA(cudaGetDevice(&dev));
  if (dev != p->device) return fail(TASR_ERR_BAD_ARG, "tasr_encoder_block_f32: plan was created on device %d, current device is %d", p->device, dev);
  const size_t att_smem = ((size_t)2 * T * 32 + 32) * sizeof(float);
  if (att_smem > 200 * 1024) return fail(TASR_ERR_UNSUPPORTED, "tasr_encoder_block_f32: T = %d exceeds the attention kernel's shared-memory K/V (T <= 790)", T);
  if (p->rope_T < T) {
    // (allocates: call tasr_encoder_block_prepare(plan, T_max) once before capturing a CUDA graph)
    const int rc = ensure_rope(p, T);
    if (rc != TASR_OK) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long M64 = (long long)B * T;
  if (M64 > 0x7fffffffLL / 1024) return fail(TASR_ERR_UNSUPPORTED, "tasr_encoder_block_f32: too many tokens");
  const int M = (int)M64;
  float* q = workspace;
  float* k = q + (size_t)M * kNT;
  float* v = k + (size_t)M * kNT;
  float* ao = v + (size_t)M * kNT;
  float* h1 = ao + (size_t)M * kNT;
  float* f1 = h1 + (size_t)M * kNT;

  LinArgs a = {};
  a.M = M; a.T = T; a.eps = p->W.ln_eps; a.qscale = 1.0f / sqrtf((float)p->W.head_dim);
  // q, k, v = x Wq, x Wk, x Wv; RoPE; q / sqrt(head_dim)          attention.py:86-90, 190-191, 102
  a.x = x; a.bpack = p->d_pack_qkv; a.bias = nullptr; a.y = q; a.y2 = k; a.y3 = v; a.K = p->d; a.N = 3 * kNT; a.n_chunks = p->d / kKC;
  a.rope = p->d_rope;
  int rc = launch_linear<EPI_QKV>(a, 3, st);
  if (rc != TASR_OK) return rc;
  {
    AttArgs t = {q, k, v, len, ao, B, T, p->W.num_heads, use_causal_mask ? 1 : 0};
    static bool attr_set[64] = {false};
    if (dev < 64 && !attr_set[dev]) {
      TASR_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set[dev] = true;
    }
    attention_kernel<<<(unsigned)(B * p->W.num_heads), kAttThreads, att_smem, st>>>(t);
    TASR_LAUNCH_CHECK("attention_kernel");
  }
  // h1 = LayerNorm(x + attention Wo)                               attention.py:120, 598-600
  a.x = ao; a.bpack = p->d_pack_o; a.bias = nullptr; a.y = h1; a.res = x; a.gamma = p->W.ln1_gamma; a.beta = p->W.ln1_beta;
  a.K = p->hd; a.N = kNT; a.n_chunks = p->hd / kKC;
  rc = launch_linear<EPI_RES_LN>(a, 1, st);
  if (rc != TASR_OK) return rc;
  // f1 = gelu(h1 W1 + b1)                                          mlp.py:51
  a.x = h1; a.bpack = p->d_pack_1; a.bias = p->W.b1; a.y = f1; a.K = p->d; a.N = p->F; a.n_chunks = p->d / kKC;
  rc = launch_linear<EPI_BIAS_GELU>(a, p->F / kNT, st);
  if (rc != TASR_OK) return rc;
  // out = LayerNorm(f1 W2 + b2 + h1)                               mlp.py:53-55
  a.x = f1; a.bpack = p->d_pack_2; a.bias = p->W.b2; a.y = out; a.res = h1; a.gamma = p->W.ln2_gamma; a.beta = p->W.ln2_beta;
  a.K = p->F; a.N = kNT; a.n_chunks = p->F / kKC;
  return launch_linear<EPI_RES_LN>(a, 1, st);
}
