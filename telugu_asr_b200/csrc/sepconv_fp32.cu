// SeparableConv1D forward, FP32 on CUDA cores (TASR_MATH_FP32).
//
// Replaces one tf.keras.layers.SeparableConv1D call (src/models/moonshine/encoder.py:31-40,60):
// depthwise cross-correlation (k=9, stride 2, VALID, depth multiplier 1, no bias) -> 1x1
// pointwise -> bias_add -> activation, fused in one kernel: the depthwise result never leaves
// shared memory.  This is the exact-arithmetic path (FP32 FMA, fixed summation order over the
// input channels); the tensor-core path is sepconv_tf32.cu.
//
// CTA tile: 64 output frames x 64 output channels, 256 threads with a 4x4 register tile each,
// input channels consumed in chunks of 16: stage x rows -> depthwise into Ys[c][t] ->
// rank-16 update from Ys and the pointwise chunk Ws[c][o].
#include "sepconv_common.cuh"

using namespace tasr;

namespace {

constexpr int TO = 64;    // output frames per CTA
constexpr int CT = 64;    // output channels per CTA
constexpr int KC = 16;    // input channels per chunk
constexpr int kThreads = 256;

__device__ __forceinline__ float apply_act(float z, int act) {
  switch (act) {
    case TASR_ACT_TANH: return tanhf(z);
    case TASR_ACT_GELU_ERF: return 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f));
    case TASR_ACT_RELU: return fmaxf(z, 0.0f);
    default: return z;
  }
}

template <int K, int STRIDE>
__global__ void __launch_bounds__(kThreads)
sepconv_fp32_kernel(const float* __restrict__ x, int T_in, int C_in, const float* __restrict__ dw,
                    const float* __restrict__ pw, const float* __restrict__ bias, int C_out, int act,
                    float* __restrict__ y, int T_out, int row_off) {
  constexpr int XR = (TO - 1) * STRIDE + K;   // input rows per tile (135)
  __shared__ __align__(16) float Xs[XR][KC + 1];
  __shared__ __align__(16) float Ys[KC][TO + 4];
  __shared__ __align__(16) float Ws[KC][CT];
  __shared__ float Dw[K][KC];

  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * TO;
  const int n0 = blockIdx.y * CT;
  const int tx = tid & 15, ty = tid >> 4;
  const float* xb = x + (size_t)b * T_in * C_in;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int c0 = 0; c0 < C_in; c0 += KC) {
    // stage x[2*t0 .. 2*t0+XR) x [c0, c0+16), depthwise taps and the pointwise chunk
    for (int i = tid; i < XR * KC; i += kThreads) {
      const int r = i / KC, c = i - r * KC;
      const int tg = t0 * STRIDE + r + row_off;     // row_off = -pad_left for padding='same': rows outside the tensor are zero
      Xs[r][c] = (tg >= 0 && tg < T_in && c0 + c < C_in) ? xb[(size_t)tg * C_in + c0 + c] : 0.0f;
    }
    for (int i = tid; i < K * KC; i += kThreads) {
      const int k = i / KC, c = i - k * KC;
      Dw[k][c] = (c0 + c < C_in) ? dw[k * C_in + c0 + c] : 0.0f;
    }
    for (int i = tid; i < KC * CT; i += kThreads) {
      const int c = i / CT, o = i - c * CT;
      Ws[c][o] = (c0 + c < C_in && n0 + o < C_out) ? pw[(size_t)(c0 + c) * C_out + n0 + o] : 0.0f;
    }
    __syncthreads();
    // depthwise: Ys[c][t] = sum_k Xs[STRIDE*t+k][c] * Dw[k][c], ascending k
    {
      const int c = tid & (KC - 1);
#pragma unroll
      for (int i = 0; i < (TO * KC) / kThreads; ++i) {
        const int t = (tid >> 4) + i * (kThreads / KC);
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) s = fmaf(Xs[STRIDE * t + k][c], Dw[k][c], s);
        Ys[c][t] = s;
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      const float4 a4 = *reinterpret_cast<const float4*>(&Ys[c][4 * ty]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Ws[c][4 * tx]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bv[j] = (n0 + 4 * tx + j < C_out) ? bias[n0 + 4 * tx + j] : 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + 4 * ty + i;
    if (t >= T_out) continue;
    float* yr = y + ((size_t)b * T_out + t) * C_out + n0 + 4 * tx;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = apply_act(acc[i][j] + bv[j], act);
    if (n0 + 4 * tx + 3 < C_out && (C_out & 3) == 0) {
      *reinterpret_cast<float4*>(yr) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + 4 * tx + j < C_out) yr[j] = o[j];
    }
  }
}

}  // namespace

namespace tasr {
int validate_sepconv(const char* who, const void* x, int32_t B, int32_t T_in, const TasrSepConvLayer* L,
                     const void* y, int32_t T_out) {
  if (!x || !L || !y) return fail(TASR_ERR_BAD_ARG, "%s: null argument", who);
  if (!L->dw || !L->pw || !L->bias) return fail(TASR_ERR_BAD_ARG, "%s: null weight pointer", who);
  if (B < 0 || T_in < 0 || T_out < 0 || L->c_in < 1 || L->c_out < 1) return fail(TASR_ERR_BAD_ARG, "%s: bad shape", who);
  if (L->kernel != 9 || L->stride != 2)
    return fail(TASR_ERR_UNSUPPORTED, "%s: kernels are built for kernel=9, stride=2 (config/model.yaml:24-25); got k=%d s=%d",
                who, L->kernel, L->stride);
  if (L->activation < TASR_ACT_NONE || L->activation > TASR_ACT_RELU) return fail(TASR_ERR_BAD_ARG, "%s: unknown activation %d", who, L->activation);
  const int32_t t_full = L->same ? (T_in + L->stride - 1) / L->stride : ((T_in >= L->kernel) ? (T_in - L->kernel) / L->stride + 1 : 0);
  if (T_out > t_full) return fail(TASR_ERR_BAD_ARG, "%s: t_out=%d exceeds the %s conv length %d of t_in=%d", who, T_out, L->same ? "same" : "valid", t_full, T_in);
  if (!aligned16(x) || !aligned16(y)) return fail(TASR_ERR_MISALIGNED, "%s: x/y must be 16-byte aligned", who);
  if (B > 65535) return fail(TASR_ERR_UNSUPPORTED, "%s: batch > 65535", who);
  return TASR_OK;
}
}  // namespace tasr

extern "C" int tasr_sepconv1d_f32(const float* x, int32_t B, int32_t T_in, const TasrSepConvLayer* L,
                                  float* y, int32_t T_out, tasr_stream_t stream) {
  int rc = validate_sepconv("tasr_sepconv1d_f32", x, B, T_in, L, y, T_out);
  if (rc != TASR_OK) return rc;
  if (B == 0 || T_out == 0) return TASR_OK;
  dim3 grid((T_out + TO - 1) / TO, (L->c_out + CT - 1) / CT, B);
  sepconv_fp32_kernel<9, 2><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
      x, T_in, L->c_in, L->dw, L->pw, L->bias, L->c_out, L->activation, y, T_out,
      L->same ? -tasr_sep::tasr_same_pad_left(T_in, L->kernel, L->stride) : 0);
  TASR_LAUNCH_CHECK("sepconv_fp32_kernel");
  return TASR_OK;
}
