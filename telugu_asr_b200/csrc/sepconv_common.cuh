// Shared pieces of the TF32 separable-convolution kernels (sepconv_tf32.cu: one CTA per tile;
// sepconv_ws.cu: persistent, warp-specialised): the plan, PTX wrappers for mbarrier / TMA bulk copy /
// tcgen05, UMMA descriptors, the SFU activations and the argument block.
#pragma once
#include "common.cuh"

struct TasrSepConvPlan {
  TasrSepConvLayer L;   // dw / bias pointers are borrowed from the caller (must outlive the plan)
  int device;
  int n_split;          // output channels are processed in n_split slices of NT
  int NT;
  int n_chunks;         // ceil(c_in / 32)
  float* d_bpack;       // [n_split][n_chunks][NT*32] shared-memory images of pw^T
  void* kernel;         // template instance for (c_in, activation)
  float* d_pad_in;      // [9][c_in]  nine copies of the input's padding row (ragged mode)
  float* d_pad_out;     // [c_out]    this layer's output for an all-padding receptive field
  int pad_ready;
  int use_ws;           // 1: persistent warp-specialised kernel (sepconv_ws.cu) when the shape allows it
  int ws_roles;         // role counts of the persistent kernel: 28 (two depthwise groups, 8 epilogue warps; default), 18, 116;
                        // TASR_WS_ROLES at plan creation (development aid, all bit-identical)
};

namespace tasr_sep {

constexpr int kThreads = 256;
constexpr int kMT = 128;                 // output frames per tile (UMMA M)
constexpr int kKC = 32;                  // input channels per chunk (one 128-byte swizzle row)
constexpr int kRun = kMT / (kThreads / 32);   // 16 output frames per warp
constexpr int kWin = 2 * (kRun - 1) + 9; // 39 input rows per run
constexpr int kABytes = kMT * kKC * 4;   // 16 KiB per stage
constexpr int kTmemCols = 256;
constexpr int kStgStride = 36;           // floats; conflict-free for 128-bit row writes and reads

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, TF32 inputs, FP32 accumulate.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 B
// apart (stride byte offset), descriptor version 1 (sm_100).  `saddr` must be 1024-byte aligned
// (+ 32*k bytes to step along K inside the swizzle row).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
  d |= (uint64_t)1 << 46;                 // version
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// Instruction descriptor: D=F32, A=B=TF32, both K-major, N, M.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// tanh(z) = 1 - 2/(1 + e^{2z}) on the SFU (ex2 + rcp): absolute error < 3e-7 over the whole range,
// saturates correctly (e -> inf gives 1, e -> 0 gives -1).
__device__ __forceinline__ float tanh_fast(float z) {
  const float e = ex2_approx(z * 2.88539008177792681472f);   // 2*log2(e)
  return fmaf(-2.0f, rcp_approx(1.0f + e), 1.0f);
}
// Exact-erf GELU (Keras approximate=False): 0.5 z (1 + erf(z/sqrt2)) = max(z,0) - |0.5 z erfc(|z|/sqrt2)|, with
// erfc(u/sqrt2) = 2^q(u) and q a degree-5 weighted-minimax fit of log2(erfc(u/sqrt2)) on [0,8] (tools/fit_gelu.py; the
// weight is the sensitivity of GELU to q, the leading coefficient is negative so 2^q -> 0 beyond the interval).
// |error| <= 8.5e-7 over the whole range evaluated in float32 (the float32 ulp at z = 4 is 4.8e-7): one MUFU and ten
// instructions per value instead of the seventeen of the Abramowitz-Stegun 7.1.26 form used before.
__device__ __forceinline__ float gelu_erf_fast(float z) {
  const float a = fabsf(z);
  float q = fmaf(-4.732913936e-04f, a, 7.084427742e-03f);
  q = fmaf(q, a, -5.182704213e-02f);
  q = fmaf(q, a, -4.599928375e-01f);
  q = fmaf(q, a, -1.150787652e+00f);
  q = fmaf(q, a, -3.765495672e-05f);
  const float t = (0.5f * z) * ex2_approx(q);
  return fmaxf(z, 0.0f) - fabsf(t);
}

// ---- packed FP32x2 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2): one issue slot for two IEEE-identical operations.
// The epilogues are bound by instruction issue, not by the FMA pipe, so the bias add and the GELU polynomial of two
// neighbouring columns share their instructions; every result bit equals the scalar form above.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// Depthwise run of 16 output frames from a 39-row register window (stride 2, 9 taps): outputs j and j + 8 share their
// nine FMAs as packed FP32x2 instructions (taps ascending, like the scalar loop: identical bits).
__device__ __forceinline__ void depthwise_run16(const float (&v)[39], const float (&w)[9], float (&acc)[16]) {
  u64 wp[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) wp[k] = pack2(w[k], w[k]);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    u64 a2 = pack2(0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < 9; ++k) a2 = fma2(pack2(v[2 * j + k], v[2 * j + 16 + k]), wp[k], a2);
    unpack2(a2, acc[j], acc[j + 8]);
  }
}

// bias add + activation of two neighbouring accumulator columns: (acc0 + b0, acc1 + b1) -> act.
template <int ACT>
__device__ __forceinline__ void act_apply2(float acc0, float acc1, float b0, float b1, float& o0, float& o1);

__device__ __forceinline__ void gelu_erf_fast2(u64 z2, float& o0, float& o1) {
  float z0, z1;
  unpack2(z2, z0, z1);
  const u64 a2 = pack2(fabsf(z0), fabsf(z1));
  u64 q = fma2(pack2(-4.732913936e-04f, -4.732913936e-04f), a2, pack2(7.084427742e-03f, 7.084427742e-03f));
  q = fma2(q, a2, pack2(-5.182704213e-02f, -5.182704213e-02f));
  q = fma2(q, a2, pack2(-4.599928375e-01f, -4.599928375e-01f));
  q = fma2(q, a2, pack2(-1.150787652e+00f, -1.150787652e+00f));
  q = fma2(q, a2, pack2(-3.765495672e-05f, -3.765495672e-05f));
  float q0, q1;
  unpack2(q, q0, q1);
  const u64 t2 = mul2(mul2(pack2(0.5f, 0.5f), z2), pack2(ex2_approx(q0), ex2_approx(q1)));
  float t0, t1;
  unpack2(t2, t0, t1);
  o0 = fmaxf(z0, 0.0f) - fabsf(t0);
  o1 = fmaxf(z1, 0.0f) - fabsf(t1);
}

// tanh of two columns: the scale, the 1 + e and the final fma are packed; same operations as tanh_fast (identical bits)
__device__ __forceinline__ void tanh_fast2(u64 z2, float& o0, float& o1) {
  float s0, s1;
  unpack2(mul2(z2, pack2(2.88539008177792681472f, 2.88539008177792681472f)), s0, s1);
  float d0, d1;
  unpack2(add2(pack2(1.0f, 1.0f), pack2(ex2_approx(s0), ex2_approx(s1))), d0, d1);
  unpack2(fma2(pack2(-2.0f, -2.0f), pack2(rcp_approx(d0), rcp_approx(d1)), pack2(1.0f, 1.0f)), o0, o1);
}

template <int ACT>
__device__ __forceinline__ float act_apply(float z);

template <int ACT>
__device__ __forceinline__ void act_apply2(float acc0, float acc1, float b0, float b1, float& o0, float& o1) {
  const u64 z2 = add2(pack2(acc0, acc1), pack2(b0, b1));
  if (ACT == TASR_ACT_GELU_ERF) {
    gelu_erf_fast2(z2, o0, o1);
  } else if (ACT == TASR_ACT_TANH) {
    tanh_fast2(z2, o0, o1);
  } else {
    float z0, z1;
    unpack2(z2, z0, z1);
    o0 = act_apply<ACT>(z0);
    o1 = act_apply<ACT>(z1);
  }
}

// bias add + activation of sixteen accumulator columns held as eight FP32x2 pairs, IN PLACE and STAGE BY STAGE: every stage is
// applied to all eight pairs before the next one starts, so the sixteen MUFU / FMA dependency chains are in flight together.
// (Written value by value — act_apply2 in a loop — ptxas reuses the same four registers for every group of four columns and the
// chains run one after another: LDS -> FADD2 -> FMUL2 -> MUFU -> FADD2 -> MUFU -> FFMA2 -> STS, ~150 cycles each, 8 per block of 32
// columns; that serial chain, not issue slots or the MUFU pipe, was what an epilogue warp spent its time on.)  Same operations in
// the same order per value as act_apply2: identical bits.
template <int ACT>
__device__ __forceinline__ void act_block16(u64 (&z)[8]) {
  if (ACT == TASR_ACT_TANH) {
    float e[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) unpack2(mul2(z[i], pack2(2.88539008177792681472f, 2.88539008177792681472f)), e[2 * i], e[2 * i + 1]);
#pragma unroll
    for (int j = 0; j < 16; ++j) e[j] = ex2_approx(e[j]);
#pragma unroll
    for (int i = 0; i < 8; ++i) unpack2(add2(pack2(1.0f, 1.0f), pack2(e[2 * i], e[2 * i + 1])), e[2 * i], e[2 * i + 1]);
#pragma unroll
    for (int j = 0; j < 16; ++j) e[j] = rcp_approx(e[j]);
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = fma2(pack2(-2.0f, -2.0f), pack2(e[2 * i], e[2 * i + 1]), pack2(1.0f, 1.0f));
  } else if (ACT == TASR_ACT_GELU_ERF) {
    u64 a2[8], q[8];
    float x[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      unpack2(z[i], x[2 * i], x[2 * i + 1]);
      a2[i] = pack2(fabsf(x[2 * i]), fabsf(x[2 * i + 1]));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = fma2(pack2(-4.732913936e-04f, -4.732913936e-04f), a2[i], pack2(7.084427742e-03f, 7.084427742e-03f));
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = fma2(q[i], a2[i], pack2(-5.182704213e-02f, -5.182704213e-02f));
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = fma2(q[i], a2[i], pack2(-4.599928375e-01f, -4.599928375e-01f));
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = fma2(q[i], a2[i], pack2(-1.150787652e+00f, -1.150787652e+00f));
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = fma2(q[i], a2[i], pack2(-3.765495672e-05f, -3.765495672e-05f));
    float e[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) unpack2(q[i], e[2 * i], e[2 * i + 1]);
#pragma unroll
    for (int j = 0; j < 16; ++j) e[j] = ex2_approx(e[j]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t0, t1;
      unpack2(mul2(mul2(pack2(0.5f, 0.5f), z[i]), pack2(e[2 * i], e[2 * i + 1])), t0, t1);
      z[i] = pack2(fmaxf(x[2 * i], 0.0f) - fabsf(t0), fmaxf(x[2 * i + 1], 0.0f) - fabsf(t1));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float z0, z1;
      unpack2(z[i], z0, z1);
      z[i] = pack2(act_apply<ACT>(z0), act_apply<ACT>(z1));
    }
  }
}

template <int ACT>
__device__ __forceinline__ float act_apply(float z) {
  if (ACT == TASR_ACT_TANH) return tanh_fast(z);
  if (ACT == TASR_ACT_GELU_ERF) return gelu_erf_fast(z);
  if (ACT == TASR_ACT_RELU) return fmaxf(z, 0.0f);
  return z;
}

struct SepArgs {
  const float* x;
  const float* dw;
  const float* bpack;
  const float* bias;
  float* y;
  int32_t T_in, T_out, C_in, C_out, NT, n_chunks, act;
  // ragged mode (len0 != nullptr): rows t >= ceil(len0[b] / 2^shift) of x[b] all equal the padding row the
  // plan was given, so a tile whose receptive field starts there is the constant row pad_out: filled, not computed.
  const int32_t* len0;
  const float* pad_out;
  int32_t shift;
  // lean mode (fill_rows >= 0): only the first fill_rows padding rows of y[b] — rows ceil(len0[b] / 2^(shift+1)) + i,
  // i < fill_rows, rounded up to whole tiles — are guaranteed to be written; tiles beyond are left untouched
  // (for an intermediate tensor whose only reader is the next ragged layer).  < 0: every row is written.
  int32_t fill_rows;
  // deferred input gain (in_peak != nullptr; first layer after tasr_logmel_f32_single_pass): a row t < cf of x is read as
  // max(x + in_scale2 * lg2(1 / (in_peak[b] + 1e-9)), in_floor)
  const float* in_peak;
  float in_scale2, in_floor;
  // padding = "same" (dense mode only): output frame t reads input rows 2t + k + row_off, row_off = -pad_left of TensorFlow's
  // SAME rule for the tensor length T_in; rows outside [0, T_in) are zero.  0 for "valid".
  int32_t row_off;
};

// TensorFlow SAME padding before the first row: out = ceil(T/s), total = max((out-1) s + k - T, 0), left = total / 2.
inline int tasr_same_pad_left(int T, int k, int s) {
  const int out = (T + s - 1) / s;
  const int total = (out - 1) * s + k - T;
  return total > 0 ? total / 2 : 0;
}

}  // namespace tasr_sep

// sepconv_ws.cu: launches the persistent warp-specialised kernel; returns a negative value (and launches
// nothing) when the shape is outside its limits, so that the caller falls back to the per-tile kernel.
int tasr_sepconv_ws_launch(const TasrSepConvPlan* p, const tasr_sep::SepArgs& sa, int32_t B, cudaStream_t st);
