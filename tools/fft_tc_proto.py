"""Numerics prototype (CPU, numpy) for the round-2 log-mel kernel: the 512-point real FFT of a windowed frame as TWO
tensor-core GEMM stages with split-precision FP16 operands and FP32 accumulation.

  n = 16*n1 + n2 (n1 < 32, n2 < 16),  k = k1 + 32*k2 (k1 < 32, k2 < 16):
  stage 1   Y[k1, n2] = sum_n1 y[16 n1 + n2] * W32^(n1 k1)          real [.,32] x complex [32,17] GEMM (k1 = 0..16 suffice)
  twiddle   Y[k1, n2] *= W512^(n2 k1)                               FP32, CUDA cores, out of TMEM
  stage 2   X[k1 + 32 k2] = sum_n2 Y[k1, n2] * W16^(n2 k2)          complex [.,16] x complex [16,16] GEMM
  (bins k1 = 17..31 follow from conjugate symmetry of the real input: X[512 - k] = conj X[k])

Every GEMM operand is split a = hi + lo with hi = fp16(a), lo = fp16(a - hi); a product uses hi*hi + hi*lo + lo*hi
(three MMAs, FP32 accumulate).  The script emulates exactly that (FP16-rounded operands multiplied and summed in
float32) on the bench distribution and on the stress distributions, runs power -> mel -> log10 in float32 like the
kernel, and prints the max-abs log-mel error against the float64 oracle next to the float32-oracle band.
Development aid:  python tools/fft_tc_proto.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from oracle import featurizer_ref as fr

f32, f16 = np.float32, np.float16


def split(a):
    a = a.astype(f32)
    hi = a.astype(f16).astype(f32)
    lo = (a - hi).astype(f16).astype(f32)
    return hi, lo


def mm3(a, b):
    """a @ b with both operands split hi/lo in FP16, three products, float32 accumulation."""
    ah, al = split(a)
    bh, bl = split(b)
    return (ah @ bh + (ah @ bl + al @ bh)).astype(f32)


def mm3s(a, b, k=f32(2048.0)):
    """The same with the low parts stored scaled by 2^11 (FP16's subnormal range starts at 6e-5: the unscaled low part
    of a small sample loses bits); the two cross products are rescaled in FP32 after the MMA."""
    a = a.astype(f32); b = b.astype(f32)
    ah = a.astype(f16).astype(f32); al = ((a - ah) * k).astype(f16).astype(f32)
    bh = b.astype(f16).astype(f32); bl = ((b - bh) * k).astype(f16).astype(f32)
    return (ah @ bh + (ah @ bl + al @ bh) / k).astype(f32)


def mm1(a, b):   # single FP16 pass, for comparison
    return (a.astype(f16).astype(f32) @ b.astype(f16).astype(f32)).astype(f32)


n1 = np.arange(32)[:, None]; k1 = np.arange(17)[None, :]
W32 = np.exp(-2j * np.pi * n1 * k1 / 32)                       # [32, 17]
n2 = np.arange(16)[:, None]; k2 = np.arange(16)[None, :]
W16 = np.exp(-2j * np.pi * n2 * k2 / 16)                       # [16, 16]
TW = np.exp(-2j * np.pi * np.arange(17)[:, None] * np.arange(16)[None, :] / 512)   # [k1, n2]


def rfft512_tc(frames, mm):
    """frames [T, 512] float32 (windowed, zero padded) -> X [T, 257] complex64 via the two GEMM stages."""
    T = frames.shape[0]
    y = frames.reshape(T, 32, 16)                              # [T, n1, n2]
    a = y.transpose(0, 2, 1).reshape(T * 16, 32)               # rows (t, n2), K = n1
    Yr, Yi = mm(a, W32.real.astype(f32)), mm(a, W32.imag.astype(f32))          # [T*16, 17]
    Y = (Yr + 1j * Yi).reshape(T, 16, 17).transpose(0, 2, 1)   # [T, k1, n2]
    Y = (Y * TW[None]).astype(np.complex64)                    # FP32 twiddle on the CUDA cores
    b = Y.reshape(T * 17, 16)                                  # rows (t, k1), K = n2
    br, bi = b.real.astype(f32), b.imag.astype(f32)
    Wr, Wi = W16.real.astype(f32), W16.imag.astype(f32)
    Xr = mm(br, Wr) - mm(bi, Wi)
    Xi = mm(br, Wi) + mm(bi, Wr)
    X = (Xr + 1j * Xi).reshape(T, 17, 16)                      # [T, k1, k2] -> bin k1 + 32 k2
    out = np.zeros((T, 257), dtype=np.complex64)
    for kk1 in range(17):
        for kk2 in range(16):
            k = kk1 + 32 * kk2
            if k <= 256:
                out[:, k] = X[:, kk1, kk2]
            if 512 - k <= 256 and k != 0:
                out[:, 512 - k] = np.conj(X[:, kk1, kk2])
    return out


def logmel_tc(x, mm):
    x = x.astype(f32)
    g = f32(1.0) / (np.abs(x).max() + f32(1e-9))
    xn = (x * g).astype(f32)
    y = np.concatenate([xn[:1], xn[1:] - f32(0.97) * xn[:-1]]).astype(f32)
    T = 1 + (len(y) - 400) // 160
    idx = 160 * np.arange(T)[:, None] + np.arange(400)[None, :]
    fr_ = np.zeros((T, 512), dtype=f32)
    fr_[:, :400] = y[idx] * fr.hann_periodic(400).astype(f32)
    X = rfft512_tc(fr_, mm)
    P = (X.real.astype(f32) ** 2 + X.imag.astype(f32) ** 2).astype(f32)
    M = (P @ fr.htk_mel_matrix_f32()).astype(f32)
    return (np.log(np.maximum(M, f32(1e-9))) / np.log(f32(10.0))).astype(f32)


if __name__ == "__main__":
    for dist in ("tilt", "white", "tone_noise", "half_silence"):
        wav, ln = oracle.make_waveforms([48000, 16000], seed=3, dist=dist)
        worst3 = worst3s = worst1 = band = 0.0
        for b in range(2):
            x = wav[b, : ln[b]]
            r64 = oracle.logmel_ref(x, dtype=np.float64)
            r32 = oracle.logmel_ref(x, dtype=np.float32)
            worst3 = max(worst3, float(np.abs(logmel_tc(x, mm3) - r64).max()))
            worst3s = max(worst3s, float(np.abs(logmel_tc(x, mm3s) - r64).max()))
            worst1 = max(worst1, float(np.abs(logmel_tc(x, mm1) - r64).max()))
            band = max(band, float(np.abs(r32 - r64).max()))
        print(f"{dist:13s} max-abs log-mel error vs float64: 3-product FP16 split {worst3:.2e}   with scaled low parts {worst3s:.2e}   single FP16 pass {worst1:.2e}   "
              f"(float32 oracle band {band:.2e}; budget 1e-4)")
