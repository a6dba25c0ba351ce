"""Numerics prototype (CPU, numpy) of the tensor-core log-mel kernel csrc/logmel_tc.cu.

512-point real DFT of a windowed frame u[0..399] (tail zero padded), n = n1 + 32 n2, k = 16 k1 + p:

  stage 1 (FP32, CUDA cores)   V[n1, p]  = sum_{n2<13} u[n1 + 32 n2] W16^(n2 p)            p = 0..8   (real-input FFT-16)
                               Vt[n1, p] = V[n1, p] * W512^(n1 p)                           (twiddle, FP32)
  stage 2 (tcgen05 kind::f16)  F_p[k1']  = sum_{n1<32} Vt[n1, p] W32^(n1 k1')               k1' = 0..31 (one shared DFT-32
                               matrix for every p; rows of the GEMM are (frame, p) pairs, K = 64 = (n1, re/im), N = 64)
  bins                         X[16 k1' + p] = F_p[k1']  (k1' < 16),   X[512 - 16 k1' - p] = conj F_p[k1']  (k1' >= 16)

Both GEMM operands are split hi = fp16(round11(a)), lo = fp16(a - hi); the kernel issues Ahi*Bhi + Alo*Bhi + Ahi*Blo with
FP32 accumulation.  The samples of a 32-frame tile are scaled by a power of two chosen from the tile's max |x| so that
fp16 neither overflows nor goes subnormal; the power is scaled back exactly after the mel projection.

    python tools/fft_tc_proto.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from oracle import featurizer_ref as fr

f32, f16 = np.float32, np.float16
TILE = 32


def split16(a):
    a = np.asarray(a, dtype=f32)
    hi = ((a.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(f32)   # rounded to 11 bits: exact in fp16 when normal
    lo = (a - hi).astype(f16).astype(f32)
    return hi.astype(f16).astype(f32), lo                                # (values below 2^-14 round on the fp16 subnormal grid)


def dft32_matrix():
    """B [K=64, N=64] float64: rows (n1, re/im), columns (k1', re/im):  (a + ib)(c - is) = (ac + bs) + i(bc - as)."""
    n1 = np.arange(32)[:, None]
    k1 = np.arange(32)[None, :]
    th = 2 * np.pi * n1 * k1 / 32.0
    B = np.empty((64, 64))
    B[0::2, 0::2] = np.cos(th)
    B[1::2, 0::2] = np.sin(th)
    B[0::2, 1::2] = -np.sin(th)
    B[1::2, 1::2] = np.cos(th)
    B[np.abs(B) < 1e-12] = 0.0
    return B


B64 = dft32_matrix()
BH, BL = split16(B64)
TW = np.exp(-2j * np.pi * np.outer(np.arange(32), np.arange(9)) / 512.0).astype(np.complex64)   # [n1, p]


def tile_scale(xmax):
    """power of two s with s * xmax in (1024, 2048] (1 for a silent tile)."""
    if not np.isfinite(xmax) or xmax <= 0:
        return f32(1.0)
    e = int(np.ceil(np.log2(float(xmax))))
    return f32(2.0 ** int(np.clip(11 - e, -100, 100)))


def power_tc(frames, mode):
    """frames [T, 400] float32 windowed, pre-emphasised, tile-scaled samples -> |X[k]|^2 [T, 257] float32 (k = 0 and 256 zero)."""
    T = frames.shape[0]
    y = np.zeros((T, 416), dtype=f32)
    y[:, :400] = frames
    y = y.reshape(T, 13, 32)                                             # [T, n2, n1]
    V = np.fft.fft(y.astype(np.complex64), n=16, axis=1)[:, :9, :]       # [T, p, n1]
    V = V.astype(np.complex64)
    Vt = (V * TW.T[None]).astype(np.complex64)                           # FP32 twiddle
    A = np.empty((T, 9, 64), dtype=f32)
    A[:, :, 0::2] = Vt.real
    A[:, :, 1::2] = Vt.imag
    A = A.reshape(T * 9, 64)
    if mode == "fp32":
        D = (A.astype(np.float64) @ B64).astype(f32)
    else:
        ah, al = split16(A)
        if mode == "f16x1":
            D = ah @ BH
        elif mode == "f16x3":
            D = (ah.astype(np.float64) @ BH + al.astype(np.float64) @ BH + ah.astype(np.float64) @ BL).astype(f32)
        elif mode == "f16x3_rz":       # accumulator rounded toward zero after every K=16 instruction (pessimistic)
            D = np.zeros((T * 9, 64), dtype=f32)
            for k0 in range(0, 64, 16):
                s = slice(k0, k0 + 16)
                for (a_, b_) in ((ah, BH), (al, BH), (ah, BL)):
                    D = _rz32(D.astype(np.float64) + a_[:, s].astype(np.float64) @ b_[s].astype(np.float64))
    D = D.reshape(T, 9, 32, 2)
    F = D[..., 0] + 1j * D[..., 1]                                       # [T, p, k1']
    X = np.zeros((T, 257), dtype=np.complex64)
    for p in range(9):
        for k1 in range(32):
            k = 16 * k1 + p if k1 < 16 else 512 - 16 * k1 - p
            if p == 8 and k1 >= 16:
                continue
            if p == 0 and k1 > 16:
                continue
            X[:, k] = F[:, p, k1]
    P = (X.real.astype(f32) ** 2 + X.imag.astype(f32) ** 2).astype(f32)
    return P


def _rz32(x64):
    r = x64.astype(f32)
    up = np.abs(r.astype(np.float64)) > np.abs(x64)
    r[up] = np.nextafter(r[up], f32(0.0))
    return r


def logmel_tc(x, mode, normalize=True):
    x = x.astype(f32)
    g = f32(1.0) / (np.abs(x).max() + f32(1e-9)) if normalize else f32(1.0)
    xn = (x * g).astype(f32)
    y = np.concatenate([xn[:1], xn[1:] - f32(0.97) * xn[:-1]]).astype(f32)
    T = 1 + (len(y) - 400) // 160
    W = fr.htk_mel_matrix_f32()
    win = fr.hann_periodic(400).astype(f32)
    out = np.empty((T, 80), dtype=f32)
    for f0 in range(0, T, TILE):
        nv = min(TILE, T - f0)
        s0, cnt = f0 * 160, (nv - 1) * 160 + 400
        s = tile_scale(np.abs(xn[s0: s0 + cnt]).max())
        ys = (y[s0: s0 + cnt] * s).astype(f32)
        idx = 160 * np.arange(nv)[:, None] + np.arange(400)[None, :]
        frames = (ys[idx] * win).astype(f32)
        P = power_tc(frames, mode)
        M = (P @ W).astype(f32)
        M = (M * (f32(1.0) / s) * (f32(1.0) / s)).astype(f32)
        out[f0: f0 + nv] = (np.log(np.maximum(M, f32(1e-9))) / np.log(f32(10.0))).astype(f32)
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    fr_ = rng.standard_normal((3, 400)).astype(f32)
    ref = np.abs(np.fft.rfft(fr_.astype(np.float64), n=512, axis=1)) ** 2
    got = power_tc(fr_, "fp32")
    assert np.allclose(got[:, 1:256], ref[:, 1:256], rtol=2e-4, atol=1e-3), np.abs(got - ref).max()
    modes = ("fp32", "f16x3", "f16x3_rz", "f16x1")
    for dist in ("tilt", "white", "tone_noise", "half_silence"):
        wav, ln = oracle.make_waveforms([48000, 16000, 32000], seed=3, dist=dist)
        worst = {m: 0.0 for m in modes}
        band = 0.0
        for b in range(len(ln)):
            x = wav[b, : ln[b]]
            r64 = oracle.logmel_ref(x, dtype=np.float64)
            r32 = oracle.logmel_ref(x, dtype=np.float32)
            band = max(band, float(np.abs(r32 - r64).max()))
            for m in modes:
                worst[m] = max(worst[m], float(np.abs(logmel_tc(x, m) - r64).max()))
        print(f"{dist:13s} " + "  ".join(f"{m} {v:.2e}" for m, v in worst.items()) + f"   (float32 oracle band {band:.2e}; budget 1e-4)")
