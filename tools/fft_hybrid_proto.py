"""Numerics prototype (CPU, numpy) of the round-2 tensor-core log-mel kernel (csrc/logmel_tc.cu).

512-point real DFT of a windowed frame y[0..399] (tail zero padded), n = n1 + 32 n2, k = 16 k1 + k2:

  stage 1 (FP32, CUDA cores, registers)   V[n1, k2] = sum_{n2<13} y[n1 + 32 n2] W16^(n2 k2)      k2 = 0..8
  stage 2 (tcgen05, FP32 accumulate)      X[16 k1 + k2] = sum_{n1<32} V~[n1, k2] W512^(n1 (16 k1 + k2))
                                          V~[n1, k2] = V[n1, k2] (k2 <= 8), conj V[n1, 16 - k2] (k2 > 8)

Stage 2 is a real GEMM per k2-pair block p = min(k2, 16-k2): A = [Re V[:,p], Im V[:,p]] (K = 64, or 32 for the real
blocks p = 0, 8), B = the twiddles with the conjugation folded in.  Both operands are split hi = fp16(a),
lo = fp16(a - hi); the kernel issues  [hi | lo] x [Bhi ; Bhi]  and  [hi | lo] x [Blo ; Blo]  (all four products).
The samples are scaled by `scale` up front (a power of two folded into the window) so that the low parts stay in
FP16's normal range.  Prints the max-abs log-mel error against the float64 oracle next to the float32 oracle's band.

    python tools/fft_hybrid_proto.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from oracle import featurizer_ref as fr

f32, f16 = np.float32, np.float16


def split16(a):
    a = a.astype(f32)
    hi = a.astype(f16).astype(f32)
    lo = (a - hi).astype(f16).astype(f32)
    return hi, lo


def stage2_matrices():
    """B[p] for p = 0..8: real [K, N] float64.  p in 1..7: K = 64 rows (n1 re, n1 im interleaved), N = 64 columns
    = dest k2=p (k1 re/im interleaved, 32) then dest k2=16-p (32).  p = 0, 8: K = 32, N = 32."""
    n1 = np.arange(32)
    k1 = np.arange(16)
    mats = {}
    for p in range(9):
        dests = [p] if p in (0, 8) else [p, 16 - p]
        cols = []
        for k2 in dests:
            W = np.exp(-2j * np.pi * np.outer(n1, 16 * k1 + k2) / 512.0)      # [n1, k1]
            conj_in = k2 > 8
            if p in (0, 8):      # V real: X = sum V * W
                blk = np.empty((32, 32))
                blk[:, 0::2] = W.real
                blk[:, 1::2] = W.imag
            else:                # V = a + i b (conj: a - i b): X = sum (a +- i b)(Wr + i Wi)
                s = -1.0 if conj_in else 1.0
                blk = np.empty((64, 32))
                blk[0::2, 0::2] = W.real            # a -> Re
                blk[0::2, 1::2] = W.imag            # a -> Im
                blk[1::2, 0::2] = -s * W.imag       # b -> Re
                blk[1::2, 1::2] = s * W.real        # b -> Im
            cols.append(blk)
        mats[p] = np.concatenate(cols, axis=1)
    return mats


MATS = stage2_matrices()


def power_tc(frames, scale, mode):
    """frames [T, 400] float32 windowed (pre-emphasised) samples -> |X[k]|^2 / scale^2 [T, 256] float32 (k = 0..255)."""
    T = frames.shape[0]
    y = np.zeros((T, 416), dtype=f32)
    y[:, :400] = frames * f32(scale)
    y = y.reshape(T, 13, 32)                                   # [T, n2, n1]
    V = np.fft.fft(y.astype(np.complex64), n=16, axis=1)       # [T, k2, n1]  (float32 FFT: the kernel's radix-16)
    V = V.astype(np.complex64)
    X = np.zeros((T, 256), dtype=np.complex64)
    for p in range(9):
        if p in (0, 8):
            A = V[:, p, :].real.astype(f32)                    # [T, 32]
        else:
            A = np.empty((T, 64), dtype=f32)
            A[:, 0::2] = V[:, p, :].real
            A[:, 1::2] = V[:, p, :].imag
        B = MATS[p].astype(f32)
        if mode == "fp32":
            D = A @ B
        elif mode == "tf32x1":
            def t32(a):
                return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(f32)
            D = t32(A.copy()) @ t32(B.copy())
        else:
            ah, al = split16(A)
            bh, bl = split16(MATS[p])
            if mode == "f16x1":
                D = ah @ bh
            elif mode == "f16x3":
                D = ah @ bh + al @ bh + ah @ bl
            else:
                D = (ah @ bh + al @ bh) + (ah @ bl + al @ bl)
        D = D.astype(f32)
        dests = [p] if p in (0, 8) else [p, 16 - p]
        for j, k2 in enumerate(dests):
            X[:, k2::16] = D[:, 32 * j: 32 * j + 32: 2] + 1j * D[:, 32 * j + 1: 32 * j + 32: 2]
    inv = f32(1.0 / (float(scale) ** 2))
    return ((X.real.astype(f32) ** 2 + X.imag.astype(f32) ** 2) * inv).astype(f32)


def logmel_hybrid(x, mode, scale=2048.0):
    x = x.astype(f32)
    g = f32(1.0) / (np.abs(x).max() + f32(1e-9))
    xn = (x * g).astype(f32)
    y = np.concatenate([xn[:1], xn[1:] - f32(0.97) * xn[:-1]]).astype(f32)
    T = 1 + (len(y) - 400) // 160
    idx = 160 * np.arange(T)[:, None] + np.arange(400)[None, :]
    frames = (y[idx] * fr.hann_periodic(400).astype(f32)).astype(f32)
    P = power_tc(frames, scale, mode)
    W = fr.htk_mel_matrix_f32()[:256]                          # bin 256 has zero weight
    M = (P @ W).astype(f32)
    return (np.log(np.maximum(M, f32(1e-9))) / np.log(f32(10.0))).astype(f32)


if __name__ == "__main__":
    # self-check of the decomposition in float64-ish precision
    rng = np.random.default_rng(0)
    fr_ = rng.standard_normal((3, 400)).astype(f32)
    ref = np.abs(np.fft.rfft(fr_.astype(np.float64), n=512, axis=1)[:, :256]) ** 2
    got = power_tc(fr_, 1.0, "fp32")
    assert np.allclose(got, ref, rtol=2e-4, atol=1e-3), np.abs(got - ref).max()
    for dist in ("tilt", "white", "tone_noise", "half_silence"):
        wav, ln = oracle.make_waveforms([48000, 16000], seed=3, dist=dist)
        worst = {m: 0.0 for m in ("fp32", "f16x4", "f16x3", "f16x1", "tf32x1")}
        band = 0.0
        for b in range(2):
            x = wav[b, : ln[b]]
            r64 = oracle.logmel_ref(x, dtype=np.float64)
            r32 = oracle.logmel_ref(x, dtype=np.float32)
            band = max(band, float(np.abs(r32 - r64).max()))
            for m in worst:
                worst[m] = max(worst[m], float(np.abs(logmel_hybrid(x, m) - r64).max()))
        print(f"{dist:13s} " + "  ".join(f"{m} {v:.2e}" for m, v in worst.items()) + f"   (float32 oracle band {band:.2e}; budget 1e-4)")


def _rz32(x64):
    """float64 -> float32 rounding toward zero (a pessimistic model of the tensor core's accumulator)."""
    r = x64.astype(f32)
    up = np.abs(r.astype(np.float64)) > np.abs(x64)
    r[up] = np.nextafter(r[up], f32(0.0))
    return r


def mma_chunked(A, B64, chunk=8, rz=False):
    """The kernel's schedule: per chunk of `chunk` K-values one MMA pair  [hi|lo] x [Bhi;Bhi]  and  [hi|lo] x [Blo;Blo];
    products exact, summed exactly inside an instruction (float64 here), accumulator rounded once per instruction."""
    ah, al = split16(A)
    bh, bl = split16(B64)
    acc = np.zeros((A.shape[0], B64.shape[1]), dtype=f32)
    rnd = _rz32 if rz else (lambda v: v.astype(f32))
    for k0 in range(0, A.shape[1], chunk):
        s = slice(k0, k0 + chunk)
        a2 = (ah[:, s] + al[:, s]).astype(np.float64)     # hi + lo is exact in float64
        acc = rnd(acc.astype(np.float64) + a2 @ bh[s].astype(np.float64))
        acc = rnd(acc.astype(np.float64) + a2 @ bl[s].astype(np.float64))
    return acc


if __name__ == "__main__" and "--chunked" in sys.argv:
    import types
    for rz in (False, True):
        def power_chunked(frames, scale, mode, rz=rz):
            T = frames.shape[0]
            y = np.zeros((T, 416), dtype=f32)
            y[:, :400] = frames * f32(scale)
            V = np.fft.fft(y.reshape(T, 13, 32).astype(np.complex64), n=16, axis=1).astype(np.complex64)
            X = np.zeros((T, 256), dtype=np.complex64)
            for p in range(9):
                if p in (0, 8):
                    A = V[:, p, :].real.astype(f32)
                else:
                    A = np.empty((T, 64), dtype=f32)
                    A[:, 0::2] = V[:, p, :].real
                    A[:, 1::2] = V[:, p, :].imag
                D = mma_chunked(A, MATS[p], chunk=8 if p not in (0, 8) else 4, rz=rz)
                dests = [p] if p in (0, 8) else [p, 16 - p]
                for j, k2 in enumerate(dests):
                    X[:, k2::16] = D[:, 32 * j: 32 * j + 32: 2] + 1j * D[:, 32 * j + 1: 32 * j + 32: 2]
            inv = f32(1.0 / (float(scale) ** 2))
            return ((X.real.astype(f32) ** 2 + X.imag.astype(f32) ** 2) * inv).astype(f32)
        power_tc = power_chunked
        for dist in ("tilt", "white", "tone_noise", "half_silence"):
            wav, ln = oracle.make_waveforms([48000, 16000, 32000, 24000], seed=3, dist=dist)
            worst = band = 0.0
            ratios = []
            for b in range(4):
                x = wav[b, : ln[b]]
                r64 = oracle.logmel_ref(x, dtype=np.float64)
                r32 = oracle.logmel_ref(x, dtype=np.float32)
                e = float(np.abs(logmel_hybrid(x, "chunked") - r64).max())
                bd = float(np.abs(r32 - r64).max())
                worst = max(worst, e); band = max(band, bd); ratios.append(e / bd)
            print(f"accumulate {'RZ' if rz else 'RN'}  {dist:13s} chunked f16x4 {worst:.2e}  band {band:.2e}  per-utt ratios {[round(r, 2) for r in ratios]}")
