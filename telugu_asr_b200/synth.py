"""Seeded synthetic 16 kHz utterances.  numpy only, no oracle, no torch.

No audio ships with the reference (its TSVs point at /home/hemanth/..., config/model.yaml:62),
so every test and benchmark runs on these.  Distributions:
  tilt         AR(1) x[n] = 0.97 x[n-1] + e[n], e ~ N(0,1): speech-like spectral tilt that
               pre-emphasis (src/speech_featurizer.py:74-79) whitens.  Primary.
  white        N(0, 0.1^2)
  tone_noise   440 Hz + 3.1 kHz tones + N(0, 0.01^2)
  half_silence tilt in the first half, exact zeros after (exercises the 1e-9 log floor)
  zeros        all-zero utterance
Every utterance is scaled to peak 0.5 and quantised to k/32768, i.e. the float32 values
tf.audio.decode_wav gives for int16 PCM (src/utils/data_util.py:31).
"""
from __future__ import annotations

import numpy as np

__all__ = ["make_waveforms", "draw_lengths", "DISTRIBUTIONS", "to_pcm16"]

DISTRIBUTIONS = ("tilt", "white", "tone_noise", "half_silence", "zeros")


def _ar1(e: np.ndarray, rho: float) -> np.ndarray:
    try:
        from scipy.signal import lfilter
        return lfilter([1.0], [1.0, -rho], e, axis=-1)
    except Exception:  # pragma: no cover - scipy is in the image
        x = np.empty_like(e)
        acc = np.zeros(e.shape[:-1], dtype=e.dtype)
        for n in range(e.shape[-1]):
            acc = rho * acc + e[..., n]
            x[..., n] = acc
        return x


def _quantise(x: np.ndarray, peak: float = 0.5) -> np.ndarray:
    m = np.max(np.abs(x), axis=-1, keepdims=True)
    m = np.where(m > 0, m, 1.0)
    q = np.round(x / m * (peak * 32768.0))
    return (q / 32768.0).astype(np.float32)


def _one(dist: str, n: int, rng: np.random.Generator, sample_rate: int) -> np.ndarray:
    if n == 0:
        return np.zeros(0, dtype=np.float32)
    if dist == "tilt":
        return _quantise(_ar1(rng.standard_normal(n), 0.97))
    if dist == "white":
        return _quantise(rng.standard_normal(n) * 0.1)
    if dist == "tone_noise":
        t = np.arange(n) / float(sample_rate)
        x = np.sin(2 * np.pi * 440.0 * t) + 0.3 * np.sin(2 * np.pi * 3100.0 * t + 0.5)
        return _quantise(x + 0.01 * rng.standard_normal(n))
    if dist == "half_silence":
        x = _ar1(rng.standard_normal(n), 0.97)
        x[n // 2:] = 0.0
        return _quantise(x)
    if dist == "zeros":
        return np.zeros(n, dtype=np.float32)
    raise ValueError(f"unknown distribution {dist!r}; choose from {DISTRIBUTIONS}")


def make_waveforms(lengths, seed: int = 0, dist: str = "tilt", sample_rate: int = 16000,
                   n_max: int | None = None, align: int = 4):
    """Returns (wav [B, N_max] float32 zero-padded, lengths [B] int32).  N_max is rounded up
    to a multiple of `align` samples so every row starts 16-byte aligned."""
    lengths = np.asarray(lengths, dtype=np.int32).reshape(-1)
    B = lengths.shape[0]
    nm = int(lengths.max()) if B else 0
    if n_max is not None:
        nm = max(nm, int(n_max))
    nm = -(-nm // align) * align
    wav = np.zeros((B, nm), dtype=np.float32)
    for b in range(B):
        rng = np.random.default_rng([seed, b])
        wav[b, : lengths[b]] = _one(dist, int(lengths[b]), rng, sample_rate)
    return wav, lengths


def draw_lengths(B: int, lo: int, hi: int, seed: int, first_is_max: bool = True) -> np.ndarray:
    """N_b ~ U{lo..hi} with N_0 = hi (config 3 of BASELINE.md)."""
    rng = np.random.default_rng([seed, 12345])
    n = rng.integers(lo, hi + 1, size=B).astype(np.int32)
    if first_is_max and B:
        n[0] = hi
    return n


def to_pcm16(wav: np.ndarray) -> np.ndarray:
    """Exact inverse of the k/32768 quantisation: the int16 PCM these utterances decode from."""
    return np.round(wav * 32768.0).astype(np.int16)
