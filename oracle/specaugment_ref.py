"""CPU restatement of the reference's SpecAugment (TEST INFRASTRUCTURE — only tests/, smoke() and
bench.py's CPU legs may import oracle/).

Follows src/augmentations/specaugment.py:6-62 (FreqMasking.augment, TimeMasking.augment) and
src/augmentations/augmentation.py:19-35 (Augmentation._augment: each augmentation is applied when a
uniform draw is below `prob`).  The reference draws with TensorFlow's RNG, which cannot be reproduced
offline, so parity is split in two: the deterministic part — the mask arithmetic for GIVEN (t0,t,f0,f)
— is checked bit for bit, and the draws are checked for their distribution/bounds.
"""
from __future__ import annotations

import numpy as np


def freq_mask_ref(spectrogram: np.ndarray, f0: int, f: int) -> np.ndarray:
    """specaugment.py:20-31: spectrogram [T,F,V] * concat(ones[T,f0,V], zeros[T,f,V], ones[T,F-f-f0,V])."""
    T, F, V = spectrogram.shape
    mask = np.concatenate([np.ones((T, f0, V), spectrogram.dtype), np.zeros((T, f, V), spectrogram.dtype),
                           np.ones((T, F - f - f0, V), spectrogram.dtype)], axis=1)
    return spectrogram * mask


def time_mask_ref(spectrogram: np.ndarray, t0: int, t: int) -> np.ndarray:
    """specaugment.py:44-61: spectrogram [T,F,V] * concat(ones[t0,F,V], zeros[t,F,V], ones[T-t0-t,F,V])."""
    T, F, V = spectrogram.shape
    mask = np.concatenate([np.ones((t0, F, V), spectrogram.dtype), np.zeros((t, F, V), spectrogram.dtype),
                           np.ones((T - t0 - t, F, V), spectrogram.dtype)], axis=0)
    return spectrogram * mask


def draw_freq_mask(rng: np.random.Generator, F: int, mask_factor: int = 27):
    """specaugment.py:17-19: f ~ U{0..mask_factor-1}, f = min(f, F), f0 ~ U{0..F-f-1}."""
    f = min(int(rng.integers(0, mask_factor)), F)
    f0 = int(rng.integers(0, F - f)) if F - f > 0 else 0
    return f0, f


def draw_time_mask(rng: np.random.Generator, T: int, mask_factor: float = 100, p_upperbound: float = 1.0):
    """specaugment.py:45-50: t ~ U{0..mask_factor-1}, t = min(t, int(float32(T)*p_upperbound)), t0 ~ U{0..T-t-1}."""
    t = int(rng.integers(0, int(mask_factor)))
    t = min(t, int(np.float32(T) * np.float32(p_upperbound)))
    t0 = int(rng.integers(0, T - t)) if T - t > 0 else 0
    return t0, t


def apply_batch_ref(feat: np.ndarray, n_frames, time_masks, freq_masks) -> np.ndarray:
    """feat [B,T_max,F,1] zero padded; per utterance the un-padded [T_b,F,1] block goes through
    time_mask_ref / freq_mask_ref for every (t0,t) / (f0,f) of its row (width 0 = not applied)."""
    out = feat.copy()
    for b in range(feat.shape[0]):
        T = int(n_frames[b])
        x = out[b, :T]
        for t0, t in np.asarray(time_masks[b]).reshape(-1, 2):
            if t > 0:
                x = time_mask_ref(x, int(t0), int(t))
        for f0, f in np.asarray(freq_masks[b]).reshape(-1, 2):
            if f > 0:
                x = freq_mask_ref(x, int(f0), int(f))
        out[b, :T] = x
    return out
