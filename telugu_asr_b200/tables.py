"""Host-side float32 constant tables, built ONCE per featurizer instead of on every call
(the reference rebuilds both per call: src/speech_featurizer.py:96-101 via tf.signal.stft's
window_fn, and :114-120 linear_to_mel_weight_matrix).  Each numpy op below is one float32
rounding, in the op order of TF 2.15's window_ops._raised_cosine_window and
mel_ops.linear_to_mel_weight_matrix, so the tables match what the reference's ops produce.
The kernels take these as data and never recompute them on the device."""
from __future__ import annotations

import math

import numpy as np

_f32 = np.float32


def enclosing_power_of_two(n: int) -> int:
    """tf.signal.stft's default fft_length (spectral_ops._enclosing_power_of_two)."""
    return 1 << max(0, math.ceil(math.log2(n)))


def hann_window_f32(length: int, periodic: bool = True) -> np.ndarray:
    even = 1 - length % 2
    n = _f32(length + int(periodic) * even - 1)
    k = np.arange(length, dtype=np.float32)
    arg = _f32(2.0 * np.pi) * k / n
    return (_f32(0.5) - _f32(0.5) * np.cos(arg, dtype=np.float32)).astype(np.float32)


def _linspace_f32(start, stop, num: int) -> np.ndarray:
    start, stop = _f32(start), _f32(stop)
    if num == 1:
        return np.array([start], dtype=np.float32)
    delta = (stop - start) / _f32(num - 1)
    inner = start + delta * np.arange(1, num - 1, dtype=np.float32)
    return np.concatenate([[start], inner.astype(np.float32), [stop]]).astype(np.float32)


def _hz_to_mel_f32(f) -> np.ndarray:
    f = np.asarray(f, dtype=np.float32)
    return (_f32(1127.0) * np.log(_f32(1.0) + f / _f32(700.0), dtype=np.float32)).astype(np.float32)


def mel_weight_matrix_f32(num_mel_bins: int, num_spectrogram_bins: int, sample_rate: int,
                          lower_edge_hertz: float, upper_edge_hertz: float) -> np.ndarray:
    """[num_spectrogram_bins, num_mel_bins] HTK triangles, linear in mel, DC row zero."""
    nyquist = _f32(sample_rate) / _f32(2.0)
    bins_mel = _hz_to_mel_f32(_linspace_f32(0.0, nyquist, num_spectrogram_bins)[1:])[:, None]
    edges = _linspace_f32(_hz_to_mel_f32(_f32(lower_edge_hertz)), _hz_to_mel_f32(_f32(upper_edge_hertz)),
                          num_mel_bins + 2)
    lo, ce, hi = edges[None, :-2], edges[None, 1:-1], edges[None, 2:]
    up = (bins_mel - lo) / (ce - lo)
    down = (hi - bins_mel) / (hi - ce)
    w = np.maximum(_f32(0.0), np.minimum(up, down)).astype(np.float32)
    return np.ascontiguousarray(np.pad(w, [[1, 0], [0, 0]]).astype(np.float32))
