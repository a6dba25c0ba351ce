"""CPU restatement of the reference featurizer.  TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED by the reference (no tests / vectors; TensorFlow absent) — see
oracle/__init__.py for what pins it instead.

Follows, function by function:
  src/speech_featurizer.py:19-66    constructor arithmetic (frame_length/frame_step)
  src/speech_featurizer.py:68-72    normalize_signal
  src/speech_featurizer.py:74-79    preemphasis_signal
  src/speech_featurizer.py:95-105   stft  (tf.signal.stft -> abs -> square)
  src/speech_featurizer.py:107-122  logarithm / log_mel_spectrogram
  src/speech_featurizer.py:81-93    normalize_audio_feature
  src/speech_featurizer.py:124-133  spectrogram / mfcc / waveform
  src/speech_featurizer.py:163-166  get_nframes
  src/utils/math_util.py:17-18      log10
  src/dataset.py:171-175,236-252    per-utterance call + zero padded_batch
and the TF 2.15 ops those call (hann_window, frame, rfft, linear_to_mel_weight_matrix).

Every function takes ``dtype``: np.float32 reproduces the reference's float32
op sequence op for op (each TF op is one rounding); np.float64 evaluates the same
formulas with the *same float32 tables* (window, mel weights) upcast, and is the
"exact" value the float32 implementations are compared against.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import math

import numpy as np

__all__ = [
    "FeatParams",
    "hann_periodic",
    "htk_mel_matrix_f32",
    "get_nframes",
    "logmel_ref",
    "logmel_batch_ref",
    "collate_ref",
    "featurize_ref",
]

FEATURE_TYPES = ("waveform", "spectrogram", "log_mel_spectrogram", "mfcc")


@dataclass
class FeatParams:
    """Mirror of SpeechFeaturizer.__init__ kwargs (src/speech_featurizer.py:19-39);
    defaults are the constructor defaults, `from_yaml_dict` takes config/model.yaml:1-17."""

    sample_rate: int = 16000
    frame_ms: int = 25
    stride_ms: int = 10
    num_feature_bins: int = 80
    feature_type: str = "log_mel_spectrogram"
    preemphasis: float = 0.97
    pad_end: bool = False
    lower_edge_hertz: float = 0.0
    upper_edge_hertz: float = 8000.0
    output_floor: float = 1e-9
    log_base: str = "10"
    nfft: int | None = 512
    normalize_signal: bool = False
    normalize_zscore: bool = False
    normalize_min_max: bool = False
    padding: float = 0.0
    augmentation_config: dict = field(default_factory=dict)

    def __post_init__(self):
        # src/speech_featurizer.py:40,59
        assert self.feature_type in FEATURE_TYPES, f"Unsupported feature type: {self.feature_type}"
        self.log_base = str(self.log_base)
        assert self.log_base in ("10", "e"), "log_base must be '10' or 'e'"

    # src/speech_featurizer.py:46,49
    @property
    def frame_length(self) -> int:
        return int(round(self.sample_rate * self.frame_ms / 1000.0))

    @property
    def frame_step(self) -> int:
        return int(round(self.sample_rate * self.stride_ms / 1000.0))

    @property
    def fft_length(self) -> int:
        # tf.signal.stft without fft_length: enclosing power of two of frame_length
        # (spectral_ops.py `_enclosing_power_of_two`); `nfft` is stored but unused
        # (src/speech_featurizer.py:65 vs :96-101).
        return 1 << max(0, math.ceil(math.log2(self.frame_length)))

    @property
    def num_spectrogram_bins(self) -> int:
        return self.fft_length // 2 + 1

    @classmethod
    def from_yaml_dict(cls, d: dict) -> "FeatParams":
        return cls(**dict(d))


# config/model.yaml:1-17 (== config/conformer.yaml:1-17)
def yaml_params() -> FeatParams:
    return FeatParams(
        sample_rate=16000, frame_ms=25, stride_ms=10, num_feature_bins=80,
        feature_type="log_mel_spectrogram", preemphasis=0.97, pad_end=False,
        lower_edge_hertz=0.0, upper_edge_hertz=8000.0, output_floor=1e-9, log_base="10",
        nfft=512, normalize_signal=True, normalize_zscore=False, normalize_min_max=False,
        padding=0.0,
    )


def hann_periodic(window_length: int) -> np.ndarray:
    """tf.signal.hann_window(window_length, periodic=True, dtype=float32)
    (window_ops._raised_cosine_window): n = L + periodic*even - 1;
    w = 0.5 - 0.5*cos(2*pi*count/n), every op in float32."""
    even = 1 - window_length % 2
    n = np.float32(window_length + 1 * even - 1)
    count = np.arange(window_length, dtype=np.float32)
    cos_arg = np.float32(2.0 * np.pi) * count / n
    w = np.float32(0.5) - np.float32(0.5) * np.cos(cos_arg, dtype=np.float32)
    return w.astype(np.float32)


def _tf_linspace_f32(start: np.float32, stop: np.float32, num: int) -> np.ndarray:
    """tf.linspace in float32 (math_ops.linspace_nd): delta=(stop-start)/(num-1);
    interior points start + delta*i, endpoints exact."""
    start = np.float32(start)
    stop = np.float32(stop)
    if num == 1:
        return np.array([start], dtype=np.float32)
    delta = (stop - start) / np.float32(num - 1)
    i = np.arange(1, num - 1, dtype=np.float32)
    mid = start + delta * i
    return np.concatenate([[start], mid.astype(np.float32), [stop]]).astype(np.float32)


def _hertz_to_mel_f32(f: np.ndarray) -> np.ndarray:
    # mel_ops._hertz_to_mel: 1127.0 * ln(1 + f/700.0), float32
    f = np.asarray(f, dtype=np.float32)
    return (np.float32(1127.0) * np.log(np.float32(1.0) + f / np.float32(700.0), dtype=np.float32)).astype(np.float32)


def htk_mel_matrix_f32(num_mel_bins: int = 80, num_spectrogram_bins: int = 257, sample_rate: int = 16000,
                       lower_edge_hertz: float = 0.0, upper_edge_hertz: float = 8000.0) -> np.ndarray:
    """tf.signal.linear_to_mel_weight_matrix (mel_ops.py) restated op for op in
    float32; the reference rebuilds it on every call (src/speech_featurizer.py:114-120).
    Returns [num_spectrogram_bins, num_mel_bins] float32, DC row zero, not area-normalised."""
    sr = np.float32(sample_rate)
    nyquist = sr / np.float32(2.0)
    linear_freqs = _tf_linspace_f32(np.float32(0.0), nyquist, num_spectrogram_bins)[1:]
    spec_mel = _hertz_to_mel_f32(linear_freqs)[:, None]                      # [bins-1, 1]
    edges = _tf_linspace_f32(_hertz_to_mel_f32(np.float32(lower_edge_hertz)),
                             _hertz_to_mel_f32(np.float32(upper_edge_hertz)), num_mel_bins + 2)
    lower = edges[:-2][None, :]
    center = edges[1:-1][None, :]
    upper = edges[2:][None, :]
    lower_slopes = (spec_mel - lower) / (center - lower)
    upper_slopes = (upper - spec_mel) / (upper - center)
    w = np.maximum(np.float32(0.0), np.minimum(lower_slopes, upper_slopes)).astype(np.float32)
    return np.pad(w, [[1, 0], [0, 0]]).astype(np.float32)


def get_nframes(nsamples: int, p: FeatParams | None = None, *, clamp: bool = True) -> int:
    """src/speech_featurizer.py:163-166.  tf.signal.frame yields max(0, .) frames; the
    reference's bare formula goes negative for N < frame_length, `clamp=False` returns that."""
    p = p or yaml_params()
    if p.pad_end:
        return -(-nsamples // p.frame_step)
    n = 1 + (nsamples - p.frame_length) // p.frame_step
    return max(0, n) if clamp else n


def _frames(y: np.ndarray, p: FeatParams) -> np.ndarray:
    """tf.signal.frame(y, frame_length, frame_step, pad_end) on a 1-D signal."""
    L, S = p.frame_length, p.frame_step
    n = y.shape[0]
    if p.pad_end:
        T = -(-n // S)
        need = (T - 1) * S + L if T > 0 else 0
        if need > n:
            y = np.concatenate([y, np.zeros(need - n, dtype=y.dtype)])
    else:
        T = max(0, 1 + (n - L) // S)
    if T == 0:
        return np.zeros((0, L), dtype=y.dtype)
    idx = (np.arange(T)[:, None] * S) + np.arange(L)[None, :]
    return y[idx]


def _stft_power(y: np.ndarray, p: FeatParams, dtype) -> np.ndarray:
    """src/speech_featurizer.py:95-105: frame -> * periodic Hann -> rfft zero-padded at the
    TAIL to fft_length -> |.| -> square."""
    fr = _frames(y, p)
    w = hann_periodic(p.frame_length).astype(dtype)
    fr = (fr * w).astype(dtype)
    nfft = p.fft_length
    if fr.shape[0] == 0:
        return np.zeros((0, nfft // 2 + 1), dtype=dtype)
    if dtype == np.float32:
        # A float32 rfft op (TF: Eigen TensorFFT on CPU, cuFFT on GPU) runs its butterflies in float32.  numpy's
        # np.fft.rfft does NOT: on float32 input it returns the float64 transform rounded once to complex64 (numpy 2.3:
        # bit-identical, relative rms error 2.5e-8 = pure output rounding), which made the float32 "band" of round 1
        # ~4x tighter than any genuine float32 FFT.  scipy.fft (pocketfft instantiated for float) computes in single
        # precision: 1.0e-7 relative rms, the same as torch.fft.rfft / MKL (tests/test_oracle_featurizer.py pins both).
        import scipy.fft
        X = scipy.fft.rfft(np.ascontiguousarray(fr, dtype=np.float32), n=nfft, axis=-1)
        assert X.dtype == np.complex64
        mag = np.abs(X).astype(np.float32)       # tf.abs(complex64) -> float32
        return np.square(mag).astype(np.float32)  # tf.square
    X = np.fft.rfft(fr, n=nfft, axis=-1)
    mag = np.abs(X)
    return np.square(mag)


def _logarithm(S: np.ndarray, p: FeatParams, dtype) -> np.ndarray:
    """src/speech_featurizer.py:107-110 + src/utils/math_util.py:17-18."""
    S = np.maximum(S, dtype(p.output_floor))
    if p.log_base == "10":
        return (np.log(S, dtype=dtype) / np.log(dtype(10.0), dtype=dtype)).astype(dtype)
    return np.log(S, dtype=dtype).astype(dtype)


def _normalize_audio_feature(feat: np.ndarray, p: FeatParams, dtype) -> np.ndarray:
    """src/speech_featurizer.py:81-93; axis=1 is the mel axis of the [T, F] per-utterance input."""
    if p.normalize_zscore:
        mean = feat.mean(axis=1, keepdims=True, dtype=dtype)
        var = np.mean(np.square(feat - mean), axis=1, keepdims=True, dtype=dtype)
        return ((feat - mean) / np.sqrt(var + dtype(1e-9))).astype(dtype)
    if p.normalize_min_max:
        if p.feature_type == "spectrogram":
            mn = _logarithm(np.array(p.output_floor, dtype=dtype), p, dtype)
        else:
            mn = feat.min(axis=1, keepdims=True)
        return ((feat - mn) / (feat.max(axis=1, keepdims=True) - mn)).astype(dtype)
    return feat


def _dct2_ortho_mfcc(log_mel: np.ndarray, dtype) -> np.ndarray:
    """tf.signal.mfccs_from_log_mel_spectrograms: DCT-II (unnormalised, scale 2) * rsqrt(2*M)."""
    M = log_mel.shape[-1]
    n = np.arange(M, dtype=np.float64)
    k = np.arange(M, dtype=np.float64)[:, None]
    basis = 2.0 * np.cos(np.pi * (2.0 * n + 1.0) * k / (2.0 * M))           # [k, n]
    out = log_mel.astype(np.float64) @ basis.T / np.sqrt(2.0 * M)
    return out.astype(dtype)


def featurize_ref(x: np.ndarray, p: FeatParams | None = None, dtype=np.float32,
                  mel_w: np.ndarray | None = None) -> np.ndarray:
    """SpeechFeaturizer.call(x_1d, training=False) (src/speech_featurizer.py:136-161).
    x: 1-D waveform.  Returns [T, F] (or the waveform itself for feature_type='waveform')."""
    p = p or yaml_params()
    dtype = np.dtype(dtype).type
    x = np.asarray(x)
    assert x.ndim == 1, "the reference featurizer is called per utterance on a 1-D signal (src/dataset.py:171)"
    x = x.astype(np.float32).astype(dtype)     # audio is float32 at the boundary
    if p.normalize_signal:                     # :68-72
        gain = dtype(1.0) / (np.max(np.abs(x)) + dtype(1e-9)) if x.size else dtype(1.0)
        x = (x * gain).astype(dtype)
    if p.preemphasis and p.preemphasis > 0.0:  # :74-79
        c = dtype(p.preemphasis)
        y = np.empty_like(x)
        if x.size:
            y[0] = x[0]
            y[1:] = x[1:] - (c * x[:-1]).astype(dtype)
        x = y
    if p.feature_type == "waveform":
        return x
    S = _stft_power(x, p, dtype)
    if p.feature_type == "spectrogram":        # :124-126
        feat = _logarithm(S, p, dtype)[:, : p.num_feature_bins]
    else:
        if mel_w is None:
            mel_w = htk_mel_matrix_f32(p.num_feature_bins, S.shape[-1], p.sample_rate,
                                       p.lower_edge_hertz, p.upper_edge_hertz)
        M = (S @ mel_w.astype(dtype)).astype(dtype)   # tf.matmul :121
        feat = _logarithm(M, p, dtype)
        if p.feature_type == "mfcc":           # :128-130
            feat = _dct2_ortho_mfcc(feat, dtype)
    return _normalize_audio_feature(feat, p, dtype)


def logmel_ref(x: np.ndarray, p: FeatParams | None = None, dtype=np.float32) -> np.ndarray:
    """One utterance, exactly as src/dataset.py:171 calls the featurizer.  [N] -> [T, 80]."""
    p = p or yaml_params()
    assert p.feature_type == "log_mel_spectrogram"
    return featurize_ref(x, p, dtype)


def collate_ref(feats: list[np.ndarray]) -> tuple[np.ndarray, np.ndarray]:
    """expand_dims(-1) (src/dataset.py:173), length = T (:175), padded_batch with 0.0
    (:236-252).  list of [T_i, F] -> ([B, T_max, F, 1], n_frames[B] int32)."""
    B = len(feats)
    F = feats[0].shape[1] if B else 0
    T_max = max((f.shape[0] for f in feats), default=0)
    out = np.zeros((B, T_max, F, 1), dtype=feats[0].dtype if B else np.float32)
    n = np.zeros((B,), dtype=np.int32)
    for b, f in enumerate(feats):
        out[b, : f.shape[0], :, 0] = f
        n[b] = f.shape[0]
    return out, n


def logmel_batch_ref(wav: np.ndarray, lengths: np.ndarray, p: FeatParams | None = None,
                     dtype=np.float32) -> tuple[np.ndarray, np.ndarray]:
    """The reference pipeline on a padded [B, N_max] + lengths[B] batch: per-utterance
    featurise (each on its own un-padded samples), then zero-pad collate."""
    p = p or yaml_params()
    feats = [featurize_ref(np.asarray(wav[b, : int(lengths[b])]), p, dtype) for b in range(wav.shape[0])]
    return collate_ref(feats)
