"""B200-native drop-in for the reference SpeechFeaturizer (src/speech_featurizer.py:17-192).

Same constructor keywords (the `speech_config` block of config/model.yaml:1-17 goes in
verbatim, as src/helpers/dataset_helpers.py:68 does), same derived attributes, same
`__call__(inputs, training=False)`, `get_nframes`, `compute_output_shape`, `get_config`.

Differences, all additive:
  * inputs are CUDA tensors; the per-utterance 1-D call of src/dataset.py:171 returns [T, F];
  * `featurizer(waveforms[B, N_max], lengths[B])` is the batched form that also does the
    zero-padded collate of src/dataset.py:236-252 on the device and returns
    ([B, T_max, F, 1], n_frames[B]);
  * the Hann window and the mel matrix are built once (telugu_asr_b200/tables.py), not per call.

All arithmetic runs in libtasr_b200.so (csrc/absmax.cu, csrc/logmel.cu); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import asdict, dataclass

import torch

from . import _native, tables

__all__ = ["SpeechFeaturizer", "FeaturizerConfig"]


@dataclass
class FeaturizerConfig:  # src/speech_featurizer.py:10-15
    waveform: str = "waveform"
    spectrogram: str = "spectrogram"
    log_mel_spectrogram: str = "log_mel_spectrogram"
    mfcc: str = "mfcc"


class DeferredGain:
    """What `featurize_batch(single_pass=True)` hands to the reader of its raw features: the device tensor of
    per-utterance peaks and the host struct (TasrDeferredGain) the C entry points take."""

    def __init__(self, peak: torch.Tensor):
        self.peak = peak
        self.struct = _native.TasrDeferredGain(peak=peak.data_ptr(), log_scale_x2=0.0, log_floor=0.0)


class SpeechFeaturizer:
    def __init__(
        self,
        sample_rate: int = 16000,
        frame_ms: int = 25,
        stride_ms: int = 10,
        num_feature_bins: int = 80,
        feature_type: str = "log_mel_spectrogram",
        preemphasis: float = 0.97,
        pad_end: bool = False,
        lower_edge_hertz: float = 0.0,
        upper_edge_hertz: float = 8000.0,
        output_floor: float = 1e-9,
        log_base: str = "10",
        nfft: int = 512,
        normalize_signal: bool = False,
        normalize_zscore: bool = False,
        normalize_min_max: bool = False,
        padding: float = 0.0,
        augmentation_config: dict | None = None,
        **kwargs,
    ):
        # src/speech_featurizer.py:40,59 — construction-time assertions, same messages
        assert feature_type in asdict(FeaturizerConfig()).values(), (
            f"Unsupported feature type: {feature_type}. Supported types: {asdict(FeaturizerConfig()).values()}")
        self.name = kwargs.pop("name", feature_type)
        self.sample_rate = sample_rate
        self.frame_ms = frame_ms
        self.frame_length = int(round(self.sample_rate * self.frame_ms / 1000.0))   # :46
        self.stride_ms = stride_ms
        self.frame_step = int(round(self.sample_rate * self.stride_ms / 1000.0))    # :49
        self.num_feature_bins = num_feature_bins
        self.feature_type = feature_type
        self.preemphasis = preemphasis
        self.pad_end = pad_end
        self.lower_edge_hertz = lower_edge_hertz
        self.upper_edge_hertz = upper_edge_hertz
        self.output_floor = output_floor
        self.log_base = str(log_base)
        assert self.log_base in ("10", "e"), "log_base must be '10' or 'e'"
        self._normalize_signal = normalize_signal
        self._normalize_zscore = normalize_zscore
        self._normalize_min_max = normalize_min_max
        self.padding = padding
        self.nfft = self.frame_length if nfft is None else nfft   # stored, unused — like :65
        self.augmentation_config = augmentation_config or {}
        # tf.signal.stft is called without fft_length (:96-101): enclosing power of two.
        self.fft_length = tables.enclosing_power_of_two(self.frame_length)
        self.dtype = torch.float32
        self._handles: dict[int, int] = {}      # device index -> TasrFeaturizer*
        self._hann = None
        self._mel_w = None
        self.profile_events: list | None = None   # set to [] to collect (start, end) events per logmel launch

    # ------------------------------------------------------------------ tables / handle
    def _tables(self):
        if self._hann is None:
            self._hann = tables.hann_window_f32(self.frame_length, periodic=True)
            self._mel_w = tables.mel_weight_matrix_f32(
                self.num_feature_bins, self.fft_length // 2 + 1, self.sample_rate,
                self.lower_edge_hertz, self.upper_edge_hertz)
        return self._hann, self._mel_w

    def _check_supported(self):
        if self.padding and self.padding > 0:
            raise NotImplementedError("padding > 0 (0.0 in config/model.yaml:17; the reference branch fails on 1-D input)")
        # (frame geometries other than config/model.yaml's 400 / 160 / 512 / 80 run on the general kernel of
        #  csrc/logmel_generic.cu: log-mel and spectrogram, no mfcc / per-frame normalisation, no single-pass mode)

    def _handle(self, device: torch.device) -> int:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is not None:
            return h
        self._check_supported()
        hann, mel_w = self._tables()
        p = _native.TasrFeatParams(
            sample_rate=self.sample_rate, frame_length=self.frame_length, frame_step=self.frame_step,
            fft_length=self.fft_length, num_mel_bins=self.num_feature_bins,
            normalize_signal=int(bool(self._normalize_signal)), log_base_e=int(self.log_base == "e"),
            pad_end=int(bool(self.pad_end)),
            preemphasis=float(self.preemphasis) if self.preemphasis else 0.0,
            output_floor=float(self.output_floor), feature_type=_native.FEATURE_TYPES[self.feature_type],
            normalize_zscore=int(bool(self._normalize_zscore)), normalize_min_max=int(bool(self._normalize_min_max)))
        out = C.c_void_p()
        with torch.cuda.device(idx):
            _native.check(_native.lib().tasr_featurizer_create(
                C.byref(p), hann.ctypes.data_as(C.c_void_p), mel_w.ctypes.data_as(C.c_void_p), C.byref(out)))
        self._handles[idx] = out.value
        return out.value

    def __del__(self):
        try:
            for h in self._handles.values():
                _native.lib().tasr_featurizer_destroy(h)
        except Exception:
            pass

    # ------------------------------------------------------------------ reference API
    def get_nframes(self, nsamples):  # src/speech_featurizer.py:163-166
        if self.pad_end:
            return -(-nsamples // self.frame_step)
        return 1 + (nsamples - self.frame_length) // self.frame_step

    def compute_output_shape(self, input_shape):  # :168-178
        B, nsamples = input_shape
        if nsamples is None:
            return (B, None, self.num_feature_bins, 1)
        if self.feature_type == FeaturizerConfig.waveform:
            return (B, None, 1)
        return (B, int(self.get_nframes(nsamples + self.padding)), self.num_feature_bins, 1)

    def get_config(self):  # :180-190
        return {
            "name": self.name,
            "sample_rate": self.sample_rate,
            "feature_type": self.feature_type,
            "normalize_signal": self._normalize_signal,
            "preemphasis": self.preemphasis,
            "padding": self.padding,
            "augmentation_config": self.augmentation_config,
        }

    def __call__(self, inputs, lengths=None, training: bool = False, out=None):
        """1-D `inputs` [N] -> [T, F]   (reference semantics, src/dataset.py:171).
        2-D `inputs` [B, N_max] with `lengths` [B] (int32, CUDA) -> ([B, T_max, F, 1], n_frames[B]);
        `lengths=None` means every row is N_max samples long."""
        if training:
            # src/speech_featurizer.py:158-159 dereferences self.augmentation, which :66 removed.
            raise AttributeError("'SpeechFeaturizer' object has no attribute 'augmentation' "
                                 "(the reference raises here too; call with training=False)")
        x = _native.require_cuda(inputs, "inputs")
        if x.dtype != torch.float32:
            raise ValueError(f"inputs must be float32 (the reference decodes audio to float32); got {x.dtype}")
        if self.feature_type == FeaturizerConfig.waveform:
            return self._waveform(x, lengths)
        if x.dim() == 1:
            if lengths is not None:
                raise ValueError("lengths is only meaningful for batched [B, N_max] inputs")
            feats, _ = self.featurize_batch(x.unsqueeze(0), None)
            return feats[0, :, :, 0]
        if x.dim() != 2:
            raise ValueError(f"inputs must be [N] or [B, N_max]; got shape {tuple(x.shape)}")
        return self.featurize_batch(x, lengths, out=out)

    call = __call__

    def _waveform(self, x: torch.Tensor, lengths):
        """feature_type 'waveform' (src/speech_featurizer.py:132-133): the normalised, pre-emphasised signal.
        [N] -> [N]; [B, N_max] (+ lengths) -> [B, N_max] with zeros beyond each length."""
        one = x.dim() == 1
        wav = (x.unsqueeze(0) if one else x).contiguous()
        if wav.dim() != 2:
            raise ValueError(f"inputs must be [N] or [B, N_max]; got shape {tuple(x.shape)}")
        B, n_max = wav.shape
        n_pad = -(-max(n_max, 1) // 4) * 4
        buf = torch.zeros((B, n_pad), dtype=torch.float32, device=wav.device)
        buf[:, :n_max] = wav
        if lengths is None:
            lengths = torch.full((B,), n_max, dtype=torch.int32, device=wav.device)
        lengths = _native.require_cuda(lengths, "lengths").to(torch.int32).contiguous()
        out = torch.zeros_like(buf)
        if B and n_max:
            h = self._handle(wav.device)
            L = _native.lib()
            with torch.cuda.device(wav.device):
                st = _native.stream_ptr()
                peak_ptr = None
                if self._normalize_signal:
                    peak = torch.empty((B,), dtype=torch.float32, device=wav.device)
                    _native.check(L.tasr_absmax_f32(buf.data_ptr(), lengths.data_ptr(), B, n_pad, peak.data_ptr(), st))
                    peak_ptr = peak.data_ptr()
                _native.check(L.tasr_waveform_f32(h, buf.data_ptr(), lengths.data_ptr(), peak_ptr, B, n_pad, out.data_ptr(), st))
        out = out[:, :n_max]
        return out[0] if one else out

    def is_reference_geometry(self) -> bool:
        return (self.frame_length, self.frame_step, self.fft_length, self.num_feature_bins) == (400, 160, 512, 80)

    def supports_single_pass(self) -> bool:
        return bool(self.is_reference_geometry() and self._normalize_signal and not self.pad_end and not self._normalize_zscore and not self._normalize_min_max
                    and self.feature_type in (FeaturizerConfig.log_mel_spectrogram, FeaturizerConfig.spectrogram))

    def apply_deferred_gain(self, raw: torch.Tensor, n_frames: torch.Tensor, gain: "DeferredGain") -> torch.Tensor:
        """In place: raw [B, T, F(,1)] from `featurize_batch(single_pass=True)` -> the reference's feature values
        max(raw + 2 log(1/(peak+1e-9)), log(output_floor)) on rows t < n_frames[b]."""
        _native.require_cuda(raw, "raw")
        B, T, F = raw.shape[0], raw.shape[1], raw.shape[2]
        with torch.cuda.device(raw.device):
            _native.check(_native.lib().tasr_apply_deferred_gain(raw.data_ptr(), n_frames.data_ptr(), B, T, F,
                                                                 C.byref(gain.struct), _native.stream_ptr()))
        return raw

    # ------------------------------------------------------------------ batched device path
    def featurize_batch(self, wav: torch.Tensor, lengths: torch.Tensor | None, out: torch.Tensor | None = None,
                        t_max: int | None = None, pad_fill_rows: int | None = None, single_pass: bool = False):
        """wav [B, N_max] float32 CUDA (rows zero padded; padding is never read), lengths [B] int32
        CUDA -> (features [B, T_max, F, 1] with rows >= n_frames[b] equal to 0.0, n_frames [B] int32).
        T_max defaults to get_nframes(N_max) clamped at 0, i.e. the collate's batch maximum.
        `pad_fill_rows` (lean mode, for a feature tensor that only feeds the ragged subsampling): write just that
        many 0.0 rows past each utterance's frames and leave the rest of the collate padding untouched.
        `single_pass` (see `supports_single_pass`): the waveform is read once — the kernel finds max|x| itself and
        featurises the un-normalised signal; returns (raw_features, n_frames, DeferredGain), and the reader
        (Conv1DSubsamplingLayer(input_gain=...) or `apply_deferred_gain`) adds the gain and the floor."""
        if single_pass and not self.supports_single_pass():
            raise ValueError("single_pass needs normalize_signal=True, pad_end=False, a log-mel or spectrogram featurizer "
                             "and no per-frame normalisation")
        _native.require_cuda(wav, "wav")
        dev = wav.device
        B, n_max = wav.shape
        if wav.stride(1) != 1 and n_max > 1:
            wav = wav.contiguous()
        row_stride = wav.stride(0) if B > 1 else max(n_max, 0)   # one row: the stride only has to cover the row
        if (row_stride % 4) or (wav.data_ptr() % 16):   # (a single row of n_max % 4 != 0 samples is padded too: the kernels
            # read whole 16-byte vectors up to the row's end)
            # keep rows 16-byte aligned for the 128-bit loads (costs one copy; pad N_max to a
            # multiple of 4 upstream to avoid it)
            n_pad = -(-n_max // 4) * 4
            buf = torch.zeros((B, n_pad), dtype=torch.float32, device=dev)
            buf[:, :n_max] = wav
            wav, row_stride = buf, n_pad
        if lengths is None:
            lengths = torch.full((B,), n_max, dtype=torch.int32, device=dev)
        else:
            _native.require_cuda(lengths, "lengths")
            if lengths.dtype != torch.int32:
                lengths = lengths.to(torch.int32)
            lengths = lengths.contiguous()
            if lengths.numel() != B:
                raise ValueError(f"lengths has {lengths.numel()} entries for a batch of {B}")
        if t_max is None:
            t_max = max(0, self.get_nframes(n_max)) if n_max >= 0 else 0
        F = self.num_feature_bins
        if out is None:
            out = _native.empty((B, t_max, F, 1), torch.float32, dev)
        else:
            if tuple(out.shape) != (B, t_max, F, 1) or not out.is_contiguous() or out.dtype != torch.float32:
                raise ValueError(f"out must be a contiguous float32 [B={B}, T_max={t_max}, {F}, 1] tensor")
        n_frames = torch.empty((B,), dtype=torch.int32, device=dev)
        gain = None
        if single_pass:
            gain = DeferredGain(_native.empty((B,), torch.float32, dev))
        if B == 0:
            return (out, n_frames, gain) if single_pass else (out, n_frames)
        if t_max == 0:   # nothing to featurise: every utterance is shorter than one frame
            if single_pass:
                gain.peak.zero_()
            return (out, n_frames.zero_(), gain) if single_pass else (out, n_frames.zero_())
        h = self._handle(dev)
        L = _native.lib()
        with torch.cuda.device(dev):
            st = _native.stream_ptr()
            peak_ptr = None
            _native.mark("begin")
            if single_pass:
                ev = self.profile_events
                if ev is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                _native.check(L.tasr_logmel_f32_single_pass(
                    h, wav.data_ptr(), lengths.data_ptr(), B, row_stride, out.data_ptr(), t_max, n_frames.data_ptr(),
                    -1 if pad_fill_rows is None else int(pad_fill_rows), gain.peak.data_ptr(), C.byref(gain.struct), st))
                if ev is not None:
                    e1.record()
                    ev.append((e0, e1))
                _native.mark("logmel_kernel")
                return out, n_frames, gain
            if self._normalize_signal:
                peak = torch.empty((B,), dtype=torch.float32, device=dev)
                _native.check(L.tasr_absmax_f32(wav.data_ptr(), lengths.data_ptr(), B, row_stride, peak.data_ptr(), st))
                peak_ptr = peak.data_ptr()
                _native.mark("absmax_kernel")
            ev = self.profile_events
            if ev is not None:   # bench.py: CUDA events around the dominant kernel, on its own stream
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            lean = (pad_fill_rows is not None and not self._normalize_zscore and not self._normalize_min_max
                    and self.feature_type in (FeaturizerConfig.log_mel_spectrogram, FeaturizerConfig.spectrogram))
            if lean:
                _native.check(L.tasr_logmel_f32_lean(h, wav.data_ptr(), lengths.data_ptr(), peak_ptr, B, row_stride,
                                                     out.data_ptr(), t_max, n_frames.data_ptr(), int(pad_fill_rows), st))
            else:
                _native.check(L.tasr_logmel_f32(h, wav.data_ptr(), lengths.data_ptr(), peak_ptr, B, row_stride,
                                                out.data_ptr(), t_max, n_frames.data_ptr(), st))
            if ev is not None:
                e1.record()
                ev.append((e0, e1))
            _native.mark("logmel_kernel")
        return out, n_frames
