"""The whole hot path as one object: waveforms -> log-mel -> 3x separable conv -> encoder input.

Mirrors what the reference does across three places — the per-utterance featurizer call and
padded_batch (src/dataset.py:171-175, 236-252), the mask derivation (model.py:80) and
Conv1DSubsamplingLayer.call (encoder.py:50-71) — with every stage on the device."""
from __future__ import annotations


import torch
import yaml

from .speech_featurizer import SpeechFeaturizer
from .subsampling import Conv1DSubsamplingLayer

__all__ = ["FrontEnd", "ConformerFrontEnd", "CapturedFrontEnd", "InterleavedFrontEnd", "REFERENCE_SPEECH_CONFIG", "REFERENCE_SUBSAMPLING_CONFIG", "load_reference_yaml"]

# config/model.yaml:1-17
REFERENCE_SPEECH_CONFIG = dict(
    sample_rate=16000, frame_ms=25, stride_ms=10, num_feature_bins=80,
    feature_type="log_mel_spectrogram", preemphasis=0.97, pad_end=False, lower_edge_hertz=0.0,
    upper_edge_hertz=8000.0, output_floor=1e-9, log_base="10", nfft=512, normalize_signal=True,
    normalize_zscore=False, normalize_min_max=False, padding=0.0)
# config/model.yaml:21-27 (note the key is `activation`, which the layer does not read)
REFERENCE_SUBSAMPLING_CONFIG = dict(
    name="conv1d", kernel_size=[9, 9, 9], strides=[2, 2, 2], padding=["valid", "valid", "valid"],
    activation=["gelu", "gelu", "gelu"])
REFERENCE_D_MODEL = 192  # config/model.yaml:20


def load_reference_yaml(path: str):
    """Read speech_config / model_config.subsampling_config / d_model from a reference YAML
    (config/model.yaml) with PyYAML — hydra is not needed for the values."""
    with open(path) as fh:
        cfg = yaml.safe_load(fh)
    mc = cfg.get("model_config", {})
    speech = dict(cfg["speech_config"])
    # The reference loads its YAML through hydra/OmegaConf, whose resolver reads `1e-9` (no decimal point) as a
    # float; plain PyYAML (YAML 1.1) leaves it a string.  Same result here.
    for k in ("output_floor", "preemphasis", "lower_edge_hertz", "upper_edge_hertz", "padding"):
        if isinstance(speech.get(k), str):
            speech[k] = float(speech[k])
    return speech, dict(mc.get("subsampling_config", {})), int(mc.get("d_model", REFERENCE_D_MODEL))


class FrontEnd:
    def __init__(self, speech_config: dict | None = None, subsampling_config: dict | None = None,
                 model_dim: int = REFERENCE_D_MODEL, math: str = "tf32", device=None, seed: int = 0,
                 lean_intermediates: bool = True, single_pass: bool = True, assume_collated: bool = True):
        self.featurizer = SpeechFeaturizer(**(speech_config or REFERENCE_SPEECH_CONFIG))
        # The reference's batches come out of `padded_batch` (src/dataset.py:236-252): padded to their longest utterance,
        # so max(lengths) == N_max and every shape of the step (T_max, the mask width max(len3)) follows from the tensor
        # shape on the host.  With assume_collated=True a call without `max_length` takes N_max for it and only enqueues;
        # False restores the data-dependent mask width for over-padded batches (one device->host read of max(len3)).
        self.assume_collated = bool(assume_collated)
        self.subsampling = Conv1DSubsamplingLayer(
            model_dim=model_dim, subsampling_config=subsampling_config or REFERENCE_SUBSAMPLING_CONFIG,
            input_dim=self.featurizer.num_feature_bins, math=math, seed=seed, name="asr_encoder_conv_subsampling")
        self.device = torch.device(device) if device is not None else None
        # The features and the first two layers' activations are intermediates of this object: far inside the
        # collate padding nobody reads them (the ragged layers treat those rows as one constant row), so they are
        # not written there.  Outputs are bit-identical; `return_features=True` always gets the full 0.0 padding.
        self.lean_intermediates = bool(lean_intermediates)
        # normalize_signal needs max|x| of the whole utterance before its first frame (src/speech_featurizer.py:68-72):
        # a second pass over the waveform.  The featurizer is scale-covariant, so the log-mel kernel can find the
        # peak while it reads the samples, work on the un-normalised signal, and let the first separable conv add
        # 2 log(gain) and the floor as it reads the features (float32 rounding differences only).  Used when the
        # features are not returned; `return_features=True` takes the two-pass path and gets the reference values.
        self.single_pass = bool(single_pass)

    def set_weights(self, weights, device=None):
        self.subsampling.set_weights(weights, device or self.device)

    def __call__(self, wav: torch.Tensor, lengths: torch.Tensor | None = None, return_features: bool = False,
                 max_length: int | None = None):
        """wav [B, N_max] float32 CUDA, lengths [B] int32 CUDA ->
        (encoder_input [B, T3, d], padding_mask [B, max(len3)], len3 [B] int32[, features, n_frames]).

        `max_length` = max(lengths) when the caller knows it on the host (a collate does): the
        feature tensor is then padded to exactly the batch maximum and the mask width is derived
        on the host, so the step enqueues without any device->host synchronisation."""
        t_max = None
        if max_length is None and (lengths is None or self.assume_collated):
            max_length = int(wav.shape[1])
        if max_length is not None:
            t_max = max(0, int(self.featurizer.get_nframes(int(max_length))))
        sub = self.subsampling
        lean = (self.lean_intermediates and sub.math == "tf32" and sub.assume_zero_padding
                and all(p == "valid" for p in sub.padding))
        ragged = sub.math == "tf32" and sub.assume_zero_padding and all(p == "valid" for p in sub.padding)
        fill = sub.ragged_margin() if (lean and not return_features) else None
        gain = None
        if (self.single_pass and ragged and not return_features and self.featurizer.supports_single_pass()
                and self.featurizer.feature_type == "log_mel_spectrogram"):
            feats, n_frames, gain = self.featurizer.featurize_batch(wav, lengths, t_max=t_max, pad_fill_rows=fill, single_pass=True)
        else:
            feats, n_frames = self.featurizer.featurize_batch(wav, lengths, t_max=t_max, pad_fill_rows=fill)
        out, mask, len_all = sub(feats, mask=n_frames, return_lengths=True,
                                 max_frames=t_max if max_length is not None else None, lean_intermediates=lean,
                                 input_gain=gain)
        len3 = len_all[-1]
        if return_features:
            return out, mask, len3, feats, n_frames
        return out, mask, len3


class ConformerFrontEnd:
    """The conformer configuration's front end (config/conformer.yaml): waveforms -> log-mel -> Conv2dSubsampling
    (src/models/conformer/encoder.py:9-67) -> ([B, ceil(T/4), 20*filters], lengths ceil(n_frames/2) as the reference
    computes them).  Same device-side economy as FrontEnd: one pass over the waveform (the first convolution applies
    the deferred gain and floor), the feature tensor written only where it is read, the collate padding of the conv
    outputs filled with its constant pattern instead of being computed."""

    def __init__(self, speech_config: dict | None = None, subsampling_config: dict | None = None, device=None, seed: int = 0,
                 single_pass: bool = True):
        from .conv2d_subsampling import Conv2dSubsampling
        self.featurizer = SpeechFeaturizer(**(speech_config or REFERENCE_SPEECH_CONFIG))
        self.subsampling = Conv2dSubsampling(subsampling_config or dict(name="conv2d", filters=144, kernel_size=3, strides=2,
                                                                        padding="same"), seed=seed)
        self.device = torch.device(device) if device is not None else None
        self.single_pass = bool(single_pass)

    def set_weights(self, weights, device=None):
        self.subsampling.set_weights(weights, device or self.device)

    def __call__(self, wav: torch.Tensor, lengths: torch.Tensor | None = None, max_length: int | None = None):
        t_max = None
        if max_length is not None:
            t_max = max(0, int(self.featurizer.get_nframes(int(max_length))))
        sub = self.subsampling
        if (self.single_pass and sub.assume_zero_padding and self.featurizer.supports_single_pass()
                and self.featurizer.feature_type == "log_mel_spectrogram"):
            feats, n_frames, gain = self.featurizer.featurize_batch(wav, lengths, t_max=t_max, pad_fill_rows=0, single_pass=True)
            out, out_len = sub([feats, n_frames], input_gain=gain)
        else:
            feats, n_frames = self.featurizer.featurize_batch(wav, lengths, t_max=t_max)
            out, out_len = sub([feats, n_frames])
        return out, out_len


class CapturedFrontEnd:
    """The whole step (peak -> log-mel -> 3x separable conv -> lengths/mask) captured once into a CUDA graph
    for a fixed batch shape and replayed with one launch: the seven kernel launches of a step and the
    allocator calls between them disappear from the host path, and the ~2 us inter-kernel gaps shrink.

    Static shape contract: `wav` [batch, n_max] and `lengths` [batch] are device buffers owned by this object —
    write the batch into them (e.g. `PackedBatch.unpack` can target them, or `load()` copies), then `replay()`.
    Ragged batches are handled by the kernels themselves (they read `lengths` on the device), so one graph
    serves every batch of that shape.  Outputs are static tensors too, padded for n_max: `encoder_input`
    [batch, T3(n_max), d], `padding_mask` [batch, T3(n_max)], `len3` [batch]; the reference's mask width
    max(len3) (encoder.py:44) is `mask_width(max_length)` columns of it."""

    def __init__(self, frontend: FrontEnd, batch: int, n_max: int, device):
        self.fe = frontend
        self.device = torch.device(device)
        self.n_max = -(-int(n_max) // 4) * 4
        self.max_length = int(n_max)
        with torch.cuda.device(self.device):
            self.wav = torch.zeros((batch, self.n_max), dtype=torch.float32, device=self.device)
            self.lengths = torch.zeros((batch,), dtype=torch.int32, device=self.device)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):           # warm-up outside capture: plans, attributes, table uploads
                for _ in range(2):
                    self.fe(self.wav, self.lengths, max_length=self.max_length)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self.device)
            from . import _native
            l0 = _native.lib().tasr_launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.encoder_input, self.padding_mask, self.len3 = self.fe(self.wav, self.lengths, max_length=self.max_length)
            self.kernels_per_replay = int(_native.lib().tasr_launch_count() - l0)
        # the graph holds raw pointers to the plans' packed weights: a later set_weights() would free them
        self._plans_at_capture = tuple(self.fe.subsampling._plans or ())

    def load(self, wav: torch.Tensor, lengths: torch.Tensor) -> None:
        """Copy a [batch, <= n_max] device batch into the static input buffers (stream ordered)."""
        self.wav[:, : wav.shape[1]].copy_(wav, non_blocking=True)
        self.lengths.copy_(lengths.to(torch.int32), non_blocking=True)

    def replay(self):
        if tuple(self.fe.subsampling._plans or ()) != self._plans_at_capture:
            raise RuntimeError("CapturedFrontEnd: the subsampling weights changed after capture (set_weights destroys the "
                               "plans the graph points to); capture a new CapturedFrontEnd")
        self.graph.replay()
        return self.encoder_input, self.padding_mask, self.len3

    __call__ = replay

    def mask_width(self, max_length: int) -> int:
        """max(len3) for a batch whose longest utterance has `max_length` samples (host arithmetic)."""
        from .subsampling import get_conv_length
        w = max(0, int(self.fe.featurizer.get_nframes(int(max_length))))
        sub = self.fe.subsampling
        for i in range(len(sub.kernel_size)):
            w = get_conv_length(w, sub.kernel_size[i], sub.padding[i], sub.strides[i])
        return max(w, 0)


class InterleavedFrontEnd:
    """`n_streams` captured steps (each with its own static buffers) replayed round-robin on their own CUDA
    streams, so that consecutive batches overlap on the device: the kernels of this path are latency-bound at
    two CTAs per SM, and the tail of one batch's kernel fills with the next batch's work (B200, config 3:
    389 -> 352 us per step with two streams; a third adds nothing).

        il = InterleavedFrontEnd(frontend, batch, n_max, device)
        for i, (wav, lengths) in enumerate(batches):
            slot = il.slot(i)                 # CapturedFrontEnd: write the batch into slot.wav / slot.lengths
            il.wait(i)                        # (stream-ordered) the slot's previous replay has finished
            slot.load(wav, lengths); enc, mask, len3 = il.replay(i)   # enqueued on the slot's stream
        il.join()                             # current stream waits for every slot
    """

    def __init__(self, frontend: FrontEnd, batch: int, n_max: int, device, n_streams: int = 2):
        self.device = torch.device(device)
        self.slots = [CapturedFrontEnd(frontend, batch, n_max, self.device) for _ in range(n_streams)]
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream() for _ in range(n_streams)]
        self.kernels_per_replay = self.slots[0].kernels_per_replay

    def slot(self, i: int) -> CapturedFrontEnd:
        return self.slots[i % len(self.slots)]

    def stream(self, i: int) -> torch.cuda.Stream:
        return self.streams[i % len(self.streams)]

    def fork(self) -> None:
        """Make every slot stream wait for what is already enqueued on the current stream."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)

    def replay(self, i: int):
        with torch.cuda.stream(self.stream(i)):
            return self.slot(i).replay()

    def load(self, i: int, wav: torch.Tensor, lengths: torch.Tensor) -> None:
        with torch.cuda.stream(self.stream(i)):
            self.slot(i).load(wav, lengths)

    def join(self) -> None:
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)
