"""Device-side SpecAugment: mirrors src/augmentations/specaugment.py (FreqMasking,
TimeMasking) and src/augmentations/augmentation.py (Augmentation: `prob`, `signal_augment`,
`feature_augment`, the `AUGMENTATIONS` registry and its KeyError) with the same constructor keywords.

The reference augments one utterance at a time on a CPU thread of the loader (src/dataset.py:170-172).
Here the draws for the whole batch are made on the host from the frame counts the collate already has
(same distributions: f ~ U{0..mask_factor-1} clipped to F, f0 ~ U{0..F-f-1}; t ~ U{0..mask_factor-1}
clipped to int(T*p_upperbound), t0 ~ U{0..T-t-1}; each augmentation applied when U[0,1) < prob) and the
masks are applied in place on the device by one launch of tasr_specaugment_f32.

Note: a time-masked frame becomes all zeros, which the reference's own mask rule
(model.py:80, any(x != 0)) would then miscount as padding; pass `n_frames` to the subsampling layer (as
FrontEnd does) rather than a mask derived from the augmented features.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native

__all__ = ["FreqMasking", "TimeMasking", "Augmentation", "AUGMENTATIONS"]


class FreqMasking:
    def __init__(self, num_masks: int = 1, mask_factor: int = 27):   # specaugment.py:7-13
        self.num_masks = num_masks
        self.mask_factor = mask_factor

    def draw(self, rng: np.random.Generator, T: int, F: int):
        """num_masks x (f0, f) for one utterance (specaugment.py:16-19)."""
        out = []
        for _ in range(self.num_masks):
            f = min(int(rng.integers(0, self.mask_factor)), F)
            f0 = int(rng.integers(0, F - f)) if F - f > 0 else 0
            out.append((f0, f))
        return out

    axis = "freq"


class TimeMasking:
    def __init__(self, num_masks: int = 1, mask_factor: float = 100, p_upperbound: float = 1.0):   # :34-37
        self.num_masks = num_masks
        self.mask_factor = mask_factor
        self.p_upperbound = p_upperbound

    def draw(self, rng: np.random.Generator, T: int, F: int):
        """num_masks x (t0, t) for one utterance (specaugment.py:43-50)."""
        out = []
        for _ in range(self.num_masks):
            t = int(rng.integers(0, int(self.mask_factor)))
            t = min(t, int(np.float32(T) * np.float32(self.p_upperbound)))
            t0 = int(rng.integers(0, T - t)) if T - t > 0 else 0
            out.append((t0, t))
        return out

    axis = "time"


AUGMENTATIONS = {"freq_masking": FreqMasking, "time_masking": TimeMasking}   # augmentation.py:5-8


class Augmentation:
    def __init__(self, config: dict | None = None, seed: int | None = None):
        if not config:
            config = {}
        self.prob = float(config.get("prob", 0.5))                                  # augmentation.py:14
        self.signal_augmentations = self.parse(config.get("signal_augment", {}))    # :15
        self.feature_augmentations = self.parse(config.get("feature_augment", {}))  # :16
        self.rng = np.random.default_rng(seed)

    @staticmethod
    def parse(config: dict) -> list:                                                # :61-79
        augmentations = []
        for key, value in (config or {}).items():
            au = AUGMENTATIONS.get(key, None)
            if au is None:
                raise KeyError(f"No tf augmentation named: {key}\n"
                               f"Available tf augmentations: {AUGMENTATIONS.keys()}")
            augmentations.append(au(**value) if value is not None else au())
        return augmentations

    def draw_masks(self, n_frames_host, F: int, augmentations=None):
        """Host draws for a batch: (time_masks [B,n_time,2], freq_masks [B,n_freq,2]) int32; an
        augmentation that loses its `prob` draw (augmentation.py:31-34) contributes width-0 masks."""
        augs = self.feature_augmentations if augmentations is None else augmentations
        n_time = sum(a.num_masks for a in augs if a.axis == "time")
        n_freq = sum(a.num_masks for a in augs if a.axis == "freq")
        B = len(n_frames_host)
        tm = np.zeros((B, max(n_time, 1), 2), dtype=np.int32)
        fm = np.zeros((B, max(n_freq, 1), 2), dtype=np.int32)
        for b in range(B):
            T = int(n_frames_host[b])
            it = jf = 0
            for a in augs:
                apply = self.rng.random() < self.prob
                masks = a.draw(self.rng, T, F) if (apply and T > 0) else [(0, 0)] * a.num_masks
                for m in masks:
                    if a.axis == "time":
                        tm[b, it] = m
                        it += 1
                    else:
                        fm[b, jf] = m
                        jf += 1
        return tm[:, :n_time], fm[:, :n_freq]

    @staticmethod
    def apply_masks(features: torch.Tensor, n_frames: torch.Tensor, time_masks, freq_masks) -> torch.Tensor:
        """In place on features [B,T_max,F,1] (or [B,T_max,F]) CUDA float32; masks as numpy/torch int32."""
        x = _native.require_cuda(features, "features")
        _native.require_cuda(n_frames, "n_frames")
        if x.dtype != torch.float32 or not x.is_contiguous():
            raise ValueError("features must be a contiguous float32 tensor")
        if x.dim() == 4 and x.shape[-1] != 1:
            raise ValueError("features must be [B,T,F,1] or [B,T,F]")
        B, T, F = x.shape[0], x.shape[1], x.shape[2]
        tm = torch.as_tensor(np.ascontiguousarray(time_masks), dtype=torch.int32).reshape(B, -1, 2).to(x.device)
        fm = torch.as_tensor(np.ascontiguousarray(freq_masks), dtype=torch.int32).reshape(B, -1, 2).to(x.device)
        nf = n_frames.to(torch.int32).contiguous()
        with torch.cuda.device(x.device):
            _native.check(_native.lib().tasr_specaugment_f32(
                x.data_ptr(), nf.data_ptr(), B, T, F, tm.data_ptr() if tm.numel() else None, tm.shape[1],
                fm.data_ptr() if fm.numel() else None, fm.shape[1], _native.stream_ptr()))
        return features

    def feature_augment(self, features: torch.Tensor, n_frames: torch.Tensor, n_frames_host=None) -> torch.Tensor:
        """Batched form of augmentation.py:48-58: draws on the host, masks on the device, in place."""
        if not self.feature_augmentations:
            return features
        host = n_frames.cpu().numpy() if n_frames_host is None else np.asarray(n_frames_host)
        tm, fm = self.draw_masks(host, features.shape[2])
        return self.apply_masks(features, n_frames, tm, fm)

    def signal_augment(self, signals: torch.Tensor) -> torch.Tensor:
        """augmentation.py:37-46.  The registry holds no waveform augmentation (both entries mask
        spectrograms), so a non-empty `signal_augment` config cannot be honoured."""
        if self.signal_augmentations:
            raise NotImplementedError("signal_augment: the reference's registry only has spectrogram augmentations")
        return signals
