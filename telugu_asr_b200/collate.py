"""Device-side collate: the padding/masking half of src/dataset.py:223-255 moved onto the GPU.

The reference featurises each utterance on a CPU thread and then `padded_batch`es the features
with 0.0 (src/dataset.py:236-252).  Here the ragged waveforms are packed once into a pinned,
16-byte-row-aligned [B, N_max] staging buffer, copied H2D asynchronously, and everything after
that (featurisation, zero padding of features, lengths, masks) happens on the device."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

__all__ = ["pack_waveforms", "PinnedBatch", "shard_by_length"]


class PinnedBatch:
    """Reusable pinned host staging buffer + matching device buffers (no per-step allocation)."""

    def __init__(self, batch: int, n_max: int, device, align: int = 4):
        self.batch = batch
        self.n_max = -(-n_max // align) * align
        self.host_wav = torch.zeros((batch, self.n_max), dtype=torch.float32).pin_memory()
        self.host_len = torch.zeros((batch,), dtype=torch.int32).pin_memory()
        self.dev_wav = torch.empty((batch, self.n_max), dtype=torch.float32, device=device)
        self.dev_len = torch.empty((batch,), dtype=torch.int32, device=device)

    def fill(self, waveforms: Sequence) -> None:
        if len(waveforms) != self.batch:
            raise ValueError(f"expected {self.batch} utterances, got {len(waveforms)}")
        hw = self.host_wav.numpy()
        hl = self.host_len.numpy()
        for b, w in enumerate(waveforms):
            w = np.asarray(w, dtype=np.float32).reshape(-1)
            if w.shape[0] > self.n_max:
                raise ValueError(f"utterance {b} has {w.shape[0]} samples > n_max={self.n_max}")
            hw[b, : w.shape[0]] = w
            hw[b, w.shape[0]:] = 0.0
            hl[b] = w.shape[0]

    def to_device(self, non_blocking: bool = True):
        self.dev_wav.copy_(self.host_wav, non_blocking=non_blocking)
        self.dev_len.copy_(self.host_len, non_blocking=non_blocking)
        return self.dev_wav, self.dev_len

    @property
    def h2d_bytes(self) -> int:
        return self.host_wav.numel() * 4 + self.host_len.numel() * 4


def pack_waveforms(waveforms: Sequence, device, align: int = 4):
    """list of 1-D float32 waveforms -> (wav [B, N_max] CUDA zero padded, lengths [B] int32 CUDA)."""
    n_max = max((int(np.asarray(w).shape[0]) for w in waveforms), default=0)
    pb = PinnedBatch(len(waveforms), max(n_max, align), device, align)
    pb.fill(waveforms)
    wav, ln = pb.to_device()
    torch.cuda.current_stream(wav.device).synchronize()  # pinned buffer is dropped on return
    return wav, ln


def shard_by_length(lengths, world_size: int) -> list[list[int]]:
    """Length-balanced sharding of utterances over ranks: longest first, each to
    the least-loaded rank; ties break toward the lower rank so the result is deterministic.
    The union of the shards is every index exactly once; no utterance is split."""
    lengths = np.asarray(lengths).reshape(-1)
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world_size
    shards: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (load[j], j))
        shards[r].append(i)
        load[r] += int(lengths[i])
    for s in shards:
        s.sort()
    return shards
