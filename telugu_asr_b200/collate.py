"""Device-side collate: the padding/masking half of src/dataset.py:223-255 moved onto the GPU.

The reference featurises each utterance on a CPU thread and then `padded_batch`es the features
with 0.0 (src/dataset.py:236-252).  Here the ragged waveforms are packed once into a pinned,
16-byte-row-aligned [B, N_max] staging buffer, copied H2D asynchronously, and everything after
that (featurisation, zero padding of features, lengths, masks) happens on the device."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

__all__ = ["pack_waveforms", "PinnedBatch", "PackedBatch", "shard_by_length"]


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


class PackedBatch:
    """Ragged host batch, packed: only the valid samples cross PCIe (no padding), optionally still
    as the 16-bit PCM the reference decodes its wav files from (src/utils/data_util.py:31,
    tf.audio.decode_wav: float32 = int16 / 32768, exact).  `unpack()` runs the device half of the
    collate (csrc/ingest.cu: tasr_unpack_pcm16 / tasr_unpack_f32) into the padded
    [B, N_max] float32 layout the featurizer kernels read; everything is preallocated.

    Utterance b starts at sample offset[b] of the packed buffer, a multiple of 8 samples so that
    it is 16-byte aligned in either format."""

    ALIGN = 8

    def __init__(self, batch: int, n_max: int, device, pcm16: bool = True, capacity: int | None = None):
        self.batch, self.pcm16 = batch, bool(pcm16)
        self.n_max = -(-max(n_max, 4) // 4) * 4
        cap = capacity if capacity is not None else batch * (-(-self.n_max // self.ALIGN) * self.ALIGN)
        self.capacity = -(-cap // self.ALIGN) * self.ALIGN
        dt = torch.int16 if self.pcm16 else torch.float32
        self.host_packed = torch.zeros((self.capacity,), dtype=dt).pin_memory()
        self.dev_packed = torch.empty((self.capacity,), dtype=dt, device=device)
        self.dev_off = torch.empty((batch,), dtype=torch.int64, device=device)
        self.dev_len = torch.empty((batch,), dtype=torch.int32, device=device)
        self.host_off = torch.zeros((batch,), dtype=torch.int64).pin_memory()
        self.host_len = torch.zeros((batch,), dtype=torch.int32).pin_memory()
        self.dev_wav = torch.empty((batch, self.n_max), dtype=torch.float32, device=device)
        self.used = 0
        self.max_len = 0

    @staticmethod
    def offsets_for(lengths, align: int = 8):
        lengths = np.asarray(lengths, dtype=np.int64).reshape(-1)
        padded = -(-lengths // align) * align
        off = np.zeros(len(lengths), dtype=np.int64)
        if len(lengths) > 1:
            off[1:] = np.cumsum(padded)[:-1]
        return off, int(padded.sum())

    def fill(self, waveforms: Sequence) -> None:
        """waveforms: B 1-D arrays, int16 PCM or float32 (float32 is converted to PCM with
        round(x*32768) when pcm16=True; values must then be k/32768 for the ingest to be exact)."""
        if len(waveforms) != self.batch:
            raise ValueError(f"expected {self.batch} utterances, got {len(waveforms)}")
        lens = [int(np.asarray(w).reshape(-1).shape[0]) for w in waveforms]
        if lens and max(lens) > self.n_max:
            raise ValueError(f"an utterance has {max(lens)} samples > n_max={self.n_max}")
        off, used = self.offsets_for(lens, self.ALIGN)
        if used > self.capacity:
            raise ValueError(f"packed batch needs {used} samples > capacity={self.capacity}")
        hp = self.host_packed.numpy()
        for b, w in enumerate(waveforms):
            w = np.asarray(w).reshape(-1)
            if self.pcm16 and w.dtype != np.int16:
                q = np.round(w.astype(np.float32) * 32768.0)
                if q.size and (q.max() > 32767 or q.min() < -32768):
                    raise ValueError(f"utterance {b}: float samples outside the int16 PCM range")
                w = q.astype(np.int16)
            elif not self.pcm16:
                w = w.astype(np.float32, copy=False)
            hp[off[b]: off[b] + lens[b]] = w
        self.host_off.numpy()[:] = off
        self.host_len.numpy()[:] = np.asarray(lens, dtype=np.int32)
        self.used = used
        self.max_len = max(lens) if lens else 0

    def to_device(self, non_blocking: bool = True):
        n = self.used
        self.dev_packed[:n].copy_(self.host_packed[:n], non_blocking=non_blocking)
        self.dev_off.copy_(self.host_off, non_blocking=non_blocking)
        self.dev_len.copy_(self.host_len, non_blocking=non_blocking)
        return self.dev_packed, self.dev_off, self.dev_len

    def unpack(self):
        """Device collate: packed -> (wav [B, N_max] float32, lengths [B] int32) on the current stream.
        Samples beyond lengths[b] are left as they are (no kernel reads them)."""
        from . import _native
        _native.require_cuda(self.dev_packed, "packed")
        L = _native.lib()
        fn = L.tasr_unpack_pcm16 if self.pcm16 else L.tasr_unpack_f32
        with torch.cuda.device(self.dev_wav.device):
            _native.check(fn(self.dev_packed.data_ptr(), self.dev_off.data_ptr(), self.dev_len.data_ptr(),
                             self.batch, int(self.max_len), self.dev_wav.data_ptr(), self.n_max,
                             _native.stream_ptr()))
        return self.dev_wav, self.dev_len

    @property
    def h2d_bytes(self) -> int:
        return self.used * (2 if self.pcm16 else 4) + self.batch * (8 + 4)
