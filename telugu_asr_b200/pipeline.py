"""Host-to-host front end: the reference's loader + featurizer + subsampling hand-off as a
three-stage device pipeline.

The reference decodes each wav file to float32 on a CPU thread, featurises it there, pads the batch
and only then ships features to the GPU (src/dataset.py:167-197, 223-255; src/utils/data_util.py:31).
Here the host hands over the raw utterances — packed, valid samples only, optionally still int16 PCM —
and receives the encoder input, padding mask and lengths in pinned host memory:

    copy-in stream   H2D of the packed batch                      (PackedBatch.to_device)
    compute stream   unpack -> peak -> log-mel -> 3x sepconv -> lengths/mask
    copy-out stream  D2H of [B,T3,d], mask, len3

Slots are double buffered, so batch i+1 crosses PCIe while batch i computes and batch i-1 returns.
Stream ordering is by CUDA events only; the host never blocks inside `submit`.

`graph=True` (default): the device side of a slot — unpack, the whole front end and the three D2H copies — is
captured once per slot into a CUDA graph on the slot's own stream, so a submit costs the host three
`cudaMemcpyAsync`, a few event calls and ONE graph launch instead of ~25 launches / allocator calls; a slow or
contended host thread (measured: 0.25 -> 1.7 ms per eager submit on a busy box, which made the path host-bound)
no longer shows in the throughput.  Shapes are then static, sized for `n_max`; the ticket returns views cut to
the batch's own maximum length, i.e. the same tensors as the eager mode.
"""
from __future__ import annotations

from typing import Sequence

import torch

from .collate import PackedBatch
from .frontend import FrontEnd

__all__ = ["FrontEndPipeline", "Ticket"]


class Ticket:
    """Result handle of one submitted batch; `wait()` blocks until its D2H has landed."""

    def __init__(self, slot: int, done: torch.cuda.Event, out, mask, len3, offsets=None):
        self.slot, self.done = slot, done
        self._out, self._mask, self._len3, self._offsets = out, mask, len3, offsets

    def wait(self):
        """(encoder_input [B, T3, d], padding_mask [B, max(len3)], len3 [B]) in pinned host memory; with the pipeline's
        `packed_output=True`: (rows [sum(len3), d], offsets [B + 1] int64, len3 [B]) - utterance b is rows[offsets[b]:offsets[b+1]]."""
        self.done.synchronize()
        if self._offsets is not None:
            return self._out, self._offsets, self._len3
        return self._out, self._mask, self._len3


class FrontEndPipeline:
    def __init__(self, frontend: FrontEnd, batch: int, n_max: int, device, pcm16: bool = True, slots: int = 2,
                 graph: bool = True, packed_output: bool = False):
        self.fe = frontend
        # packed_output: only the valid rows of the encoder input cross the host link on the way back ([sum(len3), d] +
        # offsets instead of the zero-padded [B, T3, d] + mask: about half the D2H bytes of a ragged batch).  The row count
        # of a batch is known on the host (integer arithmetic on the staged lengths), so the copy is sized without a sync.
        self.packed_output = bool(packed_output)
        self.device = torch.device(device)
        self.batch, self.pcm16, self.slots = batch, pcm16, slots
        self.n_max = int(n_max)
        self.graph = bool(graph)
        self._graphs = [None] * slots          # per slot: (CUDAGraph, host out/mask/len3, plans at capture)
        self.kernels_per_submit = None         # kernels inside one captured submit (graph mode)
        self.staging = [PackedBatch(batch, n_max, self.device, pcm16=pcm16) for _ in range(slots)]
        with torch.cuda.device(self.device):
            self.s_in = torch.cuda.Stream()
            self.s_out = torch.cuda.Stream()
            self.s_comp = [torch.cuda.Stream() for _ in range(slots)]      # graph mode: one stream per slot
            self.ev_ready = [torch.cuda.Event() for _ in range(slots)]     # H2D of slot landed
            self.ev_free = [torch.cuda.Event() for _ in range(slots)]      # packed device buffer consumed
            self.ev_done = [torch.cuda.Event() for _ in range(slots)]      # D2H of slot landed
            self.ev_comp = [torch.cuda.Event() for _ in range(slots)]      # compute of slot finished
        self._host = [None] * slots    # pinned (out, mask, len3) per slot, sized on first use
        self._used = [False] * slots
        self.d2h_bytes = 0

    def stage(self, slot: int, waveforms: Sequence) -> None:
        """Host half of the collate: pack the utterances into the slot's pinned buffer."""
        if self._used[slot]:
            self.ev_free[slot].synchronize()     # the previous H2D+unpack of this slot must be over
        self.staging[slot].fill(waveforms)

    def _host_buffers(self, slot: int, out, mask, len3):
        hb = self._host[slot]
        if hb is None or hb[0].shape != out.shape or hb[1].shape != mask.shape:
            hb = (torch.empty(out.shape, dtype=out.dtype).pin_memory(),
                  torch.empty(mask.shape, dtype=mask.dtype).pin_memory(),
                  torch.empty(len3.shape, dtype=len3.dtype).pin_memory())
            self._host[slot] = hb
        return hb

    # ------------------------------------------------------------------ graph mode
    def _capture(self, slot: int):
        from . import _native
        from .subsampling import get_conv_length
        pb = self.staging[slot]
        st = self.s_comp[slot]
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            st.wait_stream(cur)
            st.wait_stream(self.s_in)
            saved_max = pb.max_len
            pb.max_len = self.n_max                  # static launch geometry: the kernels read the lengths on the device
            try:
                with torch.cuda.stream(st):
                    if not self._used[slot]:         # nothing staged yet: lengths must not be garbage for the warm-up
                        pb.dev_len.zero_()
                        pb.dev_off.zero_()
                    for _ in range(2):               # warm-up outside capture: plans, attributes, table uploads
                        wav, lens = pb.unpack()
                        out, mask, len3 = self.fe(wav, lens, max_length=self.n_max)
                    hb = (torch.empty(out.shape, dtype=out.dtype).pin_memory(),
                          torch.empty(mask.shape, dtype=mask.dtype).pin_memory(),
                          torch.empty(len3.shape, dtype=len3.dtype).pin_memory())
                    dev_rows = dev_offs = None
                    if self.packed_output:
                        dev_rows = torch.empty((out.shape[0] * out.shape[1], out.shape[2]), dtype=out.dtype, device=self.device)
                        dev_offs = torch.empty((out.shape[0] + 1,), dtype=torch.int64, device=self.device)
                st.synchronize()
                l0 = _native.lib().tasr_launch_count()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    wav, lens = pb.unpack()
                    out, mask, len3 = self.fe(wav, lens, max_length=self.n_max)
                    if self.packed_output:      # compaction inside the graph; the (batch-sized) D2H follows the replay
                        _native.check(_native.lib().tasr_pack_valid_rows(out.data_ptr(), len3.data_ptr(), out.shape[0], out.shape[1],
                                                                         out.shape[2], dev_rows.data_ptr(), dev_offs.data_ptr(),
                                                                         _native.stream_ptr()))
                    else:
                        hb[0].copy_(out, non_blocking=True)
                        hb[1].copy_(mask, non_blocking=True)
                    hb[2].copy_(len3, non_blocking=True)
                self.kernels_per_submit = int(_native.lib().tasr_launch_count() - l0)
            finally:
                pb.max_len = saved_max
        self._graphs[slot] = (g, hb, tuple(self.fe.subsampling._plans or ()), (out, mask, len3, dev_rows, dev_offs))
        self.d2h_bytes = (hb[0].numel() + hb[1].numel()) * 4 + hb[2].numel() * 4

    def _views(self, slot: int, max_len: int):
        """The static host buffers cut to what the eager path returns for a batch whose longest utterance has
        `max_len` samples: [B, T3(max_len), d], [B, max(len3)], [B]."""
        from .subsampling import get_conv_length
        sub = self.fe.subsampling
        t = max(0, int(self.fe.featurizer.get_nframes(int(max_len))))
        for i in range(len(sub.kernel_size)):
            t = max(0, get_conv_length(t, sub.kernel_size[i], sub.padding[i], sub.strides[i]))
        h_out, h_mask, h_len = self._graphs[slot][1]
        return h_out[:, :t], h_mask[:, :t], h_len

    def _submit_graph(self, slot: int) -> Ticket:
        pb = self.staging[slot]
        if self._graphs[slot] is None:
            self._capture(slot)
        g, hb, plans, _keep = self._graphs[slot]
        if tuple(self.fe.subsampling._plans or ()) != plans:
            raise RuntimeError("FrontEndPipeline: the subsampling weights changed after the slot's graph was captured; "
                               "build a new pipeline")
        st = self.s_comp[slot]
        with torch.cuda.device(self.device):
            if self._used[slot]:
                self.s_in.wait_event(self.ev_free[slot])
            else:
                self.s_in.wait_stream(st)            # the capture warm-up wrote the slot's device buffers
            with torch.cuda.stream(self.s_in):
                pb.to_device(non_blocking=True)
                self.ev_ready[slot].record(self.s_in)
            st.wait_event(self.ev_ready[slot])
            with torch.cuda.stream(st):
                g.replay()
                self.ev_free[slot].record(st)
                if self.packed_output:
                    len3_h = self._host_len3(pb.host_len.numpy())
                    offs = torch.zeros((len(len3_h) + 1,), dtype=torch.int64)
                    offs[1:] = torch.from_numpy(len3_h.clip(0, hb[0].shape[1]).astype("int64")).cumsum(0)
                    rows = int(offs[-1])
                    h_rows = hb[0].view(-1, hb[0].shape[-1])[:rows]
                    h_rows.copy_(_keep[3][:rows], non_blocking=True)
                    self.d2h_bytes = rows * hb[0].shape[-1] * 4 + hb[2].numel() * 4
                self.ev_done[slot].record(st)
        self._used[slot] = True
        if self.packed_output:
            return Ticket(slot, self.ev_done[slot], h_rows, None, hb[2], offsets=offs)
        return Ticket(slot, self.ev_done[slot], *self._views(slot, pb.max_len))

    def _host_len3(self, lengths):
        """len3 of every utterance from its sample count, on the host, with the reference's arithmetic
        (get_nframes, then get_conv_length per layer: float32, truncating cast) - the values the device computes."""
        import numpy as np
        sub, fz = self.fe.subsampling, self.fe.featurizer
        n = np.asarray(lengths, dtype=np.int64)
        if fz.pad_end:
            t = -(-n // fz.frame_step)
        else:
            t = 1 + (n - fz.frame_length) // fz.frame_step
        cur = np.maximum(t, 0).astype(np.int32)
        for k, p, s_ in zip(sub.kernel_size, sub.padding, sub.strides):   # src/utils/math_util.py:20-32, vectorised
            f = cur.astype(np.float32)
            if p == "same":
                f = np.ceil(f / np.float32(s_))
            else:
                f = (f - np.float32(k)) / np.float32(s_) + np.float32(1.0)
            cur = np.trunc(f).astype(np.int32)
        return cur

    def drain(self) -> None:
        """Block until everything submitted so far has landed in host memory."""
        self.s_out.synchronize()
        for st in self.s_comp:
            st.synchronize()

    def submit(self, slot: int) -> Ticket:
        """Enqueue H2D -> compute -> D2H for the staged slot; returns immediately."""
        if self.graph:
            return self._submit_graph(slot)
        pb = self.staging[slot]
        with torch.cuda.device(self.device):
            comp = torch.cuda.current_stream()
            if self._used[slot]:
                self.s_in.wait_event(self.ev_free[slot])
            with torch.cuda.stream(self.s_in):
                pb.to_device(non_blocking=True)
                self.ev_ready[slot].record(self.s_in)
            comp.wait_event(self.ev_ready[slot])
            wav, lens = pb.unpack()
            out, mask, len3 = self.fe(wav, lens, max_length=pb.max_len)
            # `lens` IS the slot's device length buffer and every kernel of the batch reads it, so the slot is free for
            # the next H2D only when the whole front end has run (not right after the unpack)
            self.ev_free[slot].record(comp)
            self.ev_comp[slot].record(comp)
            h_out, h_mask, h_len = self._host_buffers(slot, out, mask, len3)
            self.s_out.wait_event(self.ev_comp[slot])
            with torch.cuda.stream(self.s_out):
                h_out.copy_(out, non_blocking=True)
                h_mask.copy_(mask, non_blocking=True)
                h_len.copy_(len3, non_blocking=True)
                for t in (out, mask, len3):
                    t.record_stream(self.s_out)
                self.ev_done[slot].record(self.s_out)
        self._used[slot] = True
        self.d2h_bytes = (h_out.numel() + h_mask.numel()) * 4 + h_len.numel() * 4
        return Ticket(slot, self.ev_done[slot], h_out, h_mask, h_len)

    def run(self, waveforms: Sequence, slot: int = 0):
        """Convenience, one batch, blocking: utterances -> (encoder_input, mask, len3) on the host."""
        self.stage(slot, waveforms)
        return self.submit(slot).wait()

    @property
    def h2d_bytes(self) -> int:
        return self.staging[0].h2d_bytes
