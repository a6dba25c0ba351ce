"""B200-native drop-in for the conformer configuration's Conv2dSubsampling (src/models/conformer/encoder.py:9-73).

Same constructor (`subsampling_config` with filters / kernel_size / strides / padding as in config/conformer.yaml:22-27;
regularizer / initializer kwargs accepted), same call convention `layer([outputs, outputs_length], training=False)
-> (outputs [B, T'', F''*filters], outputs_length)`: conv1 -> relu -> conv2 -> relu -> merge_two_last_dims, the
lengths passed through get_conv_length ONCE (encoder.py:59-64 — ceil(L/2), although time shrinks by four; kept).
The second convolution (filters -> filters, 9 taps) runs as an implicit GEMM on tcgen05 (FP16 operands, FP32
accumulate); everything runs in libtasr_b200.so, there is no CPU path.  Only the reference configuration's geometry is
built: kernel_size 3, strides 2, padding "same"."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _native

__all__ = ["Conv2dSubsampling"]


class Conv2dSubsampling:
    def __init__(self, subsampling_config: dict | None = None, kernel_regularizer=None, bias_regularizer=None,
                 kernel_initializer=None, bias_initializer=None, name: str = "Conv2dSubsampling", seed: int | None = None,
                 assume_zero_padding: bool = True, **kwargs):
        subsampling_config = subsampling_config or {}
        self.name = name
        self.filter = int(subsampling_config.get("filters", 128))          # encoder.py:22
        self.kernel_size = subsampling_config.get("kernel_size", 3)         # :23
        self.stride = subsampling_config.get("strides", 2)                  # :24
        self.padding = subsampling_config.get("padding", "same")            # :25
        if self.kernel_size != 3 or self.stride != 2 or self.padding != "same":
            raise NotImplementedError("Conv2dSubsampling kernels are built for kernel_size=3, strides=2, padding='same' "
                                      "(config/conformer.yaml:22-27)")
        self._seed = seed
        # ragged mode: feature rows t >= outputs_length[b] are taken to be the collate's 0.0 padding (src/dataset.py:241);
        # tiles deep inside the padding are filled with the pattern the convolutions produce there.  Same values everywhere.
        self.assume_zero_padding = bool(assume_zero_padding)
        self._ragged_w = None
        self.weights = None       # [(w1 [3,3,1,F], b1 [F]), (w2 [3,3,F,F], b2 [F])] on the device
        self._plan = None
        self._device = None

    # ------------------------------------------------------------------ weights
    def build(self, device, seed: int | None = None):
        """Keras builds lazily with glorot_uniform kernels and zero biases; same here, seeded."""
        g = torch.Generator(device="cpu")
        g.manual_seed(self._seed if seed is None and self._seed is not None else (seed or 0))
        ws = []
        for cin in (1, self.filter):
            lim = math.sqrt(6.0 / (9 * cin + 9 * self.filter))
            ws.append(((torch.rand((3, 3, cin, self.filter), generator=g) * 2 - 1) * lim, torch.zeros(self.filter)))
        self.set_weights(ws, device)

    def set_weights(self, weights, device=None):
        """weights: [(kernel1 (3,3,1,F), bias1 (F)), (kernel2 (3,3,F,F), bias2 (F))], numpy or torch (Keras layer
        names `<name>_1`, `<name>_2`, encoder.py:31,41)."""
        device = torch.device(device) if device is not None else (self._device or torch.device("cuda"))
        F = self.filter
        out = []
        for i, ((w, b), cin) in enumerate(zip(weights, (1, F))):
            w = torch.as_tensor(np.asarray(w) if not isinstance(w, torch.Tensor) else w, dtype=torch.float32)
            b = torch.as_tensor(np.asarray(b) if not isinstance(b, torch.Tensor) else b, dtype=torch.float32)
            if tuple(w.shape) != (3, 3, cin, F) or tuple(b.shape) != (F,):
                raise ValueError(f"conv{i + 1}: expected kernel {(3, 3, cin, F)} and bias {(F,)}; got {tuple(w.shape)}, {tuple(b.shape)}")
            out.append((w.to(device).contiguous(), b.to(device).contiguous()))
        if len(out) != 2:
            raise ValueError("expected weights for two Conv2D layers")
        self._destroy_plan()
        self._ragged_w = None
        self.weights = out
        self._device = device

    def _destroy_plan(self):
        if getattr(self, "_plan", None):
            try:
                _native.lib().tasr_conv2d_plan_destroy(self._plan)
            except Exception:
                pass
        self._plan = None

    def __del__(self):
        self._destroy_plan()

    def _ensure_plan(self):
        if self._plan is not None:
            return
        (w1, b1), (w2, b2) = self.weights
        out = C.c_void_p()
        with torch.cuda.device(self._device):
            _native.check(_native.lib().tasr_conv2d_plan_create(w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                                                self.filter, C.byref(out), _native.stream_ptr()))
        self._plan = out.value

    # ------------------------------------------------------------------ reference API
    @staticmethod
    def output_shape(t: int, w: int):
        h1, w1, h2, w2 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _native.check(_native.lib().tasr_conv2d_output_shape(int(t), int(w), C.byref(h1), C.byref(w1), C.byref(h2), C.byref(w2)))
        return h1.value, w1.value, h2.value, w2.value

    def compute_output_shape(self, input_shape):
        _, _, h2, w2 = self.output_shape(input_shape[1], input_shape[2])
        return (input_shape[0], h2, w2 * self.filter)

    def __call__(self, inputs, training: bool = False, input_gain=None, **kwargs):
        """`input_gain` (a speech_featurizer.DeferredGain): the features are the raw output of
        `featurize_batch(single_pass=True)`; the first convolution adds the gain and the floor as it reads them
        (needs lengths and assume_zero_padding)."""
        outputs, outputs_length = inputs                                   # encoder.py:56
        x = _native.require_cuda(outputs, "outputs")
        if x.dim() != 4 or x.shape[-1] != 1:
            raise ValueError(f"outputs must be [B, T, F, 1] (Conv2D over one input channel); got {tuple(x.shape)}")
        if x.dtype != torch.float32:
            raise ValueError("outputs must be float32")
        B, T, W, _ = x.shape
        if self.weights is None:
            self.build(x.device)
        if self._device != x.device:
            raise ValueError(f"weights live on {self._device}, inputs on {x.device}")
        self._ensure_plan()
        x = x.contiguous()
        h1, w1, h2, w2 = self.output_shape(T, W)
        F = self.filter
        work = _native.empty((B, h1, w1, F), torch.float16, x.device)   # conv1's output: FP16, only the tensor cores read it
        out = _native.empty((B, h2, w2 * F), torch.float32, x.device)
        L = _native.lib()
        with torch.cuda.device(x.device):
            st = _native.stream_ptr()
            ln = None
            if outputs_length is not None:
                ln = _native.require_cuda(outputs_length, "outputs_length").to(torch.int32).contiguous()
            if ln is not None and self.assume_zero_padding and B and T:
                if self._ragged_w != W:        # once per feature width (synchronises; not inside a graph capture)
                    _native.check(L.tasr_conv2d_plan_prepare_ragged(self._plan, W, st))
                    self._ragged_w = W
                _native.check(L.tasr_conv2d_subsample_ragged(self._plan, x.data_ptr(), ln.data_ptr(), B, T, W,
                                                             work.data_ptr(), out.data_ptr(),
                                                             C.byref(input_gain.struct) if input_gain is not None else None, st))
            elif input_gain is not None:
                raise ValueError("input_gain needs outputs_length and assume_zero_padding=True (the ragged path)")
            else:
                _native.check(L.tasr_conv2d_subsample(self._plan, x.data_ptr(), B, T, W, work.data_ptr(), out.data_ptr(), st))
            out_len = None
            if ln is not None:
                out_len = torch.empty((1, ln.numel()), dtype=torch.int32, device=x.device)
                if ln.numel():
                    k, s, same = (C.c_int32 * 1)(3), (C.c_int32 * 1)(2), (C.c_int32 * 1)(1)
                    _native.check(L.tasr_conv_lengths_mask(ln.data_ptr(), ln.numel(), 1, k, s, same, out_len.data_ptr(), None, 0, st))
                out_len = out_len[0]
        return out, out_len

    call = __call__

    def get_config(self):
        return {"name": self.name, "filters": self.filter, "kernel_size": self.kernel_size, "strides": self.stride,
                "padding": self.padding}
