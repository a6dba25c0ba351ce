"""B200-native drop-in for Conv1DSubsamplingLayer (src/models/moonshine/encoder.py:9-105).

Same constructor (`model_dim`, `subsampling_config`, regularizer kwargs accepted and unused like
Keras' separable convs ignore `kernel_regularizer`), same `__call__(inputs[B,T,F,1],
training=False, mask=None) -> (outputs[B,T3,model_dim], padding_mask[B,max(len3)] or None)`,
static `lengths_to_padding_mask`, `compute_output_shape`, `get_config`.

Config-key quirk kept on purpose: the layer reads `subsampling_config["activations"]` and the
YAML spells it `activation` (encoder.py:25 vs config/model.yaml:27), so the effective default is
tanh, gelu, gelu.

Additive: `mask` may also be int32 `lengths[B]` (frames per utterance, as the featurizer returns
them) which skips the any(bin != 0) reduction of model.py:80 / encoder.py:53-56;
`math="fp32"|"tf32"` picks CUDA-core FP32 or tcgen05 TF32 for the pointwise contraction.
All arithmetic runs in libtasr_b200.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _native

__all__ = ["Conv1DSubsamplingLayer", "get_conv_length"]


def get_conv_length(input_length: int, kernel_size: int, padding: str, strides: int) -> int:
    """Host-side scalar version of math_util.get_conv_length (src/utils/math_util.py:20-32) for
    static shapes: float32 arithmetic, truncating cast."""
    length = np.float32(input_length)
    if padding == "same":
        length = np.ceil(length / np.float32(strides))
    elif padding == "valid":
        length = (length - np.float32(kernel_size)) / np.float32(strides) + np.float32(1.0)
    return int(np.trunc(length))


class Conv1DSubsamplingLayer:
    def __init__(self, model_dim: int = 288, subsampling_config: dict | None = None,
                 kernel_regularizer=None, bias_regularizer=None, name: str = "conv1d_subsampling",
                 input_dim: int = 80, math: str = "tf32", seed: int | None = None,
                 assume_zero_padding: bool = True, **kwargs):
        subsampling_config = subsampling_config or {}
        self.name = name
        self.filters = [model_dim, 2 * model_dim, model_dim]                        # encoder.py:21
        self.kernel_size = list(subsampling_config.get("kernel_size", [9, 9, 9]))   # :22
        self.strides = list(subsampling_config.get("strides", [2, 2, 2]))           # :23
        self.padding = list(subsampling_config.get("padding", ["same", "same", "same"]))  # :24
        self.activations = list(subsampling_config.get("activations", ["tanh", "gelu", "gelu"]))  # :25
        if (len(self.kernel_size) != len(self.strides) or len(self.kernel_size) != len(self.padding)
                or len(self.kernel_size) != len(self.activations)):
            raise ValueError("kernel_size, strides, padding, and activation must have the same length.")  # :26-27
        for a in self.activations:
            if a not in _native.ACT_CODES:
                raise ValueError(f"unsupported activation {a!r}; supported: tanh, gelu, relu, linear")
        if math not in ("fp32", "tf32"):
            raise ValueError("math must be 'fp32' or 'tf32'")
        self.math = math
        # ragged mode (TF32 path, when lengths are known): rows t >= lengths[b] of the input are taken to be
        # the collate's 0.0 padding (src/dataset.py:241), so tiles deep inside the padding are filled with the
        # constant row the convolution produces there instead of being computed.  Same values everywhere.
        self.assume_zero_padding = bool(assume_zero_padding)
        self.input_dim = input_dim
        self._seed = seed
        self.weights: list[tuple[torch.Tensor, torch.Tensor, torch.Tensor]] | None = None  # per layer (dw, pw, bias)
        self._plans: list[int] | None = None
        self._device = None
        self._side_streams: dict = {}

    # ------------------------------------------------------------------ weights
    def layer_dims(self):
        cin = self.input_dim
        for cout in self.filters[: len(self.kernel_size)]:
            yield cin, cout
            cin = cout

    def build(self, device, seed: int | None = None):
        """Keras creates variables lazily at first call with glorot-uniform kernels and zero bias
        (encoder.py:31-40 forwards no initializers).  Same here, seeded."""
        g = torch.Generator(device="cpu")
        g.manual_seed(self._seed if seed is None and self._seed is not None else (seed or 0))
        ws = []
        for (cin, cout), k in zip(self.layer_dims(), self.kernel_size):
            lim_dw = math.sqrt(6.0 / (k * cin + k))
            lim_pw = math.sqrt(6.0 / (cin + cout))
            dw = (torch.rand((k, cin), generator=g) * 2 - 1) * lim_dw
            pw = (torch.rand((cin, cout), generator=g) * 2 - 1) * lim_pw
            ws.append((dw, pw, torch.zeros(cout)))
        self.set_weights(ws, device)

    def set_weights(self, weights, device=None):
        """weights: per layer (depthwise_kernel, pointwise_kernel, bias) in Keras shapes
        (k,Cin,1)/(1,Cin,Cout)/(Cout) or squeezed (k,Cin)/(Cin,Cout)/(Cout); numpy or torch.
        Keras layer names: `<name>_conv_{1,2,3}` (encoder.py:39)."""
        device = torch.device(device) if device is not None else (self._device or torch.device("cuda"))
        out = []
        for i, ((dw, pw, b), (cin, cout), k) in enumerate(zip(weights, self.layer_dims(), self.kernel_size)):
            dw = torch.as_tensor(np.asarray(dw) if not isinstance(dw, torch.Tensor) else dw, dtype=torch.float32)
            pw = torch.as_tensor(np.asarray(pw) if not isinstance(pw, torch.Tensor) else pw, dtype=torch.float32)
            b = torch.as_tensor(np.asarray(b) if not isinstance(b, torch.Tensor) else b, dtype=torch.float32)
            dw = dw.reshape(k, cin) if dw.numel() == k * cin else dw
            pw = pw.reshape(cin, cout) if pw.numel() == cin * cout else pw
            if tuple(dw.shape) != (k, cin) or tuple(pw.shape) != (cin, cout) or tuple(b.shape) != (cout,):
                raise ValueError(f"layer {i + 1}: expected dw {(k, cin)}, pw {(cin, cout)}, bias {(cout,)}; got "
                                 f"{tuple(dw.shape)}, {tuple(pw.shape)}, {tuple(b.shape)}")
            out.append((dw.to(device).contiguous(), pw.to(device).contiguous(), b.to(device).contiguous()))
        if len(out) != len(self.kernel_size):
            raise ValueError(f"expected weights for {len(self.kernel_size)} layers, got {len(out)}")
        self._destroy_plans()
        self.weights = out
        self._device = device

    def _layer_struct(self, i: int) -> _native.TasrSepConvLayer:
        dw, pw, b = self.weights[i]
        (cin, cout) = list(self.layer_dims())[i]
        return _native.TasrSepConvLayer(
            dw=dw.data_ptr(), pw=pw.data_ptr(), bias=b.data_ptr(), c_in=cin, c_out=cout,
            kernel=self.kernel_size[i], stride=self.strides[i], same=int(self.padding[i] == "same"),
            activation=_native.ACT_CODES[self.activations[i]])

    def _destroy_plans(self):
        if getattr(self, "_plans", None):
            for p in self._plans:
                try:
                    _native.lib().tasr_sepconv_plan_destroy(p)
                except Exception:
                    pass
        self._plans = None

    def __del__(self):
        self._destroy_plans()

    def _ensure_plans(self):
        if self._plans is not None:
            return
        L = _native.lib()
        plans = []
        with torch.cuda.device(self._device):
            st = _native.stream_ptr()
            for i in range(len(self.kernel_size)):
                ls = self._layer_struct(i)
                out = C.c_void_p()
                _native.check(L.tasr_sepconv_plan_create(C.byref(ls), C.byref(out), st))
                plans.append(out.value)
            # chain the padding rows: zeros into layer 1, each layer's constant output row into the next
            prev = None
            for pl in plans:
                _native.check(L.tasr_sepconv_plan_set_pad_row(pl, prev, st))
                prev = L.tasr_sepconv_plan_pad_row(pl)
        self._plans = plans

    # ------------------------------------------------------------------ reference API
    @staticmethod
    def lengths_to_padding_mask(lengths: torch.Tensor) -> torch.Tensor:
        """encoder.py:43-48: float32 [B, max(lengths)] with 1.0 where t < lengths[b]."""
        _native.require_cuda(lengths, "lengths")
        lengths = lengths.to(torch.int32).contiguous()
        B = lengths.numel()
        width = max(int(lengths.max().item()), 0) if B else 0
        mask = torch.empty((B, width), dtype=torch.float32, device=lengths.device)
        if B == 0 or width == 0:
            return mask
        # identity "layer" (k=1, s=1, valid keeps L) reuses the C entry point for the mask alone
        k = (C.c_int32 * 1)(1)
        s = (C.c_int32 * 1)(1)
        same = (C.c_int32 * 1)(0)
        tmp = torch.empty((1, B), dtype=torch.int32, device=lengths.device)
        with torch.cuda.device(lengths.device):
            _native.check(_native.lib().tasr_conv_lengths_mask(
                lengths.data_ptr(), B, 1, k, s, same, tmp.data_ptr(), mask.data_ptr(), width, _native.stream_ptr()))
        return mask

    def compute_output_shape(self, input_shape):  # encoder.py:73-92
        bsz, seq = input_shape[0], input_shape[1]
        cur = seq
        for i in range(len(self.kernel_size)):
            if cur is not None:
                cur = get_conv_length(cur, self.kernel_size[i], self.padding[i], self.strides[i])
        return (bsz, cur, self.filters[-1])

    def get_config(self):  # encoder.py:94-105
        return {
            "model_dim": self.filters[-1], "kernel_size": self.kernel_size, "filters": self.filters,
            "strides": self.strides, "padding": self.padding, "name": self.name,
            "activations": self.activations,
        }

    def conv_lengths(self, lengths: torch.Tensor, with_mask: bool = True, max_frames: int | None = None):
        """lengths [B] int32 CUDA -> (len_per_layer [n_layers, B] int32, padding_mask or None).
        `max_frames` = max(lengths) if the caller knows it on the host: the mask width
        max(len_last) (encoder.py:44) is then computed on the host (the length map is monotone),
        otherwise it costs one small device->host read."""
        lengths = lengths.to(torch.int32).contiguous()
        B = lengths.numel()
        n = len(self.kernel_size)
        len_out = torch.empty((n, B), dtype=torch.int32, device=lengths.device)
        k = (C.c_int32 * n)(*self.kernel_size)
        s = (C.c_int32 * n)(*self.strides)
        same = (C.c_int32 * n)(*[int(p == "same") for p in self.padding])
        L = _native.lib()
        if B == 0:
            return len_out, (torch.empty((0, 0), dtype=torch.float32, device=lengths.device) if with_mask else None)
        with torch.cuda.device(lengths.device):
            st = _native.stream_ptr()
            mask = None
            if with_mask:
                if max_frames is not None:
                    width = int(max_frames)
                    for i in range(n):
                        width = get_conv_length(width, self.kernel_size[i], self.padding[i], self.strides[i])
                    width = max(width, 0)
                else:
                    # the reference sizes the mask by max(lengths) (encoder.py:44): one small D2H read
                    _native.check(L.tasr_conv_lengths_mask(lengths.data_ptr(), B, n, k, s, same, len_out.data_ptr(), None, 0, st))
                    width = max(int(len_out[-1].max().item()), 0)
                mask = torch.empty((B, width), dtype=torch.float32, device=lengths.device)
                _native.check(L.tasr_conv_lengths_mask(lengths.data_ptr(), B, n, k, s, same, len_out.data_ptr(),
                                                       mask.data_ptr() if width else None, width, st))
            else:
                _native.check(L.tasr_conv_lengths_mask(lengths.data_ptr(), B, n, k, s, same, len_out.data_ptr(), None, 0, st))
        return len_out, mask

    @staticmethod
    def ragged_margin() -> int:
        """Rows past an utterance's data that the ragged kernels may read from their input (see
        tasr_sepconv_ragged_margin): what a lean producer has to keep filled."""
        return int(_native.lib().tasr_sepconv_ragged_margin())

    def _lengths_on_side_stream(self, lengths: torch.Tensor, max_frames: int | None):
        """The lengths + mask kernel enqueued on a per-device side stream that forks from the current stream;
        returns (len_all, mask, side_stream) — the caller joins with `current.wait_stream(side)`.  Outputs are
        allocated under the CURRENT stream before the fork (so the caching allocator never sees a cross-stream
        hand-over).  Without `max_frames` the mask width needs a device->host read, which would stall the fork:
        that case stays on the current stream, as does stage timing."""
        if max_frames is None or _native.stage_marks is not None:
            return None
        dev = lengths.device
        n = len(self.kernel_size)
        B = lengths.numel()
        width = int(max_frames)
        for i in range(n):
            width = get_conv_length(width, self.kernel_size[i], self.padding[i], self.strides[i])
        width = max(width, 0)
        lengths = lengths.to(torch.int32).contiguous()
        len_out = torch.empty((n, B), dtype=torch.int32, device=dev)
        mask = torch.empty((B, width), dtype=torch.float32, device=dev)
        k = (C.c_int32 * n)(*self.kernel_size)
        s = (C.c_int32 * n)(*self.strides)
        same = (C.c_int32 * n)(*[int(p == "same") for p in self.padding])
        side = self._side_streams.get(dev)
        with torch.cuda.device(dev):
            if side is None:
                side = self._side_streams[dev] = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            _native.check(_native.lib().tasr_conv_lengths_mask(
                lengths.data_ptr(), B, n, k, s, same, len_out.data_ptr(), mask.data_ptr() if width else None, width,
                side.cuda_stream))
        return len_out, mask, side

    def __call__(self, inputs: torch.Tensor, training: bool = False, mask=None, return_lengths: bool = False,
                 max_frames: int | None = None, lean_intermediates: bool = False, input_gain=None):
        """`lean_intermediates` (ragged TF32 path only): the outputs of all layers but the last are not
        materialised far inside the padding — nobody reads them there (the reference keeps no intermediate,
        encoder.py:58-68).  The returned tensor is bit-identical either way.
        `input_gain` (a speech_featurizer.DeferredGain): `inputs` are the raw features of
        `featurize_batch(single_pass=True)`; the first layer adds the gain and the floor as it reads them."""
        x = _native.require_cuda(inputs, "inputs")
        if x.dim() != 4 or x.shape[-1] != 1:
            raise ValueError(f"inputs must be [B, T, F, 1] (encoder.py:51 squeezes the last axis); got {tuple(x.shape)}")
        if x.dtype != torch.float32:
            raise ValueError("inputs must be float32")
        B, T, F, _ = x.shape
        if F != self.input_dim:
            raise ValueError(f"inputs have {F} feature bins, layer was built for {self.input_dim}")
        if self.weights is None:
            self.build(x.device)
        if self._device != x.device:
            raise ValueError(f"weights live on {self._device}, inputs on {x.device}")
        h = x.reshape(B, T, F)
        if not h.is_contiguous():
            h = h.contiguous()
        L = _native.lib()

        lengths = None
        prefix_lengths = mask is not None and mask.dim() == 1   # explicit frame counts: padding is a zero suffix
        if mask is not None:
            _native.require_cuda(mask, "mask")
            if mask.dim() == 1:                      # additive: frame counts straight from the featurizer
                lengths = mask.to(torch.int32).contiguous()
            elif mask.dim() in (2, 3):               # encoder.py:53-56 on a [B,T] / [B,T,F] 0/1 mask
                m = mask.to(torch.float32).contiguous()
                lengths = torch.empty((B,), dtype=torch.int32, device=x.device)
                with torch.cuda.device(x.device):
                    _native.check(L.tasr_count_nonzero_frames(m.data_ptr(), B, m.shape[1], m.shape[2] if m.dim() == 3 else 1,
                                                              lengths.data_ptr(), _native.stream_ptr()))
            else:
                raise ValueError("mask must be lengths [B], [B,T] or [B,T,F]")

        # lengths + padding mask depend on the frame counts only: they run on a side stream (forked here, joined after
        # the last layer; inside a CUDA-graph capture this becomes a parallel branch) instead of trailing the stack
        side_result = None
        if lengths is not None and B:
            side_result = self._lengths_on_side_stream(lengths, max_frames)

        use_tf32 = self.math == "tf32"
        # the ragged kernels (constant padding rows, tiles skipped by length) are written for 'valid' receptive fields;
        # a stack with a padding='same' layer (the reference constructor's default, encoder.py:24) runs dense
        all_valid = all(p == "valid" for p in self.padding)
        if input_gain is not None and not (use_tf32 and prefix_lengths and self.assume_zero_padding and all_valid):
            raise ValueError("input_gain needs the ragged TF32 path: math='tf32', assume_zero_padding=True and mask=n_frames [B]")
        if use_tf32:
            self._ensure_plans()
        with torch.cuda.device(x.device):
            st = _native.stream_ptr()
            t_in = T
            for i in range(len(self.kernel_size)):
                if self.padding[i] not in ("valid", "same"):
                    raise ValueError(f"padding must be 'valid' or 'same', got {self.padding[i]!r}")
                t_out = max(0, get_conv_length(t_in, self.kernel_size[i], self.padding[i], self.strides[i]))
                cout = self.filters[i]
                y = _native.empty((B, t_out, cout), torch.float32, x.device)
                if B and t_out:
                    if use_tf32 and prefix_lengths and self.assume_zero_padding and all_valid:
                        lean_i = lean_intermediates and i + 1 < len(self.kernel_size)
                        gain_i = input_gain if i == 0 else None
                        if lean_i or gain_i is not None:
                            _native.check(L.tasr_sepconv1d_tf32_ragged_lean(
                                self._plans[i], h.data_ptr(), lengths.data_ptr(), i, B, t_in, y.data_ptr(), t_out,
                                self.ragged_margin() if lean_i else -1,
                                C.byref(gain_i.struct) if gain_i is not None else None, st))
                        else:
                            _native.check(L.tasr_sepconv1d_tf32_ragged(self._plans[i], h.data_ptr(), lengths.data_ptr(), i,
                                                                       B, t_in, y.data_ptr(), t_out, st))
                    elif use_tf32:
                        _native.check(L.tasr_sepconv1d_tf32(self._plans[i], h.data_ptr(), B, t_in, y.data_ptr(), t_out, st))
                    else:
                        ls = self._layer_struct(i)
                        _native.check(L.tasr_sepconv1d_f32(h.data_ptr(), B, t_in, C.byref(ls), y.data_ptr(), t_out, st))
                _native.mark(f"sepconv_layer{i + 1}")
                h, t_in = y, t_out
        padding_mask, len_all = None, None
        if side_result is not None:
            len_all, padding_mask, side = side_result
            torch.cuda.current_stream(x.device).wait_stream(side)
        elif lengths is not None:
            len_all, padding_mask = self.conv_lengths(lengths, with_mask=True, max_frames=max_frames)
            _native.mark("lengths_mask")
        if return_lengths:
            return h, padding_mask, len_all
        return h, padding_mask

    call = __call__

    @staticmethod
    def create_audio_mask(audio_inputs: torch.Tensor, pad_value: float = 0.0) -> torch.Tensor:
        """ASRModel.create_masks' audio half (model.py:80): any(audio != pad, axis=-1) -> [B,T,F] float32.
        Kept for callers that still build the mask the reference's way; passing n_frames is cheaper."""
        x = _native.require_cuda(audio_inputs, "audio_inputs")
        if x.dtype != torch.float32:
            raise ValueError("audio_inputs must be float32")
        x = x.contiguous()
        V = x.shape[-1] if x.dim() > 0 else 1
        out = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _native.check(_native.lib().tasr_audio_mask(x.data_ptr(), out.numel(), int(V), float(pad_value), out.data_ptr(),
                                                        _native.stream_ptr()))
        return out
