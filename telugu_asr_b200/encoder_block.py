"""B200-native drop-in for the first block of the Moonshine-style encoder:
`EncoderBlock` of src/models/moonshine/encoder.py:109-180 = MHSAModule (src/models/layers/attention.py:519-615, RoPE of
positional_encoding.py:20-93) followed by FFNModule (src/models/layers/mlp.py:9-60).

Same constructor keywords as the reference (input_dim, dropout, activation, num_heads, head_dim, fc_factor, initializer /
regularizer kwargs accepted) and the same call convention `block(inputs [B, T, d], training=False, use_causal_mask=False,
mask=padding_mask [B, T] or None) -> [B, T, d]`; it consumes the `[B, T3, 192]` tensor and the float padding mask (or the
`len3` lengths it was built from) that `Conv1DSubsamplingLayer` / `FrontEnd` return.  Inference only: `training=True` (dropout)
raises.  The dense layers run on tcgen05 (TF32, FP32 accumulate), attention on the CUDA cores; everything is in
libtasr_b200.so and there is no CPU path.  Built for the config/model.yaml shape (d = 192 = 6 heads x 32)."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _native

__all__ = ["EncoderBlock"]

_WEIGHT_NAMES = ("wq", "wk", "wv", "wo", "ln1_gamma", "ln1_beta", "w1", "b1", "w2", "b2", "ln2_gamma", "ln2_beta")


class EncoderBlock:
    def __init__(self, input_dim: int = 288, dropout: float = 0.1, activation: str = "gelu", num_heads: int = 8, head_dim: int = 36,
                 fc_factor: int = 1, kernel_initializer=None, bias_initializer=None, kernel_regularizer=None, bias_regularizer=None,
                 name: str = "odv_encoder_block", seed: int | None = None, **kwargs):
        # encoder.py:111-150 (the defaults are the reference's; config/model.yaml:27-33 builds 192 / 6 x 32 / fc_factor 1 / gelu)
        self.name = name
        self.input_dim = int(input_dim)
        self.dropout = float(dropout)
        self.activation = activation
        self.num_heads = int(num_heads)
        self.head_dim = int(head_dim)
        self.fc_factor = int(fc_factor)
        if activation != "gelu":
            raise NotImplementedError("EncoderBlock kernels implement activation='gelu' (config/model.yaml:32)")
        self.ln_eps = 1e-3                   # tf.keras.layers.LayerNormalization default
        self._seed = seed
        self.weights = None
        self._plan = None
        self._device = None

    # ------------------------------------------------------------------ weights
    def build(self, device, seed: int | None = None):
        """Keras builds lazily: glorot_uniform kernels, zero biases, LayerNorm gamma 1 / beta 0; same here, seeded."""
        g = torch.Generator(device="cpu")
        g.manual_seed(self._seed if seed is None and self._seed is not None else (seed or 0))
        d, hd, F = self.input_dim, self.num_heads * self.head_dim, self.input_dim * self.fc_factor

        def glorot(i, o):
            return (torch.rand((i, o), generator=g) * 2 - 1) * math.sqrt(6.0 / (i + o))

        self.set_weights({"wq": glorot(d, hd), "wk": glorot(d, hd), "wv": glorot(d, hd), "wo": glorot(hd, d),
                          "ln1_gamma": torch.ones(d), "ln1_beta": torch.zeros(d), "w1": glorot(d, F), "b1": torch.zeros(F),
                          "w2": glorot(F, d), "b2": torch.zeros(d), "ln2_gamma": torch.ones(d), "ln2_beta": torch.zeros(d)}, device)

    def set_weights(self, weights: dict, device=None):
        """weights: query_kernel / key_kernel / value_kernel [d, H*Dh], projection_kernel [H*Dh, d] (attention.py:47-70),
        the two LayerNormalization gamma / beta, dense1 / dense2 kernel + bias (mlp.py:31-47), under the short names
        wq wk wv wo ln1_gamma ln1_beta w1 b1 w2 b2 ln2_gamma ln2_beta; numpy or torch."""
        device = torch.device(device) if device is not None else (self._device or torch.device("cuda"))
        d, hd, F = self.input_dim, self.num_heads * self.head_dim, self.input_dim * self.fc_factor
        shapes = {"wq": (d, hd), "wk": (d, hd), "wv": (d, hd), "wo": (hd, d), "ln1_gamma": (d,), "ln1_beta": (d,),
                  "w1": (d, F), "b1": (F,), "w2": (F, d), "b2": (d,), "ln2_gamma": (d,), "ln2_beta": (d,)}
        ws = {}
        for n in _WEIGHT_NAMES:
            if n not in weights:
                raise ValueError(f"EncoderBlock.set_weights: missing '{n}'")
            t = torch.as_tensor(np.asarray(weights[n]) if not isinstance(weights[n], torch.Tensor) else weights[n], dtype=torch.float32)
            if tuple(t.shape) != shapes[n]:
                raise ValueError(f"EncoderBlock.set_weights: '{n}' has shape {tuple(t.shape)}, expected {shapes[n]}")
            ws[n] = t.to(device).contiguous()
        if device.type != "cuda":
            raise RuntimeError("telugu_asr_b200: EncoderBlock weights must live on a CUDA device (there is no CPU path)")
        self._release()
        self.weights, self._device = ws, device
        lib = _native.lib()
        W = _native.TasrEncoderBlockWeights(*[ws[n].data_ptr() for n in _WEIGHT_NAMES], self.input_dim, self.num_heads, self.head_dim,
                                            self.fc_factor, self.ln_eps)
        plan = C.c_void_p()
        with torch.cuda.device(device):
            _native.check(lib.tasr_encoder_block_plan_create(C.byref(W), C.byref(plan), _native.stream_ptr()))
        self._plan = plan

    def _release(self):
        if self._plan is not None:
            _native.lib().tasr_encoder_block_plan_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def prepare(self, t_max: int):
        """Size the RoPE table for sequences up to t_max (so that later calls only enqueue, e.g. inside a CUDA graph)."""
        _native.check(_native.lib().tasr_encoder_block_prepare(self._plan, int(t_max)))

    # ------------------------------------------------------------------ call
    def __call__(self, inputs, training: bool = False, use_causal_mask: bool = False, mask=None, lengths=None):
        """inputs [B, T, d] float32 CUDA; mask = the float / bool padding mask [B, T] of `lengths_to_padding_mask`
        (encoder.py:43-48: a prefix of ones per utterance) or `lengths` = the int32 lengths it was made from."""
        if training:
            raise NotImplementedError("EncoderBlock: inference only (training=True would apply dropout)")
        x = _native.require_cuda(inputs, "inputs")
        if x.dim() != 3 or x.shape[-1] != self.input_dim:
            raise ValueError(f"EncoderBlock: inputs must be [B, T, {self.input_dim}], got {tuple(x.shape)}")
        if self._plan is None:
            self.build(x.device)
        x = x.contiguous().float()
        B, T, d = x.shape
        if lengths is None and mask is not None:
            m = _native.require_cuda(mask, "mask")
            if tuple(m.shape) != (B, T):
                raise ValueError(f"EncoderBlock: mask must be [B, T] = {(B, T)}, got {tuple(m.shape)}")
            lengths = (m != 0).sum(dim=1).to(torch.int32)            # the reference's masks are prefixes (encoder.py:46-47)
        len_ptr = None
        if lengths is not None:
            lengths = _native.require_cuda(lengths, "lengths").to(torch.int32).contiguous()
            if lengths.numel() != B:
                raise ValueError("EncoderBlock: lengths must have one entry per utterance")
            len_ptr = lengths.data_ptr()
        lib = _native.lib()
        n_ws = lib.tasr_encoder_block_workspace_floats(self._plan, B, T)
        ws = _native.empty((max(int(n_ws), 4),), torch.float32, x.device)
        out = _native.empty((B, T, d), torch.float32, x.device)
        _native.check(lib.tasr_encoder_block_f32(self._plan, x.data_ptr(), len_ptr, B, T, 1 if use_causal_mask else 0,
                                                 ws.data_ptr(), out.data_ptr(), _native.stream_ptr()))
        return out

    def compute_output_shape(self, input_shape):
        return tuple(input_shape[:-1]) + (self.input_dim,)            # encoder.py:156-161

    def get_config(self):
        return {"input_dim": self.input_dim, "dropout": self.dropout, "activation": self.activation, "num_heads": self.num_heads,
                "head_dim": self.head_dim, "fc_factor": self.fc_factor, "name": self.name}
