"""CPU restatement of the reference Conv1D subsampling stack.  TEST INFRASTRUCTURE
(see oracle/__init__.py).  PARITY UNPINNED by the reference (no tests / vectors).

Follows:
  src/models/moonshine/encoder.py:11-41   layer construction (filters, k, s, padding, activations)
  src/models/moonshine/encoder.py:43-48   lengths_to_padding_mask
  src/models/moonshine/encoder.py:50-71   call (squeeze, mask->lengths, 3x SeparableConv1D, lengths)
  src/utils/math_util.py:20-32            get_conv_length (float32 arithmetic, cast truncates)
  src/models/moonshine/model.py:73-82     create_masks (audio_mask = any(bin != 0.0))
and Keras 2.15 SeparableConv1D: depthwise cross-correlation (depth_multiplier 1,
no bias) -> 1x1 pointwise -> bias_add -> activation; "gelu" is the exact erf form.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
    "conv_length_f32_trunc",
    "conv_lengths_ref",
    "lengths_to_padding_mask_ref",
    "create_audio_mask_ref",
    "sepconv1d_ref",
    "subsample_ref",
    "glorot_subsampling_weights",
    "ACTIVATIONS",
]

# Effective defaults of the trained reference: the YAML key is `activation`, the layer
# reads `activations` (encoder.py:25 vs config/model.yaml:27) -> tanh, gelu, gelu.
DEFAULT_ACTIVATIONS = ("tanh", "gelu", "gelu")

try:  # scipy is test-side only
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


def _act(name, x, dtype):
    if name in (None, "linear", "none"):
        return x
    if name == "tanh":
        return np.tanh(x).astype(dtype)
    if name == "gelu":  # keras.activations.gelu(approximate=False)
        return (dtype(0.5) * x * (dtype(1.0) + _erf(x / np.sqrt(dtype(2.0))))).astype(dtype)
    if name == "relu":
        return np.maximum(x, dtype(0.0))
    raise ValueError(f"unsupported activation {name!r}")


ACTIVATIONS = ("linear", "tanh", "gelu", "relu")


def conv_length_f32_trunc(length, kernel_size: int, padding: str, stride: int):
    """math_util.get_conv_length (src/utils/math_util.py:20-32): all operands cast to
    float32, result cast to int32 (truncation toward zero, NOT floor: L=6,k=9,s=2 -> 0)."""
    length = np.asarray(length).astype(np.float32)
    k = np.float32(kernel_size)
    s = np.float32(stride)
    if padding == "same":
        length = np.ceil(length / s)
    elif padding == "valid":
        length = (length - k) / s + np.float32(1.0)
    return np.trunc(length).astype(np.int32)


def conv_lengths_ref(lengths, kernel_size=(9, 9, 9), strides=(2, 2, 2), padding=("valid",) * 3):
    """Lengths after every layer: returns [n_layers, B] int32 (encoder.py:60-68)."""
    out = []
    cur = np.asarray(lengths).astype(np.int32)
    for k, s, pad in zip(kernel_size, strides, padding):
        cur = conv_length_f32_trunc(cur, k, pad, s)
        out.append(cur)
    return np.stack(out, axis=0)


def lengths_to_padding_mask_ref(lengths) -> np.ndarray:
    """encoder.py:43-48: width = max(lengths) (not the conv output length), float32 0/1."""
    lengths = np.asarray(lengths).astype(np.int32)
    max_len = int(lengths.max()) if lengths.size else 0
    max_len = max(max_len, 0)
    return (np.arange(max_len)[None, :] < lengths[:, None]).astype(np.float32)


def create_audio_mask_ref(audio_inputs: np.ndarray, pad_value: float = 0.0) -> np.ndarray:
    """model.py:80: reduce_any(audio != pad, axis=-1) on [B,T,F,1] -> [B,T,F] float32."""
    return np.any(audio_inputs != pad_value, axis=-1).astype(np.float32)


def _same_pad(T: int, k: int, s: int) -> tuple[int, int]:
    out = -(-T // s)
    total = max((out - 1) * s + k - T, 0)
    return total // 2, total - total // 2


def sepconv1d_ref(x: np.ndarray, dw: np.ndarray, pw: np.ndarray, bias: np.ndarray,
                  stride: int = 2, padding: str = "valid", activation: str | None = None,
                  dtype=np.float32) -> np.ndarray:
    """Keras SeparableConv1D forward.  x [B,T,Cin]; dw [k,Cin] (Keras depthwise_kernel
    (k,Cin,1) squeezed); pw [Cin,Cout] (pointwise_kernel (1,Cin,Cout) squeezed); bias [Cout].
    y[b,t,c] = sum_k x[b, s*t+k, c]*dw[k,c];  z = y @ pw + bias;  act(z)."""
    dtype = np.dtype(dtype).type
    x = x.astype(dtype)
    dw = dw.astype(dtype)
    pw = pw.astype(dtype)
    bias = bias.astype(dtype)
    B, T, Cin = x.shape
    k = dw.shape[0]
    if padding == "same":
        lo, hi = _same_pad(T, k, stride)
        x = np.pad(x, [[0, 0], [lo, hi], [0, 0]])
        T = x.shape[1]
    T_out = max(0, (T - k) // stride + 1)
    y = np.zeros((B, T_out, Cin), dtype=dtype)
    for j in range(k):
        y += x[:, j: j + stride * (T_out - 1) + 1: stride, :][:, :T_out] * dw[j][None, None, :]
    z = (y.reshape(B * T_out, Cin) @ pw).reshape(B, T_out, -1).astype(dtype) + bias[None, None, :]
    return _act(activation, z.astype(dtype), dtype)


def subsample_ref(feat: np.ndarray, mask_or_lengths, weights, activations=DEFAULT_ACTIVATIONS,
                  kernel_size=(9, 9, 9), strides=(2, 2, 2), padding=("valid",) * 3, dtype=np.float32):
    """Conv1DSubsamplingLayer.call (encoder.py:50-71).
    feat [B,T,F,1]; mask_or_lengths: None, a [B,T,F] / [B,T] mask (as create_masks gives) or
    int lengths[B]; weights: list of (dw[k,Cin], pw[Cin,Cout], bias[Cout]).
    Returns (out [B,T3,C], padding_mask [B,max(len3)] or None, lengths [n_layers,B] or None).
    No zeroing between layers: padded positions hold conv-over-zero values, like the reference."""
    h = np.squeeze(feat, axis=-1)
    lengths = None
    if mask_or_lengths is not None:
        m = np.asarray(mask_or_lengths)
        if m.ndim == 1:
            lengths = m.astype(np.int32)
        else:
            m = m.astype(np.int32)
            if m.ndim == 3:
                m = m.max(axis=-1)               # encoder.py:55
            lengths = m.sum(axis=1).astype(np.int32)  # :56
    all_len = []
    for (dw, pw, b), act, k, s, pad in zip(weights, activations, kernel_size, strides, padding):
        assert dw.shape[0] == k
        h = sepconv1d_ref(h, dw, pw, b, stride=s, padding=pad, activation=act, dtype=dtype)
        if lengths is not None:
            lengths = conv_length_f32_trunc(lengths, k, pad, s)
            all_len.append(lengths)
    mask = lengths_to_padding_mask_ref(lengths) if lengths is not None else None
    return h, mask, (np.stack(all_len, 0) if all_len else None)


def glorot_subsampling_weights(model_dim: int = 192, in_dim: int = 80, kernel_size=(9, 9, 9),
                               seed: int = 7, bias_range: float = 0.1):
    """Synthetic weights of the reference's shapes (encoder.py:21; no checkpoint ships):
    Keras glorot-uniform limits — depthwise (k,Cin,1): sqrt(6/(k*Cin + k*1)) ... Keras computes
    fans for a (k, Cin, 1) kernel as fan_in = k*Cin, fan_out = k*1; pointwise (1,Cin,Cout):
    fan_in = Cin, fan_out = Cout.  Bias U(-bias_range, bias_range) instead of zeros so the bias
    path is exercised."""
    rng = np.random.default_rng(seed)
    filters = [model_dim, 2 * model_dim, model_dim]
    cin = in_dim
    out = []
    for k, cout in zip(kernel_size, filters):
        lim_dw = math.sqrt(6.0 / (k * cin + k * 1))
        lim_pw = math.sqrt(6.0 / (cin + cout))
        dw = rng.uniform(-lim_dw, lim_dw, size=(k, cin)).astype(np.float32)
        pw = rng.uniform(-lim_pw, lim_pw, size=(cin, cout)).astype(np.float32)
        b = rng.uniform(-bias_range, bias_range, size=(cout,)).astype(np.float32)
        out.append((dw, pw, b))
        cin = cout
    return out
