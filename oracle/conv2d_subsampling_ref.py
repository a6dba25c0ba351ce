"""CPU restatement of the conformer configuration's Conv2dSubsampling.  TEST INFRASTRUCTURE
(see oracle/__init__.py).  PARITY UNPINNED by the reference (no tests / vectors).

Follows:
  src/models/conformer/encoder.py:9-48    two tf.keras.layers.Conv2D(filters, 3, strides 2, padding "same")
  src/models/conformer/encoder.py:50-67   call: conv1 -> relu -> conv2 -> relu; lengths through get_conv_length ONCE
                                          (conv1's kernel/stride, "same": ceil(L/2), although time shrinks by 4);
                                          merge_two_last_dims
  src/utils/math_util.py:34-46            merge_two_last_dims: [B, T', F', C] -> [B, T', F'*C]
  config/conformer.yaml:22-27             filters 144, kernel_size 3, strides 2, padding same
and Keras 2.15 Conv2D (channels_last, kernel [kh, kw, Cin, Cout], cross-correlation, bias_add) with TensorFlow's
"SAME" rule: out = ceil(in / stride), pad_total = max((out - 1) * stride + k - in, 0), pad_before = pad_total // 2
(the extra row / column goes AFTER).
"""
from __future__ import annotations

import numpy as np

from .subsampling_ref import conv_length_f32_trunc

__all__ = ["same_pads", "conv2d_same_ref", "conv2d_subsample_ref", "glorot_conv2d_weights"]


def same_pads(n: int, k: int, s: int):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


def conv2d_same_ref(x, w, b, stride: int = 2, dtype=np.float32):
    """x [B, H, W, Cin], w [kh, kw, Cin, Cout], b [Cout] -> relu-less conv output [B, H', W', Cout]."""
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    B, H, W, Cin = x.shape
    kh, kw, _, Cout = w.shape
    Ho, pt, pb = same_pads(H, kh, stride)
    Wo, pl, pr = same_pads(W, kw, stride)
    xp = np.zeros((B, H + pt + pb, W + pl + pr, Cin), dtype=dtype)
    xp[:, pt:pt + H, pl:pl + W] = x
    out = np.zeros((B, Ho, Wo, Cout), dtype=dtype)
    for di in range(kh):
        for dj in range(kw):
            patch = xp[:, di:di + stride * (Ho - 1) + 1:stride, dj:dj + stride * (Wo - 1) + 1:stride]
            out += (patch.reshape(-1, Cin) @ w[di, dj]).reshape(B, Ho, Wo, Cout)
    return out + np.asarray(b, dtype=dtype)


def conv2d_subsample_ref(feat, lengths, weights, dtype=np.float32):
    """feat [B, T, F, 1], lengths [B], weights [(w1 [3,3,1,C], b1 [C]), (w2 [3,3,C,C], b2 [C])] ->
    (outputs [B, ceil(ceil(T/2)/2), ceil(ceil(F/2)/2) * C], lengths' [B] = ceil(lengths / 2))."""
    (w1, b1), (w2, b2) = weights
    h = np.maximum(conv2d_same_ref(feat, w1, b1, 2, dtype), dtype(0))
    h = np.maximum(conv2d_same_ref(h, w2, b2, 2, dtype), dtype(0))
    B, T2, F2, C = h.shape
    out_len = conv_length_f32_trunc(np.asarray(lengths), w1.shape[0], "same", 2)
    return h.reshape(B, T2, F2 * C), out_len


def glorot_conv2d_weights(filters: int = 144, seed: int = 7, bias_scale: float = 0.1):
    """Keras default glorot_uniform for both kernels ([3,3,1,C] and [3,3,C,C]); biases U(-bias_scale, bias_scale)
    instead of zeros so that the bias path is exercised."""
    rng = np.random.default_rng(seed)
    out = []
    for cin in (1, filters):
        lim = np.sqrt(6.0 / (9 * cin + 9 * filters))
        w = rng.uniform(-lim, lim, size=(3, 3, cin, filters)).astype(np.float32)
        b = rng.uniform(-bias_scale, bias_scale, size=(filters,)).astype(np.float32)
        out.append((w, b))
    return out
