"""CPU restatement of the reference's first encoder block.  TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED by the reference (no tests / vectors; TensorFlow absent); pinned instead against torch
(scaled_dot_product_attention, layer_norm, gelu) in tests/test_oracle_encoder_block.py.

Follows:
  src/models/moonshine/encoder.py:109-154      EncoderBlock: MHSAModule then FFNModule
  src/models/layers/attention.py:519-602       MHSAModule.call: mha(x, x, x, attention_mask=mask) -> dropout -> x + . -> LayerNorm
  src/models/layers/attention.py:44-230        MultiHeadAttention: bias-free q/k/v/projection kernels, RoPE on q and k,
                                               q / sqrt(head_dim), Keras masked softmax (scores + (1 - mask) * -1e9), @ v, projection
  src/models/layers/positional_encoding.py:20-93  RoPEPositionalEncoding: rot_dim = max(head_dim // 2, 32), inv_freq over rot_dim,
                                               interleaved pairs (x1, x2) -> (x1 cos - x2 sin, x2 cos + x1 sin),
                                               output = concat([unrotated, rotated])  (the unrotated tail comes FIRST)
  src/models/layers/mlp.py:9-60                FFNModule: dense1 (gelu, exact erf) -> dropout -> dense2 -> + inputs -> LayerNorm
  tf.keras.layers.LayerNormalization defaults  axis -1, epsilon 1e-3, (x - mean) * rsqrt(var + eps) * gamma + beta

How `inputs, pos = inputs` (attention.py:572) is resolved: EncoderBlock.call (encoder.py:152) hands MHSAModule a single tensor, and
unpacking a [B, T, d] tensor into two names only succeeds for B == 2 (and then means "utterance 0, utterance 1").  `pos` is only
read by the 'relmha' attention type; the encoder builds 'sdpa'.  The restatement takes the evident intent: `inputs` is the whole
tensor and `pos` is unused.

Padded query rows (t >= length): every score of the row gets -1e9, which in float32 swallows the score itself (|s| < 32), so
the softmax is uniform over ALL T keys and the row attends to the plain mean of v.  The restatement reproduces that.
Dropout is inference-mode identity.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = ["rope_tables", "rope_apply", "encoder_block_ref", "glorot_encoder_block_weights"]


def rope_tables(seq_len: int, head_dim: int, base: float = 10000.0, dtype=np.float32):
    """cos / sin [T, rot_dim] of positional_encoding.py:33-60 (float32 arithmetic like the reference when dtype is float32)."""
    rot_dim = max(head_dim // 2, 32)
    f32 = np.float32
    index = np.arange(0, rot_dim, 2, dtype=f32)
    inv_freq = (f32(1.0) / np.power(f32(base), index / f32(rot_dim))).astype(f32)          # InvFreqInitializer
    pos = np.arange(seq_len, dtype=f32)[:, None]
    freq = (pos * inv_freq[None, :]).astype(f32)                                            # [T, rot/2]
    freq = np.repeat(freq, 2, axis=1)                                                       # stack([f, f], -1).reshape -> interleaved
    freq = freq.astype(dtype)
    return np.cos(freq).astype(dtype), np.sin(freq).astype(dtype), rot_dim


def rope_apply(x: np.ndarray, cos: np.ndarray, sin: np.ndarray, rot_dim: int) -> np.ndarray:
    """x [B, T, H, Dh] -> positional_encoding.py:74-90."""
    if rot_dim > x.shape[-1]:
        raise ValueError("rot_dim = max(head_dim // 2, 32) exceeds head_dim: the reference's slice/multiply would not broadcast")
    t_rot, t_unrot = x[..., :rot_dim], x[..., rot_dim:]
    x1, x2 = t_rot[..., 0::2], t_rot[..., 1::2]
    half = np.empty_like(t_rot)
    half[..., 0::2] = -x2
    half[..., 1::2] = x1
    rot = t_rot * cos[None, :, None, :] + half * sin[None, :, None, :]
    return np.concatenate([t_unrot, rot], axis=-1)


def _layer_norm(x, gamma, beta, eps):
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    return (x - mean) / np.sqrt(var + eps) * gamma + beta


def _gelu_erf(x):
    from scipy.special import erf
    return 0.5 * x * (1.0 + erf(x / math.sqrt(2.0)))


def encoder_block_ref(x: np.ndarray, lengths, w: dict, num_heads: int, head_dim: int, dtype=np.float64,
                      use_causal_mask: bool = False, ln_eps: float = 1e-3) -> np.ndarray:
    """x [B, T, d]; lengths [B] (valid prefix of every utterance = the padding mask of encoder.py:43-48) or None.
    w: wq, wk, wv [d, H*Dh], wo [H*Dh, d], ln1_gamma, ln1_beta [d], w1 [d, F], b1 [F], w2 [F, d], b2 [d], ln2_gamma, ln2_beta."""
    x = np.asarray(x, dtype=dtype)
    B, T, d = x.shape
    H, Dh = num_heads, head_dim
    W = {k: np.asarray(v, dtype=dtype) for k, v in w.items()}
    q = (x @ W["wq"]).reshape(B, T, H, Dh)
    k = (x @ W["wk"]).reshape(B, T, H, Dh)
    v = (x @ W["wv"]).reshape(B, T, H, Dh)
    cos, sin, rot = rope_tables(T, Dh, dtype=np.float32)
    cos, sin = cos.astype(dtype), sin.astype(dtype)
    q = rope_apply(q, cos, sin, rot).transpose(0, 2, 1, 3)          # [B, H, T, Dh]
    k = rope_apply(k, cos, sin, rot).transpose(0, 2, 1, 3)
    v = v.transpose(0, 2, 1, 3)
    q = q * dtype(1.0 / math.sqrt(Dh))
    scores = q @ k.transpose(0, 1, 3, 2)                            # [B, H, T, T]
    if lengths is not None or use_causal_mask:
        m = np.ones((B, T, T), dtype=bool)
        if lengths is not None:
            valid = np.arange(T)[None, :] < np.asarray(lengths)[:, None]          # [B, T]
            m &= valid[:, :, None] & valid[:, None, :]                             # query & value & key masks
        if use_causal_mask:
            m &= np.tril(np.ones((T, T), dtype=bool))[None]
        # Keras Softmax(mask): inputs += (1 - mask) * -1e9, evaluated in the reference's float32
        add = np.where(m, 0.0, -1e9)[:, None, :, :]
        if dtype == np.float32:
            scores = (scores + add.astype(np.float32)).astype(np.float32)
        else:
            # float64 evaluation of what the float32 graph does: a fully masked row is uniform, masked keys weigh exactly 0
            scores = np.where(m[:, None], scores, -np.inf)
            full = ~m.any(axis=-1)                                                  # [B, T] rows with no admissible key
            scores = np.where(full[:, None, :, None], 0.0, scores)
    scores = scores - scores.max(axis=-1, keepdims=True)
    p = np.exp(scores)
    p = p / p.sum(axis=-1, keepdims=True)
    att = (p @ v).transpose(0, 2, 1, 3).reshape(B, T, H * Dh) @ W["wo"]
    h1 = _layer_norm(x + att, W["ln1_gamma"], W["ln1_beta"], dtype(ln_eps))
    f = _gelu_erf(h1 @ W["w1"] + W["b1"]).astype(dtype)
    out = _layer_norm(f @ W["w2"] + W["b2"] + h1, W["ln2_gamma"], W["ln2_beta"], dtype(ln_eps))
    return out.astype(dtype)


def glorot_encoder_block_weights(d_model: int = 192, num_heads: int = 6, head_dim: int = 32, fc_factor: int = 1, seed: int = 11):
    """Random weights with the reference's initialisers (glorot-uniform kernels; biases and LayerNorm parameters perturbed
    away from their zeros / ones defaults so that a test exercises them)."""
    rng = np.random.default_rng(seed)
    hd, F = num_heads * head_dim, d_model * fc_factor

    def glorot(i, o):
        lim = math.sqrt(6.0 / (i + o))
        return rng.uniform(-lim, lim, size=(i, o)).astype(np.float32)

    return {
        "wq": glorot(d_model, hd), "wk": glorot(d_model, hd), "wv": glorot(d_model, hd), "wo": glorot(hd, d_model),
        "ln1_gamma": (1.0 + 0.1 * rng.standard_normal(d_model)).astype(np.float32),
        "ln1_beta": (0.1 * rng.standard_normal(d_model)).astype(np.float32),
        "w1": glorot(d_model, F), "b1": rng.uniform(-0.1, 0.1, F).astype(np.float32),
        "w2": glorot(F, d_model), "b2": rng.uniform(-0.1, 0.1, d_model).astype(np.float32),
        "ln2_gamma": (1.0 + 0.1 * rng.standard_normal(d_model)).astype(np.float32),
        "ln2_beta": (0.1 * rng.standard_normal(d_model)).astype(np.float32),
    }
