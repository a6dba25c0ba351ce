"""Pins oracle/encoder_block_ref.py (src/models/moonshine/encoder.py:109-154 and the layers it calls) against
an independent torch evaluation: scaled_dot_product_attention with the reference's combined query/key mask, torch layer_norm,
torch gelu (exact erf), and a RoPE written from the rotation-matrix definition rather than the reference's rotate_half form."""
import math

import numpy as np
import torch
import torch.nn.functional as F

import oracle


def torch_block(x, lengths, w, H, Dh, causal=False, eps=1e-3):
    x = torch.from_numpy(x).double()
    B, T, d = x.shape
    W = {k: torch.from_numpy(v).double() for k, v in w.items()}
    q = (x @ W["wq"]).view(B, T, H, Dh)
    k = (x @ W["wk"]).view(B, T, H, Dh)
    v = (x @ W["wv"]).view(B, T, H, Dh)
    # RoPE as 2x2 rotations of consecutive pairs by angle t * 10000^(-2i/rot), rot = max(Dh // 2, 32)
    rot = max(Dh // 2, 32)
    inv = torch.from_numpy((np.float32(1.0) / np.power(np.float32(10000.0), np.arange(0, rot, 2, dtype=np.float32) / np.float32(rot))).astype(np.float32))
    ang = (torch.arange(T, dtype=torch.float32)[:, None] * inv[None, :]).double()           # the reference forms the angle in float32

    def rope(t):
        a, b = t[..., 0:rot:2], t[..., 1:rot:2]
        c, s = torch.cos(ang)[None, :, None, :], torch.sin(ang)[None, :, None, :]
        r = torch.stack([a * c - b * s, b * c + a * s], dim=-1).reshape(B, T, H, rot)
        return torch.cat([t[..., rot:], r], dim=-1)

    q, k = rope(q).transpose(1, 2), rope(k).transpose(1, 2)
    v = v.transpose(1, 2)
    m = torch.ones(B, T, T, dtype=torch.bool)
    if lengths is not None:
        valid = torch.arange(T)[None, :] < torch.as_tensor(lengths)[:, None]
        m &= valid[:, :, None] & valid[:, None, :]
    if causal:
        m &= torch.tril(torch.ones(T, T, dtype=torch.bool))[None]
    full = ~m.any(-1)                     # fully masked (padded) query rows: the float32 -1e9 trick leaves a uniform softmax
    m = m | full[:, :, None]              # ... over ALL keys: admit every key and zero the query (all scores equal)
    q = q * (~full)[:, None, :, None]
    att = F.scaled_dot_product_attention(q, k, v, attn_mask=m[:, None])
    att = att.transpose(1, 2).reshape(B, T, H * Dh) @ W["wo"]
    h1 = F.layer_norm(x + att, (d,), W["ln1_gamma"], W["ln1_beta"], eps)
    f = F.gelu(h1 @ W["w1"] + W["b1"])
    return F.layer_norm(f @ W["w2"] + W["b2"] + h1, (d,), W["ln2_gamma"], W["ln2_beta"], eps).numpy()


def test_rope_tables_match_the_reference_formula():
    cos, sin, rot = oracle.rope_tables(7, 32)
    assert rot == 32 and cos.shape == (7, 32)
    assert np.array_equal(cos[:, 0::2], cos[:, 1::2])                    # interleaved pairs share their angle
    assert np.allclose(cos[3, 0], math.cos(3.0)) and np.allclose(sin[3, 2], math.sin(3.0 / 10000 ** (2 / 32)), atol=1e-6)
    assert oracle.rope_tables(5, 64)[2] == 32 and oracle.rope_tables(5, 36)[2] == 32 and oracle.rope_tables(5, 128)[2] == 64


def test_rope_puts_the_unrotated_tail_first():
    x = np.arange(2 * 3 * 1 * 36, dtype=np.float64).reshape(2, 3, 1, 36)
    cos, sin, rot = oracle.rope_tables(3, 36, dtype=np.float64)
    y = oracle.rope_apply(x, cos, sin, rot)
    assert np.array_equal(y[..., :4], x[..., 32:])                       # positional_encoding.py:90 concat([unrotated, rotated])
    assert np.allclose(y[:, 0, :, 4:], x[:, 0, :, :32])                  # position 0 is the identity rotation


def test_block_against_torch():
    rng = np.random.default_rng(0)
    for (B, T, H, Dh, fc, lens, causal) in [(3, 37, 6, 32, 1, [37, 20, 1], False), (2, 16, 6, 32, 2, None, False),
                                            (2, 25, 8, 36, 1, [25, 9], False), (2, 19, 6, 32, 1, [19, 12], True)]:
        d = H * Dh
        w = oracle.glorot_encoder_block_weights(d, H, Dh, fc, seed=5)
        x = rng.standard_normal((B, T, d)).astype(np.float32)
        ref = torch_block(x, lens, w, H, Dh, causal)
        got = oracle.encoder_block_ref(x, lens, w, H, Dh, dtype=np.float64, use_causal_mask=causal)
        assert got.shape == (B, T, d)
        assert np.abs(got - ref).max() < 1e-6, (B, T, np.abs(got - ref).max())
        got32 = oracle.encoder_block_ref(x, lens, w, H, Dh, dtype=np.float32, use_causal_mask=causal)
        L = np.asarray(lens if lens is not None else [T] * B)
        for b in range(B):                                                # float32 op-for-op evaluation: valid rows agree
            assert np.abs(got32[b, :L[b]] - ref[b, :L[b]]).max() < 2e-5
        if lens is not None and not causal:                               # and the -1e9 trick really gives the uniform rows
            for b in range(B):
                if L[b] < T:
                    assert np.abs(got32[b, L[b]:] - ref[b, L[b]:]).max() < 2e-5
