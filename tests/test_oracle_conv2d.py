"""CPU tests of the Conv2dSubsampling oracle (oracle/conv2d_subsampling_ref.py) against torch's conv2d with the
TensorFlow "SAME" padding written out explicitly, and of the shape / length rules of
src/models/conformer/encoder.py:50-67."""
import numpy as np
import pytest
import torch

import oracle


@pytest.mark.parametrize("T", [1, 2, 7, 8, 37, 100])
def test_same_pads_and_shapes(T):
    out, before, after = oracle.same_pads(T, 3, 2)
    assert out == -(-T // 2) and before + after == max((out - 1) * 2 + 3 - T, 0) and before <= after


@pytest.mark.parametrize("T,F", [(37, 80), (38, 80), (5, 6), (1, 80)])
def test_conv2d_subsample_vs_torch(T, F):
    rng = np.random.default_rng(T)
    C = 16
    ws = oracle.glorot_conv2d_weights(C, seed=3)
    x = rng.standard_normal((3, T, F, 1)).astype(np.float32)
    x[1, T // 2:] = 0.0                                  # a zero-padded utterance
    lens = np.array([T, T // 2, max(T - 1, 0)], dtype=np.int32)
    out, out_len = oracle.conv2d_subsample_ref(x, lens, ws, dtype=np.float64)
    h = torch.from_numpy(x.astype(np.float64)).permute(0, 3, 1, 2)      # NCHW
    for (w, b) in ws:
        _, pt, pb = oracle.same_pads(h.shape[2], 3, 2)
        _, pl, pr = oracle.same_pads(h.shape[3], 3, 2)
        h = torch.nn.functional.pad(h, (pl, pr, pt, pb))
        h = torch.relu(torch.nn.functional.conv2d(h, torch.from_numpy(w.astype(np.float64)).permute(3, 2, 0, 1),
                                                  torch.from_numpy(b.astype(np.float64)), stride=2))
    want = h.permute(0, 2, 3, 1).reshape(3, h.shape[2], -1).numpy()     # [B, T', F'*C], channel fastest
    assert out.shape == want.shape == (3, -(-(-(-T // 2)) // 2), -(-(-(-F // 2)) // 2) * C)
    np.testing.assert_allclose(out, want, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(out_len, -(-lens // 2))             # get_conv_length applied once (encoder.py:59-64)
