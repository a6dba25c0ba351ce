"""The timed CPU baseline (oracle/torch_port.py) computes the same thing as the numpy oracle."""
import numpy as np
import torch

import oracle
from oracle import torch_port


def test_torch_port_matches_numpy_oracle():
    lens = [16000, 399, 4000, 24000]
    wav, ln = oracle.make_waveforms(lens, seed=8, dist="tilt")
    weights = oracle.glorot_subsampling_weights(192, 80, seed=7)
    out, mask, len3, feat, n = torch_port.frontend_torch(wav, ln, weights)
    rf, rn = oracle.logmel_batch_ref(wav, ln, dtype=np.float64)
    np.testing.assert_array_equal(n.numpy(), rn)
    assert np.abs(feat.numpy() - rf).max() < 1e-4
    ro, rm, rl = oracle.subsample_ref(rf.astype(np.float32), rn, weights, dtype=np.float64)
    np.testing.assert_array_equal(len3.numpy(), rl[-1])
    np.testing.assert_array_equal(mask.numpy(), rm)
    assert np.abs(out.numpy() - ro).max() / np.abs(ro).max() < 1e-4
