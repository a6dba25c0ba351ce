"""CPU tests that pin the ORACLE for the subsampling half of the path."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle.subsampling_ref import sepconv1d_ref


def test_conv_length_trunc_semantics(golden_dir):
    # float32 arithmetic then a truncating cast: differs from floor for short inputs
    assert int(oracle.conv_length_f32_trunc(6, 9, "valid", 2)) == 0      # -0.5 -> 0 (floor would be -1)
    assert int(oracle.conv_length_f32_trunc(8, 9, "valid", 2)) == 0      # 0.5 -> 0
    assert int(oracle.conv_length_f32_trunc(9, 9, "valid", 2)) == 1
    assert int(oracle.conv_length_f32_trunc(0, 9, "valid", 2)) == -3     # -3.5 -> -3
    assert int(oracle.conv_length_f32_trunc(7, 9, "same", 2)) == 4
    L = np.arange(9, 4000)
    np.testing.assert_array_equal(oracle.conv_length_f32_trunc(L, 9, "valid", 2), (L - 9) // 2 + 1)
    g = np.load(os.path.join(golden_dir, "lengths.npz"))
    np.testing.assert_array_equal(oracle.conv_lengths_ref(g["n_frames"]), g["conv_lengths"])
    np.testing.assert_array_equal(oracle.conv_lengths_ref(g["small_in"]), g["small_out"])
    # SURVEY.md §8 table
    assert oracle.conv_lengths_ref([998, 1498, 2998, 98]).T.tolist() == [[495, 244, 118], [745, 369, 181], [1495, 744, 368], [45, 19, 6]]


def test_padding_mask_width_is_max_length():
    m = oracle.lengths_to_padding_mask_ref([3, 0, 5, -2])
    assert m.shape == (4, 5) and m.dtype == np.float32
    assert m.tolist() == [[1, 1, 1, 0, 0], [0] * 5, [1] * 5, [0] * 5]
    assert oracle.lengths_to_padding_mask_ref([-1, -3]).shape == (2, 0)


def test_create_audio_mask():
    a = np.zeros((2, 4, 3, 1), dtype=np.float32)
    a[0, :2] = -9.0
    a[1, :3, 1] = 0.5
    m = oracle.create_audio_mask_ref(a)
    assert m.shape == (2, 4, 3)
    assert m[0].sum() == 6 and m[1].sum() == 3


@pytest.mark.parametrize("act", [None, "tanh", "gelu"])
def test_sepconv_matches_torch_conv1d(act):
    rng = np.random.default_rng(0)
    B, T, Cin, Cout, k = 2, 37, 12, 20, 9
    x = rng.standard_normal((B, T, Cin)).astype(np.float32)
    dw = rng.standard_normal((k, Cin)).astype(np.float32) * 0.3
    pw = rng.standard_normal((Cin, Cout)).astype(np.float32) * 0.3
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    y = sepconv1d_ref(x, dw, pw, b, stride=2, padding="valid", activation=act, dtype=np.float64)
    xt = torch.from_numpy(x).double().permute(0, 2, 1)
    d = torch.nn.functional.conv1d(xt, torch.from_numpy(dw).double().t().unsqueeze(1), stride=2, groups=Cin)
    z = torch.nn.functional.conv1d(d, torch.from_numpy(pw).double().t().unsqueeze(-1), torch.from_numpy(b).double())
    if act == "tanh":
        z = torch.tanh(z)
    elif act == "gelu":
        z = torch.nn.functional.gelu(z)  # exact erf form
    ref = z.permute(0, 2, 1).numpy()
    assert y.shape == ref.shape == (B, 15, Cout)
    np.testing.assert_allclose(y, ref, rtol=1e-10, atol=1e-12)


def test_subsample_stack_shapes_lengths_and_no_zeroing(golden_dir):
    g = np.load(os.path.join(golden_dir, "subsample_2x3s.npz"))
    weights = oracle.glorot_subsampling_weights(192, 80, seed=int(g["weight_seed"]))
    assert [w[0].shape for w in weights] == [(9, 80), (9, 192), (9, 384)]
    assert [w[1].shape for w in weights] == [(80, 192), (192, 384), (384, 192)]
    assert sum(w[0].size + w[1].size + w[2].size for w in weights) == 169488
    out, mask, len_all = oracle.subsample_ref(g["feat32"], g["n_frames"], weights, dtype=np.float64)
    np.testing.assert_allclose(out, g["out64"], rtol=0, atol=1e-10)
    np.testing.assert_array_equal(mask, g["mask"])
    np.testing.assert_array_equal(len_all, g["len_all"])
    assert out.shape == (2, 31, 192) and mask.shape == (2, int(len_all[-1].max()))
    # the reference does not zero padded positions: rows past len3 hold conv-over-zero values
    short = int(np.argmin(len_all[-1]))
    assert np.abs(out[short, len_all[-1][short]:]).max() > 0
    # mask from a [B,T,F] audio mask gives the same lengths as passing n_frames
    am = oracle.create_audio_mask_ref(g["feat32"])
    _, mask2, len2 = oracle.subsample_ref(g["feat32"], am, weights, dtype=np.float32)
    np.testing.assert_array_equal(len2, len_all)
    np.testing.assert_array_equal(mask2, mask)


def test_valid_positions_depend_only_on_valid_frames():
    """t < L_out only reads frames < L (2t+8 <= L-1), so a longer zero padding changes nothing there."""
    wav, ln = oracle.make_waveforms([20000], seed=9)
    f, nf = oracle.logmel_batch_ref(wav, ln)
    weights = oracle.glorot_subsampling_weights(32, 80, seed=1)
    a, _, la = oracle.subsample_ref(f, nf, weights)
    f_pad = np.concatenate([f, np.zeros((1, 40, 80, 1), np.float32)], axis=1)
    b, _, lb = oracle.subsample_ref(f_pad, nf, weights)
    np.testing.assert_array_equal(la, lb)
    L3 = int(la[-1][0])
    np.testing.assert_array_equal(a[0, :L3], b[0, :L3])


def test_float32_band_subsampling():
    wav, ln = oracle.make_waveforms([48000, 30000], seed=5)
    f, nf = oracle.logmel_batch_ref(wav, ln)
    weights = oracle.glorot_subsampling_weights(192, 80, seed=7)
    a, _, _ = oracle.subsample_ref(f, nf, weights, dtype=np.float32)
    b, _, _ = oracle.subsample_ref(f, nf, weights, dtype=np.float64)
    assert np.abs(a - b).max() / np.abs(b).max() < 1e-5
