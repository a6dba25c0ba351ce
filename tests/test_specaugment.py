"""SpecAugment: host mirror of src/augmentations/{augmentation,specaugment}.py on the
CPU (constructor parity, registry error, draw bounds) and, on the GPU, the mask arithmetic bit for bit
against the oracle restatement for given masks."""
import numpy as np
import pytest
import torch

import oracle
import telugu_asr_b200 as tasr
from oracle import specaugment_ref as sref


def test_augmentation_parse_mirrors_reference():
    a = tasr.Augmentation({"prob": 0.3, "feature_augment": {"time_masking": {"num_masks": 2, "mask_factor": 50, "p_upperbound": 0.5},
                                                            "freq_masking": None}})
    assert a.prob == 0.3 and len(a.feature_augmentations) == 2 and a.signal_augmentations == []
    t, f = a.feature_augmentations
    assert (t.num_masks, t.mask_factor, t.p_upperbound) == (2, 50, 0.5)
    assert (f.num_masks, f.mask_factor) == (1, 27)                       # specaugment.py:9-10 defaults
    assert tasr.Augmentation(None).prob == 0.5 and tasr.Augmentation(None).feature_augmentations == []
    with pytest.raises(KeyError, match="No tf augmentation named: pitch_shift"):
        tasr.Augmentation({"feature_augment": {"pitch_shift": None}})


def test_draws_respect_the_reference_bounds():
    a = tasr.Augmentation({"prob": 1.0, "feature_augment": {"time_masking": {"num_masks": 2, "mask_factor": 100, "p_upperbound": 0.2},
                                                            "freq_masking": {"num_masks": 2, "mask_factor": 27}}}, seed=0)
    nf = np.array([998, 98, 5, 0, 1498] * 40)
    tm, fm = a.draw_masks(nf, 80)
    assert tm.shape == (200, 2, 2) and fm.shape == (200, 2, 2)
    T = nf[:, None]
    assert (tm[..., 1] <= np.minimum(99, (T * 0.2).astype(int))).all() and (tm[..., 0] + tm[..., 1] <= T).all()
    assert (fm[..., 1] <= 26).all() and (fm[..., 0] + fm[..., 1] <= 80).all() and (fm >= 0).all() and (tm >= 0).all()
    assert tm[nf == 0].sum() == 0                                         # nothing to mask in an empty utterance
    assert 5 < tm[nf == 998][..., 1].mean() < 95 and 5 < fm[..., 1].mean() < 20
    # prob: about half of the augmentations are skipped (width-0 masks)
    b = tasr.Augmentation({"prob": 0.5, "feature_augment": {"freq_masking": {"mask_factor": 27}}}, seed=1)
    _, fm2 = b.draw_masks(np.full(2000, 500), 80)
    frac = (fm2[..., 1] > 0).mean()
    assert 0.40 < frac < 0.55                                             # 0.5 * P(f > 0) = 0.5 * 26/27
    # the oracle's scalar draws obey the same bounds
    rng = np.random.default_rng(0)
    for _ in range(200):
        f0, f = sref.draw_freq_mask(rng, 80)
        t0, t = sref.draw_time_mask(rng, 120, 100, 0.5)
        assert 0 <= f <= 26 and f0 + f <= 80 and 0 <= t <= 60 and t0 + t <= 120


def test_oracle_masks_are_the_concat_masks():
    x = -np.arange(1, 4 * 6 + 1, dtype=np.float32).reshape(4, 6, 1)
    y = sref.freq_mask_ref(x, 2, 3)
    assert (y[:, :2] == x[:, :2]).all() and (y[:, 5:] == x[:, 5:]).all() and (y[:, 2:5] == 0).all()
    assert np.signbit(y[:, 2:5]).all()                                    # negative * 0.0 = -0.0, like tf multiply
    z = sref.time_mask_ref(x, 1, 2)
    assert (z[1:3] == 0).all() and (z[0] == x[0]).all() and (z[3] == x[3]).all()
    assert np.array_equal(sref.time_mask_ref(x, 0, 0), x)


@pytest.mark.gpu
def test_specaugment_kernel_bit_exact(cuda_device):
    lens = [16000, 48000, 399, 8000, 24000]
    wav, ln = oracle.make_waveforms(lens, seed=5, dist="tilt")
    feat = tasr.SpeechFeaturizer(**tasr.REFERENCE_SPEECH_CONFIG)
    x, nf = feat(torch.from_numpy(wav).to(cuda_device), torch.from_numpy(ln).to(cuda_device))
    base = x.cpu().numpy()
    a = tasr.Augmentation({"prob": 0.7, "feature_augment": {"time_masking": {"num_masks": 2, "mask_factor": 60},
                                                            "freq_masking": {"num_masks": 2, "mask_factor": 27}}}, seed=3)
    nfh = nf.cpu().numpy()
    tm, fm = a.draw_masks(nfh, 80)
    tm[0] = [[10, 20], [25, 0]]                                            # one applied, one "not applied"
    fm[1] = [[0, 5], [78, 2]]                                              # masks touching both edges
    tm[1] = [[int(nfh[1]) - 7, 7], [0, 3]]
    got = tasr.Augmentation.apply_masks(x.clone(), nf, tm, fm).cpu().numpy()
    ref = sref.apply_batch_ref(base, nfh, tm, fm)
    assert np.array_equal(got.view(np.int32), ref.view(np.int32))          # bit for bit, including -0.0
    assert got[2].sum() == 0 and nfh[2] == 0                               # empty utterance untouched
    # end to end through feature_augment: only masked rows/columns differ from the input, padding stays 0.0
    y = a.feature_augment(x.clone(), nf, n_frames_host=nfh).cpu().numpy()
    for b, T in enumerate(nfh):
        assert not y[b, T:].any()
        changed = y[b, :T] != base[b, :T]
        assert (y[b, :T][changed] == 0).all()
    # no feature augmentations configured -> identity (same tensor back)
    assert tasr.Augmentation(None).feature_augment(x, nf) is x
