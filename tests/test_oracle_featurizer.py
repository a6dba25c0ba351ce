"""CPU tests that pin the ORACLE for the featurizer half of the path.

The reference has no tests or vectors, so the oracle is pinned by closed forms,
independent library implementations and the committed fixtures."""
import os

import numpy as np
import pytest

import oracle
from oracle.featurizer_ref import FeatParams, yaml_params, _stft_power, featurize_ref


def test_constructor_arithmetic():
    p = yaml_params()
    assert (p.frame_length, p.frame_step, p.fft_length, p.num_spectrogram_bins) == (400, 160, 512, 257)
    with pytest.raises(AssertionError):
        FeatParams(feature_type="fbank")
    with pytest.raises(AssertionError):
        FeatParams(log_base="2")


def test_nframes_known_answers(golden_dir):
    g = np.load(os.path.join(golden_dir, "lengths.npz"))
    p = yaml_params()
    for n, t in zip(g["nsamples"], g["n_frames"]):
        assert oracle.get_nframes(int(n), p) == int(t)
    # 10 s, 15 s, 30 s, 1 s
    assert [oracle.get_nframes(n) for n in (160000, 240000, 480000, 16000)] == [998, 1498, 2998, 98]
    assert oracle.get_nframes(399) == 0 and oracle.get_nframes(400) == 1
    assert oracle.get_nframes(559) == 1 and oracle.get_nframes(560) == 2
    assert oracle.get_nframes(399, clamp=False) == 0  # 1 + (-1)//160
    assert oracle.get_nframes(100, clamp=False) == -1


def test_hann_periodic_matches_scipy():
    from scipy.signal import get_window
    w = oracle.hann_periodic(400)
    ref = get_window("hann", 400, fftbins=True)
    assert w.dtype == np.float32 and w.shape == (400,)
    assert w[0] == 0.0 and abs(w[200] - 1.0) < 1e-7
    np.testing.assert_allclose(w, ref, atol=2e-7)


def test_mel_matrix_invariants():
    W = oracle.htk_mel_matrix_f32()
    assert W.shape == (257, 80) and W.dtype == np.float32
    assert not W[0].any(), "DC row is zero (HTK excludes the DC bin)"
    assert not W[256].any(), "bin at upper_edge_hertz has zero weight"
    nz = W != 0
    assert nz.sum() == 502
    assert nz.sum(axis=1).max() == 2
    assert nz.sum(axis=0).min() >= 1 and nz.sum(axis=0).max() == 16
    assert (W >= 0).all() and W.max() <= 1.0
    # non-zeros of every mel bin are one contiguous run (what the banded kernel relies on)
    for m in range(80):
        idx = np.flatnonzero(nz[:, m])
        assert (np.diff(idx) == 1).all()
    # adjacent triangles share edges: interior bins' two weights sum to 1 (linear interpolation in mel)
    rows = nz.sum(axis=1) == 2
    np.testing.assert_allclose(W[rows].sum(axis=1), 1.0, atol=2e-5)


def test_mel_matrix_vs_float64_formula_and_torchaudio():
    W = oracle.htk_mel_matrix_f32().astype(np.float64)
    # independent float64 evaluation of the published HTK formula
    mel = lambda f: 1127.0 * np.log1p(f / 700.0)
    bins = mel(np.linspace(0.0, 8000.0, 257)[1:])[:, None]
    e = np.linspace(mel(0.0), mel(8000.0), 82)
    lo, ce, hi = e[None, :-2], e[None, 1:-1], e[None, 2:]
    ref = np.maximum(0.0, np.minimum((bins - lo) / (ce - lo), (hi - bins) / (hi - ce)))
    ref = np.pad(ref, [[1, 0], [0, 0]])
    assert np.abs(W - ref).max() < 5e-5          # float32 op-order noise only
    try:                                        # torchaudio is not in this image; the HF pin below is the mandatory one
        import torchaudio
    except ImportError:
        return
    fb = torchaudio.functional.melscale_fbanks(257, 0.0, 8000.0, 80, 16000, norm=None, mel_scale="htk").numpy()
    # torchaudio's triangles are linear in Hz, TF's in mel: same supports, weights within 0.5 %
    assert np.abs(fb - W).max() < 5e-3
    assert ((fb > 1e-4) == (W > 1e-4)).mean() > 0.995


def _dft_power_f64(frame_windowed, nfft=512):
    n = np.arange(nfft)[None, :]
    k = np.arange(nfft // 2 + 1)[:, None]
    x = np.zeros(nfft)
    x[: frame_windowed.shape[0]] = frame_windowed      # zero padded at the TAIL
    X = (np.exp(-2j * np.pi * k * n / nfft) * x[None, :]).sum(axis=1)
    return np.abs(X) ** 2


def test_stft_against_direct_dft_and_closed_forms():
    p = FeatParams(normalize_signal=False, preemphasis=0.0)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(400 + 160 * 3)
    S = _stft_power(x, p, np.float64)
    assert S.shape == (4, 257)
    w = oracle.hann_periodic(400).astype(np.float64)
    for i in range(4):
        np.testing.assert_allclose(S[i], _dft_power_f64(x[160 * i: 160 * i + 400] * w), rtol=1e-9, atol=1e-9)
    # impulse at the window centre: w[200] = 1 -> flat unit power spectrum
    imp = np.zeros(400); imp[200] = 1.0
    np.testing.assert_allclose(_stft_power(imp, p, np.float64)[0], 1.0, atol=1e-6)
    # impulse at n=0 is killed by the periodic Hann (w[0] = 0)
    imp0 = np.zeros(400); imp0[0] = 1.0
    assert np.abs(_stft_power(imp0, p, np.float64)).max() < 1e-12
    # DC: X[0] = sum(w) = 200 exactly for a periodic Hann of even length
    dc = _stft_power(np.ones(400), p, np.float64)[0]
    assert abs(dc[0] - 200.0 ** 2) < 1e-2
    # frame/offset: pure cosine at a 512-bin centre peaks at that bin
    t = np.arange(400)
    tone = _stft_power(np.cos(2 * np.pi * 64 * t / 512), p, np.float64)[0]
    assert int(np.argmax(tone)) == 64


def test_preemphasis_and_gain_order():
    x = np.array([0.5, -0.25, 0.125, 0.0625] + [0.0] * 396, dtype=np.float32)
    p = FeatParams(feature_type="waveform", normalize_signal=True, preemphasis=0.97)
    y = featurize_ref(x, p, np.float64)
    g = 1.0 / (0.5 + 1e-9)
    xn = x.astype(np.float64) * g
    np.testing.assert_allclose(y[0], xn[0])
    np.testing.assert_allclose(y[1:4], xn[1:4] - 0.97 * xn[0:3], rtol=1e-12)


def test_silence_floor_and_collate_padding():
    p = yaml_params()
    z = np.zeros(1000, dtype=np.float32)
    f = oracle.logmel_ref(z, p, np.float32)
    assert f.shape == (4, 80) and np.all(f == np.float32(-9.0))
    other = oracle.logmel_ref(oracle.make_waveforms([1600], seed=1)[0][0, :1600], p, np.float32)
    out, n = oracle.collate_ref([f, other])
    assert out.shape == (2, 8, 80, 1) and list(n) == [4, 8]
    assert np.all(out[0, 4:] == 0.0) and np.all(out[0, :4] == -9.0)


def test_float32_oracle_runs_a_genuine_float32_fft():
    """The float32 oracle's FFT stage must compute in single precision, like the reference's TF op does.  numpy's
    np.fft.rfft on float32 input is the float64 transform rounded once (relative rms error 2.5e-8 = output rounding
    only) - a band no float32 FFT can live in; scipy.fft (what the oracle uses) and torch.fft (MKL / pocketfft, an
    independent float32 FFT) both sit at ~1e-7.  Pins which one the oracle uses and that the two agree."""
    import scipy.fft
    import torch
    from oracle import featurizer_ref as fr
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((256, 400)) * fr.hann_periodic(400)).astype(np.float32)
    ref = np.fft.rfft(x.astype(np.float64), n=512, axis=1)
    rms = np.sqrt((np.abs(ref) ** 2).mean())
    err = lambda X: float(np.sqrt((np.abs(X - ref) ** 2).mean()) / rms)
    e_np = err(np.fft.rfft(x, n=512, axis=1))
    e_sp = err(scipy.fft.rfft(x, n=512, axis=1))
    e_th = err(torch.fft.rfft(torch.from_numpy(x), n=512, dim=1).numpy())
    assert e_np < 4e-8                               # double inside: not a float32 FFT
    assert 6e-8 < e_sp < 2e-7 and 6e-8 < e_th < 2e-7 and 0.5 < e_sp / e_th < 2.0
    p = fr.yaml_params()
    P32 = fr._stft_power(x[0], fr.FeatParams(**{**p.__dict__, "pad_end": False}), np.float32)
    X1 = scipy.fft.rfft((x[0, :400] * fr.hann_periodic(400)).astype(np.float32), n=512)
    np.testing.assert_array_equal(P32[0], np.square(np.abs(X1).astype(np.float32)))   # the oracle's spectrum IS scipy's float32 one


@pytest.mark.parametrize("dist,band", [("tilt", 3e-5), ("white", 2e-3), ("half_silence", 3e-5)])
def test_float32_band_vs_float64(dist, band):
    """The error band any float32 implementation of the path lives in (SURVEY.md hard part 1)."""
    wav, ln = oracle.make_waveforms([48000], seed=0, dist=dist)
    a = oracle.logmel_ref(wav[0, :48000], dtype=np.float32)
    b = oracle.logmel_ref(wav[0, :48000], dtype=np.float64)
    assert a.dtype == np.float32 and b.dtype == np.float64
    assert np.abs(a - b).max() < band


def test_batched_equals_per_utterance_loop():
    lens = [399, 400, 3000, 16000, 12345]
    wav, ln = oracle.make_waveforms(lens, seed=4, dist="white")
    out, n = oracle.logmel_batch_ref(wav, ln, dtype=np.float32)
    assert out.shape == (5, 98, 80, 1)
    for b, L in enumerate(lens):
        f = oracle.logmel_ref(wav[b, :L], dtype=np.float32)
        assert n[b] == f.shape[0]
        np.testing.assert_array_equal(out[b, : f.shape[0], :, 0], f)
        assert not out[b, f.shape[0]:].any()


def test_golden_logmel_fixtures(golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_tilt_1s.npz"))
    wav, ln = oracle.make_waveforms(g["lengths"], seed=int(g["seed"]), dist=str(g["dist"]))
    np.testing.assert_allclose(oracle.logmel_ref(wav[0, : ln[0]], dtype=np.float64), g["f64"], rtol=0, atol=1e-10)
    assert np.abs(g["f32"].astype(np.float64) - g["f64"]).max() < 3e-5
    for dist in ("white", "half_silence"):
        g = np.load(os.path.join(golden_dir, f"logmel_ragged_{dist}.npz"))
        wav, ln = oracle.make_waveforms(g["lengths"], seed=int(g["seed"]), dist=dist)
        out, n = oracle.logmel_batch_ref(wav, ln, dtype=np.float64)
        np.testing.assert_array_equal(n, g["n_frames"])
        np.testing.assert_allclose(out, g["f64"], rtol=0, atol=1e-10)


def test_other_feature_types_shapes():
    x = oracle.make_waveforms([4000], seed=2)[0][0, :4000]
    spec = featurize_ref(x, FeatParams(feature_type="spectrogram", normalize_signal=True), np.float32)
    assert spec.shape == (23, 80)
    mf = featurize_ref(x, FeatParams(feature_type="mfcc", normalize_signal=True), np.float32)
    assert mf.shape == (23, 80)
    ln = featurize_ref(x, FeatParams(log_base="e", normalize_signal=True), np.float32)
    l10 = featurize_ref(x, FeatParams(log_base="10", normalize_signal=True), np.float32)
    np.testing.assert_allclose(ln, l10 * np.log(10.0), rtol=1e-5, atol=1e-5)
    pe = featurize_ref(x, FeatParams(pad_end=True, normalize_signal=True), np.float32)
    assert pe.shape == (25, 80)


def test_logmel_oracle_vs_transformers_audio_utils():
    """Independent third-party pin of the whole log-mel chain: `transformers.audio_utils` (numpy; written to reproduce
    the TF / Kaldi front ends — `triangularize_in_mel_space=True` is its switch for tf.signal.linear_to_mel_weight_matrix)
    with center=False frames, a periodic Hann window, the frame zero-padded at the TAIL to 512, power spectrum, HTK
    triangles, log10 with a 1e-9 floor.  Same support of the mel matrix (502 non-zeros, same places), weights within
    1e-5 (TF builds the matrix in float32, HF in float64), log-mel within 5e-5 end to end and within 5e-6 when HF's
    chain is handed the oracle's float32-built matrix.  (Normalisation and pre-emphasis are applied to the signal
    first, as src/speech_featurizer.py:68-79 does; HF's own per-frame Kaldi pre-emphasis is a different operation.)"""
    import transformers.audio_utils as au      # mandatory: this is the oracle's only third-party end-to-end pin
    from oracle import featurizer_ref as fr
    W = fr.htk_mel_matrix_f32()
    fb = au.mel_filter_bank(num_frequency_bins=257, num_mel_filters=80, min_frequency=0.0, max_frequency=8000.0,
                            sampling_rate=16000, norm=None, mel_scale="htk", triangularize_in_mel_space=True)
    assert fb.shape == W.shape == (257, 80)
    assert np.array_equal(W != 0, fb > 1e-12) and int((W != 0).sum()) == 502
    assert np.abs(W - fb).max() <= 1e-5
    win = au.window_function(400, "hann", periodic=True)
    wav, ln = oracle.make_waveforms([16000, 5519, 160000], seed=5, dist="tilt")
    for b in range(3):
        x = wav[b, : ln[b]].astype(np.float64)
        xn = x * (1.0 / (np.abs(x).max() + 1e-9))
        y = np.concatenate([xn[:1], xn[1:] - 0.97 * xn[:-1]])
        r64 = oracle.logmel_ref(wav[b, : ln[b]], dtype=np.float64)
        for mel, tol in ((fb, 5e-5), (W.astype(np.float64), 5e-6)):
            S = au.spectrogram(y, win, frame_length=400, hop_length=160, fft_length=512, power=2.0, center=False,
                               preemphasis=None, mel_filters=mel, mel_floor=1e-9, log_mel="log10", dtype=np.float64)
            assert S.T.shape == r64.shape                      # same frame count: 1 + (N - 400) // 160
            assert np.abs(S.T - r64).max() <= tol, (b, np.abs(S.T - r64).max())
