"""CPU tests of the host-side mirror (no GPU): constructor parity with the reference classes,
static shape / length arithmetic, tables, sharding, and that the C-ABI library loads and exports
every symbol include/tasr.h declares.  No compute call is made here."""
import ctypes
import os

import numpy as np
import pytest
import torch

import oracle
import telugu_asr_b200 as tasr
from telugu_asr_b200 import _native, tables
from telugu_asr_b200.synth import make_waveforms, draw_lengths, to_pcm16


def test_library_loads_and_exports_every_header_symbol():
    lib = _native.lib()
    syms = _native.header_symbols()
    assert "tasr_logmel_f32" in syms and "tasr_sepconv1d_f32" in syms and len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"libtasr_b200.so does not export {s}"
        assert s in _native._SIGNATURES, f"{s} has no ctypes signature"
    assert lib.tasr_version() >= 100


def test_library_is_sm100a_only():
    import subprocess, shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    archs = {l.split(".sm_")[1].split(".")[0] for l in out.splitlines() if ".sm_" in l}
    assert archs == {"100a"}, archs


def test_c_abi_argument_errors_without_gpu():
    """Validation happens before any CUDA call, so these are safe on a CPU-only box."""
    lib = _native.lib()
    p = _native.TasrFeatParams(16000, 400, 160, 512, 80, 1, 0, 0, 0.97, 1e-9)
    out = ctypes.c_void_p()
    assert lib.tasr_featurizer_create(ctypes.byref(p), None, None, ctypes.byref(out)) == _native.TASR_ERR_BAD_ARG
    assert b"null" in lib.tasr_last_error()
    hann = tables.hann_window_f32(400)
    mel = tables.mel_weight_matrix_f32(80, 257, 16000, 0.0, 8000.0)
    bad = _native.TasrFeatParams(16000, 320, 160, 500, 80, 1, 0, 0, 0.97, 1e-9)     # fft_length is not a power of two
    rc = lib.tasr_featurizer_create(ctypes.byref(bad), hann.ctypes.data_as(ctypes.c_void_p),
                                    mel.ctypes.data_as(ctypes.c_void_p), ctypes.byref(out))
    assert rc == _native.TASR_ERR_UNSUPPORTED
    with pytest.raises(NotImplementedError):
        _native.check(rc)
    assert lib.tasr_absmax_f32(None, None, 1, 4, None, None) == _native.TASR_ERR_BAD_ARG
    with pytest.raises(ValueError):
        _native.check(_native.TASR_ERR_BAD_ARG)


def test_product_tables_equal_oracle_tables_bitwise():
    assert np.array_equal(tables.hann_window_f32(400), oracle.hann_periodic(400))
    assert np.array_equal(tables.mel_weight_matrix_f32(80, 257, 16000, 0.0, 8000.0), oracle.htk_mel_matrix_f32())
    assert tables.enclosing_power_of_two(400) == 512 and tables.enclosing_power_of_two(512) == 512


def test_featurizer_constructor_mirrors_reference():
    f = tasr.SpeechFeaturizer(**tasr.REFERENCE_SPEECH_CONFIG)
    assert (f.frame_length, f.frame_step, f.num_feature_bins, f.sample_rate) == (400, 160, 80, 16000)
    assert f.fft_length == 512 and f.nfft == 512
    assert f.get_nframes(160000) == 998 and f.get_nframes(399) == 0 and f.get_nframes(100) == -1
    assert f.compute_output_shape((4, 160000)) == (4, 998, 80, 1)
    assert f.compute_output_shape((4, None)) == (4, None, 80, 1)
    assert f.get_config()["normalize_signal"] is True
    with pytest.raises(AssertionError):
        tasr.SpeechFeaturizer(feature_type="fbank")
    with pytest.raises(AssertionError):
        tasr.SpeechFeaturizer(log_base="2")
    # nfft=None falls back to frame_length like src/speech_featurizer.py:65
    assert tasr.SpeechFeaturizer(nfft=None).nfft == 400
    with pytest.raises(AttributeError):
        f(torch.zeros(400), training=True)


def test_no_cpu_fallback():
    f = tasr.SpeechFeaturizer(**tasr.REFERENCE_SPEECH_CONFIG)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        f(torch.zeros(16000))
    s = tasr.Conv1DSubsamplingLayer(192, tasr.REFERENCE_SUBSAMPLING_CONFIG)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        s(torch.zeros(1, 100, 80, 1))


def test_subsampling_constructor_mirrors_reference():
    s = tasr.Conv1DSubsamplingLayer(192, tasr.REFERENCE_SUBSAMPLING_CONFIG)
    assert s.filters == [192, 384, 192] and s.kernel_size == [9, 9, 9] and s.strides == [2, 2, 2]
    assert s.padding == ["valid"] * 3
    # YAML key is `activation`, layer reads `activations` -> default tanh/gelu/gelu
    assert s.activations == ["tanh", "gelu", "gelu"]
    assert tasr.Conv1DSubsamplingLayer(192, dict(tasr.REFERENCE_SUBSAMPLING_CONFIG, activations=["gelu"] * 3)).activations == ["gelu"] * 3
    assert s.compute_output_shape((8, 1498, 80, 1)) == (8, 181, 192)
    assert s.compute_output_shape((8, None, 80, 1)) == (8, None, 192)
    assert s.get_config()["model_dim"] == 192
    with pytest.raises(ValueError, match="same length"):
        tasr.Conv1DSubsamplingLayer(192, dict(kernel_size=[9, 9], strides=[2, 2, 2], padding=["valid"] * 3))
    d = tasr.Conv1DSubsamplingLayer()  # reference defaults: model_dim 288, 'same' padding
    assert d.filters == [288, 576, 288] and d.padding == ["same"] * 3
    for L in list(range(0, 40)) + [998, 1498, 2998]:
        assert tasr.get_conv_length(L, 9, "valid", 2) == int(oracle.conv_length_f32_trunc(L, 9, "valid", 2))
        assert tasr.get_conv_length(L, 9, "same", 2) == int(oracle.conv_length_f32_trunc(L, 9, "same", 2))


def test_reference_yaml_loader(tmp_path):
    y = tmp_path / "model.yaml"
    y.write_text("speech_config:\n  sample_rate: 16000\n  frame_ms: 25\n  stride_ms: 10\n  num_feature_bins: 80\n"
                 "  feature_type: log_mel_spectrogram\n  preemphasis: 0.97\n  log_base: \"10\"\n  nfft: 512\n"
                 "  normalize_signal: True\nmodel_config:\n  d_model: 192\n  subsampling_config:\n    name: conv1d\n"
                 "    kernel_size: [9,9,9]\n    strides: [2,2,2]\n    padding: [\"valid\", \"valid\", \"valid\"]\n"
                 "    activation: [\"gelu\", \"gelu\", \"gelu\"]\n")
    sc, sub, d = tasr.load_reference_yaml(str(y))
    f = tasr.SpeechFeaturizer(**sc)
    assert f.frame_length == 400 and f.log_base == "10"
    s = tasr.Conv1DSubsamplingLayer(d, sub)
    assert s.activations == ["tanh", "gelu", "gelu"] and s.filters == [192, 384, 192]


def test_shard_by_length_properties():
    rng = np.random.default_rng(0)
    lens = rng.integers(16000, 240001, size=256)
    for world in (1, 2, 4, 8):
        shards = tasr.shard_by_length(lens, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(256))
        loads = [int(lens[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= int(lens.max())
    assert tasr.shard_by_length([], 4) == [[], [], [], []]
    eq = tasr.shard_by_length([480000] * 1024, 8)
    assert all(len(s) == 128 for s in eq)


def test_synth_is_seeded_pcm_exact_and_aligned():
    lens = draw_lengths(8, 16000, 240000, seed=2)
    assert lens[0] == 240000 and lens.min() >= 16000
    a, la = make_waveforms(lens[:3], seed=2)
    b, lb = make_waveforms(lens[:3], seed=2)
    assert np.array_equal(a, b) and a.shape[1] % 4 == 0 and a.dtype == np.float32
    assert abs(np.abs(a[0]).max() - 0.5) < 1e-4
    pcm = to_pcm16(a)
    assert np.array_equal(pcm.astype(np.float32) / 32768.0, a)       # k/32768 values, as decode_wav gives
    assert not a[1, la[1]:].any()
    z, _ = make_waveforms([1000], seed=0, dist="zeros")
    assert not z.any()
    with pytest.raises(ValueError):
        make_waveforms([10], dist="pink")


def test_packed_batch_offsets_are_aligned_and_disjoint():
    lens = [16000, 399, 1, 0, 5, 8191, 8192, 8193]
    off, used = tasr.PackedBatch.offsets_for(lens)
    assert (off % 8 == 0).all() and off[0] == 0
    ends = off + np.asarray(lens)
    assert (ends[:-1] <= off[1:]).all() and used >= ends[-1] and used % 8 == 0
    assert used - sum(lens) < 8 * len(lens)
    o0, u0 = tasr.PackedBatch.offsets_for([])
    assert len(o0) == 0 and u0 == 0


def test_builtin_reference_config_equals_the_reference_yaml():
    """When the reference checkout is present (this container; not the GPU box), the constants this package ships
    as REFERENCE_SPEECH_CONFIG / REFERENCE_SUBSAMPLING_CONFIG / d_model must be the reference's own
    config/model.yaml values, key for key (config/conformer.yaml carries the same speech_config)."""
    path = "/root/reference/config/model.yaml"
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    sc, sub, d = tasr.load_reference_yaml(path)
    for k, v in tasr.REFERENCE_SPEECH_CONFIG.items():
        assert k in sc, k
        assert (str(sc[k]) == str(v)) or (sc[k] == v), (k, sc[k], v)
    assert set(sc) - set(tasr.REFERENCE_SPEECH_CONFIG) <= {"augmentation_config"}
    assert d == tasr.frontend.REFERENCE_D_MODEL == 192
    for k in ("kernel_size", "strides", "padding"):
        assert list(sub[k]) == list(tasr.REFERENCE_SUBSAMPLING_CONFIG[k]), k
    assert "activations" not in sub and "activation" in sub      # the key quirk the layer mirrors (encoder.py:25)
    f = tasr.SpeechFeaturizer(**sc)
    assert (f.frame_length, f.frame_step, f.fft_length, f.num_feature_bins) == (400, 160, 512, 80)
    assert isinstance(sc["output_floor"], float) and sc["output_floor"] == 1e-9   # `1e-9` in the YAML: OmegaConf reads a float
    sc2, _, _ = tasr.load_reference_yaml("/root/reference/config/conformer.yaml")
    assert sc2 == sc


def test_conv2d_subsampling_host_mirror_without_gpu():
    """Conv2dSubsampling (src/models/conformer/encoder.py:9-48): constructor defaults, the geometry it refuses, the
    static output shape against the oracle's TensorFlow "SAME" arithmetic — host logic only, no device needed."""
    import oracle
    layer = tasr.Conv2dSubsampling({"name": "conv2d", "filters": 144, "kernel_size": 3, "strides": 2, "padding": "same"})
    assert (layer.filter, layer.kernel_size, layer.stride, layer.padding) == (144, 3, 2, "same")
    assert tasr.Conv2dSubsampling({}).filter == 128                       # encoder.py:22 default
    for bad in ({"kernel_size": 5}, {"strides": 1}, {"padding": "valid"}):
        with pytest.raises(NotImplementedError):
            tasr.Conv2dSubsampling({"filters": 144, **bad})
    for T in (1, 2, 37, 38, 1497, 1498):
        h1, _, _ = oracle.same_pads(T, 3, 2)
        h2, _, _ = oracle.same_pads(h1, 3, 2)
        assert layer.output_shape(T, 80) == (h1, 40, h2, 20)
        assert layer.compute_output_shape((7, T, 80, 1)) == (7, h2, 20 * 144)
    with pytest.raises(RuntimeError):                                     # no CPU path
        layer([torch.zeros(1, 8, 80, 1), None])


def test_front_end_defaults_and_flags():
    fe = tasr.FrontEnd()
    assert fe.single_pass and fe.lean_intermediates and fe.subsampling.math == "tf32"
    assert fe.featurizer.supports_single_pass()
    assert not tasr.SpeechFeaturizer(**{**tasr.REFERENCE_SPEECH_CONFIG, "normalize_signal": False}).supports_single_pass()
    assert not tasr.SpeechFeaturizer(**{**tasr.REFERENCE_SPEECH_CONFIG, "pad_end": True}).supports_single_pass()
    assert tasr.Conv1DSubsamplingLayer.ragged_margin() >= 38               # rows a lean producer must keep filled
