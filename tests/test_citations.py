"""The reference file:line citations used throughout this repository (headers, docstrings) point at what
they claim to point at.  Runs only where the reference checkout is present (this container); skipped on the GPU box."""
import os

import pytest

REF = "/root/reference"

CITATIONS = [
    # (file, first line, last line, substring that must appear in that range)
    ("src/speech_featurizer.py", 46, 46, "self.frame_length = int(round("),
    ("src/speech_featurizer.py", 49, 49, "self.frame_step = int(round("),
    ("src/speech_featurizer.py", 65, 66, "self.nfft"),
    ("src/speech_featurizer.py", 68, 72, "tf.reduce_max(tf.abs(signal)"),
    ("src/speech_featurizer.py", 74, 79, "self.preemphasis * signal[:-1]"),
    ("src/speech_featurizer.py", 81, 93, "reduce_variance"),
    ("src/speech_featurizer.py", 95, 105, "tf.signal.stft("),
    ("src/speech_featurizer.py", 107, 110, "tf.maximum(S, self.output_floor)"),
    ("src/speech_featurizer.py", 112, 122, "linear_to_mel_weight_matrix"),
    ("src/speech_featurizer.py", 124, 126, "[:, :, :self.num_feature_bins]"),
    ("src/speech_featurizer.py", 128, 130, "mfccs_from_log_mel_spectrograms"),
    ("src/speech_featurizer.py", 158, 159, "self.augmentation.signal_augment"),
    ("src/speech_featurizer.py", 163, 166, "1 + (nsamples - self.frame_length) // self.frame_step"),
    ("src/utils/math_util.py", 17, 18, "tf.math.log(x) / tf.math.log(10.0)"),
    ("src/utils/math_util.py", 20, 32, "def get_conv_length"),
    ("src/models/moonshine/encoder.py", 21, 21, "self.filters = [model_dim"),
    ("src/models/moonshine/encoder.py", 25, 25, '"activations"'),
    ("src/models/moonshine/encoder.py", 26, 27, "must have the same length"),
    ("src/models/moonshine/encoder.py", 31, 40, "SeparableConv1D"),
    ("src/models/moonshine/encoder.py", 43, 48, "def lengths_to_padding_mask"),
    ("src/models/moonshine/encoder.py", 50, 71, "get_conv_length"),
    ("src/models/moonshine/model.py", 80, 80, "tf.not_equal"),
    ("src/dataset.py", 167, 175, "self.speech_featurizer(audio_inputs"),
    ("src/dataset.py", 172, 172, "feature_augment"),
    ("src/dataset.py", 236, 252, "padded_batch"),
    ("src/utils/data_util.py", 10, 38, "decode_wav"),
    ("src/augmentations/specaugment.py", 6, 31, "class FreqMasking"),
    ("src/augmentations/specaugment.py", 34, 62, "class TimeMasking"),
    ("src/augmentations/augmentation.py", 19, 35, "tf.less(p, self.prob)"),
    ("src/helpers/dataset_helpers.py", 68, 68, "SpeechFeaturizer("),
    ("config/model.yaml", 1, 17, "normalize_signal: True"),
    ("config/model.yaml", 24, 27, "activation:"),
]


@pytest.mark.parametrize("path,lo,hi,needle", CITATIONS)
def test_citation(path, lo, hi, needle):
    full = os.path.join(REF, path)
    if not os.path.exists(full):
        pytest.skip("reference checkout not present")
    lines = open(full, encoding="utf-8", errors="replace").read().splitlines()
    chunk = "\n".join(lines[lo - 1: hi])
    assert needle in chunk, f"{path}:{lo}-{hi} does not contain {needle!r}:\n{chunk[:600]}"
