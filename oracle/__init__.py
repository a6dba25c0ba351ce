"""CPU oracle for the Telugu-ASR front-end hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of the reference's
``SpeechFeaturizer`` (src/speech_featurizer.py), ``Conv1DSubsamplingLayer``
(src/models/moonshine/encoder.py:9-105) and the conformer configuration's ``Conv2dSubsampling``
(src/models/conformer/encoder.py:9-73).  It is the *checker* for the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``telugu_asr_b200/`` imports it, and the product path has no CPU fallback.

PARITY UNPINNED: the reference ships no tests, golden vectors, fixtures, audio
or checkpoints for this path, and its arithmetic lives in TensorFlow 2.15.0.post1
/ Keras 2.15 (requirements.txt:1), which is not installable here (no network).
The TF op semantics are restated from the published TF 2.15 sources
(tensorflow/python/ops/signal/{spectral_ops,window_ops,shape_ops,mel_ops}.py,
keras/layers/convolutional/separable_conv1d.py).  What pins the oracle instead is
listed in tests/test_oracle_*.py: closed-form spectra, independent library
implementations (torchaudio HTK filterbank, torch conv1d, direct float64 DFT),
and the committed fixtures under tests/golden/ produced by
tests/golden/make_golden.py.
"""
from .featurizer_ref import (  # noqa: F401
    FeatParams,
    hann_periodic,
    htk_mel_matrix_f32,
    get_nframes,
    logmel_ref,
    logmel_batch_ref,
    collate_ref,
)
from .subsampling_ref import (  # noqa: F401
    conv_length_f32_trunc,
    conv_lengths_ref,
    lengths_to_padding_mask_ref,
    create_audio_mask_ref,
    sepconv1d_ref,
    subsample_ref,
    glorot_subsampling_weights,
)
from .conv2d_subsampling_ref import (  # noqa: F401
    same_pads,
    conv2d_same_ref,
    conv2d_subsample_ref,
    glorot_conv2d_weights,
)
from .synth import make_waveforms  # noqa: F401
from . import specaugment_ref  # noqa: F401
from .encoder_block_ref import (  # noqa: F401
    rope_tables,
    rope_apply,
    encoder_block_ref,
    glorot_encoder_block_weights,
)
