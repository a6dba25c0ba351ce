"""telugu_asr_b200 — B200-native (sm_100a) front-end hot path of HemanthSai7/Telugu-ASR:
waveform -> log-mel (SpeechFeaturizer) -> 3x depthwise-separable Conv1D subsampling
(Conv1DSubsamplingLayer) -> [B, T/8, 192] encoder input + padding mask.

Everything numeric runs in hand-written CUDA behind the C ABI of include/tasr.h
(libtasr_b200.so).  Importing this package does not need a GPU; calling an operator does, and
fails loudly without one."""
from .speech_featurizer import SpeechFeaturizer, FeaturizerConfig  # noqa: F401
from .subsampling import Conv1DSubsamplingLayer, get_conv_length  # noqa: F401
from .conv2d_subsampling import Conv2dSubsampling  # noqa: F401
from .encoder_block import EncoderBlock  # noqa: F401
from .frontend import FrontEnd, ConformerFrontEnd, CapturedFrontEnd, InterleavedFrontEnd, REFERENCE_SPEECH_CONFIG, REFERENCE_SUBSAMPLING_CONFIG, load_reference_yaml  # noqa: F401
from .collate import pack_waveforms, PinnedBatch, PackedBatch, shard_by_length  # noqa: F401

from .augmentation import Augmentation, FreqMasking, TimeMasking, AUGMENTATIONS  # noqa: F401
from .pipeline import FrontEndPipeline, Ticket  # noqa: F401

__version__ = "0.1.0"
