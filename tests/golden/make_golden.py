"""Generates the committed fixtures under tests/golden/ from the CPU oracle.

    python tests/golden/make_golden.py

The reference ships no golden vectors and TensorFlow is not installable here, so
these fixtures pin the ORACLE (against accidental change) and give the GPU tests a file-based
target that does not depend on importing oracle/ at all.  Inputs are regenerated from seeds by
telugu_asr_b200.synth; only outputs are stored (float64 'exact' evaluation + float32 op-order
evaluation where the difference matters)."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle.featurizer_ref import yaml_params  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    p = yaml_params()
    # 1. log-mel, one 1 s "tilt" utterance (config 1 in miniature)
    wav, ln = oracle.make_waveforms([16000], seed=0, dist="tilt")
    np.savez_compressed(os.path.join(OUT, "logmel_tilt_1s.npz"),
                        seed=0, dist="tilt", lengths=ln,
                        f64=oracle.logmel_ref(wav[0, :16000], p, np.float64),
                        f32=oracle.logmel_ref(wav[0, :16000], p, np.float32))
    # 2. ragged batch incl. edge lengths (399 -> 0 frames, 400 -> 1 frame, 559/560 boundary)
    lens = [399, 400, 559, 560, 5000, 16000, 0]
    for dist in ("white", "half_silence"):
        wav, ln = oracle.make_waveforms(lens, seed=11, dist=dist)
        f64, nf = oracle.logmel_batch_ref(wav, ln, p, np.float64)
        np.savez_compressed(os.path.join(OUT, f"logmel_ragged_{dist}.npz"), seed=11, dist=dist, lengths=ln,
                            f64=f64, n_frames=nf)
    # 3. subsampling on 2 x 3 s, ragged, glorot weights seed 7, tanh/gelu/gelu
    lens = [48000, 30000]
    wav, ln = oracle.make_waveforms(lens, seed=5, dist="tilt")
    feat32, nf = oracle.logmel_batch_ref(wav, ln, p, np.float32)
    weights = oracle.glorot_subsampling_weights(192, 80, seed=7)
    out64, mask, len_all = oracle.subsample_ref(feat32, nf, weights, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "subsample_2x3s.npz"), seed=5, weight_seed=7, lengths=ln,
                        n_frames=nf, feat32=feat32, out64=out64, mask=mask, len_all=len_all)
    # 4. integer tables: frames and conv lengths (SURVEY.md §8c known-answer list)
    ns = np.array([0, 1, 399, 400, 559, 560, 719, 720, 16000, 27520, 160000, 240000, 283680, 480000], dtype=np.int64)
    frames = np.array([oracle.get_nframes(int(n), p) for n in ns], dtype=np.int32)
    convs = oracle.conv_lengths_ref(frames)
    small = np.arange(0, 64, dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "lengths.npz"), nsamples=ns, n_frames=frames, conv_lengths=convs,
                        small_in=small, small_out=oracle.conv_lengths_ref(small))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
