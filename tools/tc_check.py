"""Development aid: run the log-mel kernel (TASR_LOGMEL_TC=1 tensor-core / 0 CUDA-core) on a few small batches, print
the max-abs error against the float64 oracle and the float32 oracle's own band, and time the config-3 batch.

    TASR_LOGMEL_TC=1 python tools/tc_check.py [--time]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
import telugu_asr_b200 as tasr

dev = torch.device("cuda:0")
feat = tasr.SpeechFeaturizer(**tasr.REFERENCE_SPEECH_CONFIG)


def run(wav, ln):
    out, nf = feat(torch.from_numpy(wav).to(dev), torch.from_numpy(ln).to(dev))
    torch.cuda.synchronize()
    return out.cpu().numpy(), nf.cpu().numpy()


print("TASR_LOGMEL_TC =", os.environ.get("TASR_LOGMEL_TC", "(default 1)"), flush=True)
for dist, lens in () if "--time-only" in sys.argv else (("tilt", [16000]), ("tilt", [5520, 400, 399, 48000, 16001]), ("white", [32000, 8000]),
                   ("tone_noise", [32000]), ("half_silence", [32000, 16000]), ("zeros", [8000])):
    wav, ln = oracle.make_waveforms(lens, seed=3, dist=dist)
    t0 = time.time()
    out, nf = run(wav, ln)
    worst = band = 0.0
    for b in range(len(ln)):
        r64 = oracle.logmel_ref(wav[b, : ln[b]], dtype=np.float64)
        r32 = oracle.logmel_ref(wav[b, : ln[b]], dtype=np.float32)
        T = r64.shape[0]
        assert nf[b] == T, (nf[b], T)
        if T:
            worst = max(worst, float(np.abs(out[b, :T, :, 0] - r64).max()))
            band = max(band, float(np.abs(r32 - r64).max()))
        assert not np.any(out[b, T:]), "padding rows must be 0.0"
    print(f"{dist:13s} lens={lens} max|err|={worst:.3e} float32-oracle band={band:.3e}  ({time.time() - t0:.2f}s)", flush=True)

if "--time" in sys.argv or "--time-only" in sys.argv:
    lens = tasr.synth.draw_lengths(256, 16000, 240000, seed=2)
    wav, ln = oracle.make_waveforms(lens, seed=2, dist="tilt")
    w, l = torch.from_numpy(wav).to(dev), torch.from_numpy(ln).to(dev)
    for _ in range(3):
        feat(w, l)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        feat(w, l)
    e1.record()
    torch.cuda.synchronize()
    print(f"config-3 batch, SpeechFeaturizer two-pass call: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call", flush=True)
