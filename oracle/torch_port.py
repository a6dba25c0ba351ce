"""torch-CPU restatement of the path, used ONLY as the timed CPU baseline (bench.py's
`cpu_baseline` and `--impl reference` legs) and validated against the numpy oracle in
tests/test_oracle_torch_port.py.  TEST INFRASTRUCTURE — see oracle/__init__.py.

It runs the way the reference runs (SURVEY.md §8d "CPU baseline beside it"): one featurizer call
per utterance (src/dataset.py:171), zero-pad collate (:236-252), then the three separable convs on
the padded batch (encoder.py:50-71), with torch's multi-threaded CPU kernels standing in for
TensorFlow's Eigen kernels.  The mel matrix is rebuilt on every call, as the reference does
(src/speech_featurizer.py:114-120)."""
from __future__ import annotations

import numpy as np
import torch

from .featurizer_ref import FeatParams, yaml_params, hann_periodic, htk_mel_matrix_f32

LN10 = float(np.log(np.float32(10.0)))


def logmel_torch(x: torch.Tensor, p: FeatParams | None = None) -> torch.Tensor:
    """x [N] float32 CPU -> [T, 80] float32; op order of src/speech_featurizer.py:136-161."""
    p = p or yaml_params()
    if p.normalize_signal:
        x = x * (1.0 / (x.abs().max() + 1e-9))
    if p.preemphasis and p.preemphasis > 0:
        x = torch.cat([x[:1], x[1:] - p.preemphasis * x[:-1]])
    L, S = p.frame_length, p.frame_step
    if x.numel() < L:
        return torch.zeros((0, p.num_feature_bins), dtype=torch.float32)
    win = torch.from_numpy(hann_periodic(L))                       # rebuilt per call, like tf.signal.stft
    frames = x.unfold(0, L, S) * win
    X = torch.fft.rfft(frames, n=p.fft_length, dim=-1)             # zero padded at the tail
    P = X.abs().square()
    W = torch.from_numpy(htk_mel_matrix_f32(p.num_feature_bins, P.shape[-1], p.sample_rate,
                                            p.lower_edge_hertz, p.upper_edge_hertz))
    M = P @ W
    out = torch.log(torch.clamp_min(M, p.output_floor))
    return out / LN10 if p.log_base == "10" else out


def collate_torch(feats):
    T = max((f.shape[0] for f in feats), default=0)
    out = torch.zeros((len(feats), T, feats[0].shape[1], 1), dtype=torch.float32)
    n = torch.zeros(len(feats), dtype=torch.int32)
    for b, f in enumerate(feats):
        out[b, : f.shape[0], :, 0] = f
        n[b] = f.shape[0]
    return out, n


def subsample_torch(feat: torch.Tensor, n_frames: torch.Tensor, weights, activations=("tanh", "gelu", "gelu")):
    """feat [B,T,F,1]; weights list of numpy (dw[k,Cin], pw[Cin,Cout], b[Cout]).  Returns (out, mask, len3)."""
    h = feat.squeeze(-1).permute(0, 2, 1).contiguous()            # [B, C, T]
    L = n_frames.to(torch.float32)
    for (dw, pw, b), act in zip(weights, activations):
        dwt = torch.from_numpy(dw).t().unsqueeze(1).contiguous()   # [Cin,1,k]
        pwt = torch.from_numpy(pw).t().unsqueeze(-1).contiguous()  # [Cout,Cin,1]
        h = torch.nn.functional.conv1d(h, dwt, stride=2, groups=dwt.shape[0])
        h = torch.nn.functional.conv1d(h, pwt, torch.from_numpy(b))
        h = torch.tanh(h) if act == "tanh" else torch.nn.functional.gelu(h) if act == "gelu" else h
        L = torch.trunc((L - 9.0) / 2.0 + 1.0)
    len3 = L.to(torch.int32)
    width = max(int(len3.max()), 0) if len3.numel() else 0
    mask = (torch.arange(width)[None, :] < len3[:, None]).to(torch.float32)
    return h.permute(0, 2, 1).contiguous(), mask, len3


_POOLS = {}


def featurize_many_torch(wav: np.ndarray, lengths: np.ndarray, p: FeatParams | None = None, workers: int = 1):
    """The per-utterance featurizer calls of one batch.  workers > 1 runs them on a thread pool, one utterance per task,
    the way the reference's tf.data pipeline does (src/dataset.py:227: map(..., num_parallel_calls=AUTOTUNE)); torch
    releases the GIL inside its kernels, so the calls really overlap.  The caller sets torch's intra-op thread count
    (1 when the pool already covers the cores)."""
    p = p or yaml_params()
    xs = torch.from_numpy(wav)
    items = [xs[b, : int(lengths[b])] for b in range(wav.shape[0])]
    if workers <= 1:
        return [logmel_torch(x, p) for x in items]
    from concurrent.futures import ThreadPoolExecutor
    pool = _POOLS.get(workers)
    if pool is None:
        pool = _POOLS[workers] = ThreadPoolExecutor(max_workers=workers)
    return list(pool.map(lambda x: logmel_torch(x, p), items))


def frontend_torch(wav: np.ndarray, lengths: np.ndarray, weights, p: FeatParams | None = None, workers: int = 1,
                   conv_threads: int | None = None):
    """The whole reference-style CPU pass over one padded batch: per-utterance featurizer calls (optionally `workers` at
    a time, each single-threaded), zero-pad collate, then the three separable convs on the padded batch with
    `conv_threads` intra-op threads (default: leave torch's setting alone)."""
    if workers > 1:
        prev = torch.get_num_threads()
        torch.set_num_threads(1)
        try:
            feats = featurize_many_torch(wav, lengths, p, workers)
        finally:
            torch.set_num_threads(conv_threads or prev)
    else:
        feats = featurize_many_torch(wav, lengths, p, 1)
    feat, n = collate_torch(feats)
    return subsample_torch(feat, n, weights) + (feat, n)
