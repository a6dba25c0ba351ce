"""Device time of Conv2dSubsampling (conformer front end, filters 144) on the bench batch's feature shape
[256, 1498, 80, 1].  Development aid, run under gpurun:  python tools/time_conv2d.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import telugu_asr_b200 as tasr

dev = torch.device("cuda:0")
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 1498
layer = tasr.Conv2dSubsampling({"filters": 144, "kernel_size": 3, "strides": 2, "padding": "same"}, seed=1)
layer.build(dev)
x = torch.randn((B, T, 80, 1), device=dev)
ln = torch.full((B,), T, dtype=torch.int32, device=dev)
for _ in range(3):
    out, _ = layer([x, ln])
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out, _ = layer([x, ln]); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
h1, w1, h2, w2 = layer.output_shape(T, 80)
flop = 2.0 * B * h2 * w2 * 144 * 9 * 144 + 2.0 * B * h1 * w1 * 144 * 9
audio_s = B * (400 + 160 * (T - 1)) / 16000
print(f"Conv2dSubsampling dense B={B} T={T}: med {ts[len(ts)//2]:.3f} ms  min {ts[0]:.3f} ms; out {tuple(out.shape)}; "
      f"{flop / ts[len(ts)//2] / 1e9:.1f} TFLOP/s; {audio_s / ts[len(ts)//2] * 1e3 / 1e6:.2f} M audio-s/s for this stage")

# the bench workload's ragged lengths (configs[2]: 1..15 s, zero padded to 15 s): ragged mode fills / skips the padding
import bench
_, lens_np = bench.make_batch(0, B)
nf = np.maximum(0, 1 + (lens_np.astype(np.int64) - 400) // 160).astype(np.int32)
xr = x.clone()
for b in range(B):
    xr[b, nf[b]:] = 0.0
ln = torch.from_numpy(nf).to(dev)
for mode, layer_r in (("dense", tasr.Conv2dSubsampling({"filters": 144}, seed=1, assume_zero_padding=False)),
                      ("ragged", tasr.Conv2dSubsampling({"filters": 144}, seed=1, assume_zero_padding=True))):
    layer_r.build(dev)
    for _ in range(3):
        out, _ = layer_r([xr, ln])
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out, _ = layer_r([xr, ln]); b_.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b_))
    ts.sort()
    real_audio = float(lens_np.sum()) / 16000
    print(f"configs[2] lengths ({real_audio:.0f} real audio-s, {100 * (1 - nf.sum() / (B * T)):.0f} % padding), {mode}: "
          f"med {ts[len(ts)//2]:.3f} ms -> {real_audio / ts[len(ts)//2] * 1e3 / 1e6:.2f} M audio-s/s")

# the whole conformer front end on the bench batch: waveforms -> log-mel (single pass, lean) -> Conv2dSubsampling (ragged)
wav_np, lens_np = bench.make_batch(0, B)
wav_d, len_d = torch.from_numpy(wav_np).to(dev), torch.from_numpy(lens_np).to(dev)
cfe = tasr.ConformerFrontEnd(seed=1)
cfe.subsampling.build(dev)
for _ in range(3):
    cfe(wav_d, len_d, max_length=int(lens_np.max()))
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); cfe(wav_d, len_d, max_length=int(lens_np.max())); b_.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b_))
ts.sort()
print(f"ConformerFrontEnd (waveform -> [B, T/4, 2880]) on configs[2]'s batch: med {ts[len(ts)//2]:.3f} ms -> "
      f"{float(lens_np.sum()) / 16000 / ts[len(ts)//2] * 1e3 / 1e6:.2f} M audio-s/s (eager launches, one stream)")
