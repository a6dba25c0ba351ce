"""BASELINE.json configs[4] on one GPU: throughput of log-mel + subsampling over utterance length x batch
(equal-length utterances, 'tilt' distribution, device-resident inputs, CUDA-graph replay per step).  configs[3]
(30 s x 1024 sharded 1/8: 128 utterances per GPU) lies between two columns of the 30 s row.  Writes
gpurun_out/sweep.json and prints a markdown table.  (Product-side tool: it does not touch oracle/.)

    python tools/sweep.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, telugu_asr_b200 as tasr

dev = torch.device("cuda:0")
weights = bench.make_weights()
fe = tasr.FrontEnd(math="tf32"); fe.set_weights(weights, dev)
LENS = [1, 2, 5, 10, 15, 20, 30]
BATCHES = [1, 4, 16, 64, 256, 1024, 4096]
base, _ = tasr.synth.make_waveforms([480000] * 8, seed=100, dist="tilt") if hasattr(tasr, "synth") else (None, None)
if base is None:
    from telugu_asr_b200.synth import make_waveforms
    base, _ = make_waveforms([480000] * 8, seed=100, dist="tilt")
res = {}
for sec in LENS:
    n = sec * 16000
    for B in BATCHES:
        if B * n * 4 * 12 > 60e9:      # keep the working set (wav + features + activations) well inside HBM
            continue
        wav = torch.from_numpy(np.ascontiguousarray(np.tile(base[:, :n], (-(-B // 8), 1))[:B])).to(dev)
        lens = torch.full((B,), n, dtype=torch.int32, device=dev)
        cap = tasr.CapturedFrontEnd(fe, B, n, dev)
        cap.load(wav, lens)
        for _ in range(3): cap.replay()
        torch.cuda.synchronize()
        reps = max(3, min(50, int(0.2 / max(1e-4, B * sec / 5e6))))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): cap.replay()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[f"{sec}s x {B}"] = {"seconds": sec, "batch": B, "ms_per_step": ms, "audio_s_per_s": B * sec / (ms * 1e-3)}
        del cap, wav
        torch.cuda.empty_cache()
cpu = {}   # (the CPU figures quoted in profiles/r01_sweep.md come from bench.py's --impl reference arm on these shapes)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"gpu": res, "cpu": cpu}, open("gpurun_out/sweep.json", "w"), indent=1)
print("| length \\\\ batch | " + " | ".join(str(b) for b in BATCHES) + " |")
print("|---|" + "---:|" * len(BATCHES))
for sec in LENS:
    row = []
    for B in BATCHES:
        r = res.get(f"{sec}s x {B}")
        row.append(f"{r['audio_s_per_s'] / 1e6:.2f}" if r else "—")
    print(f"| {sec} s | " + " | ".join(row) + " |")
print("(M audio-seconds/s, one B200)")
for k, v in cpu.items():
    print(f"CPU port {k}: {v['audio_s_per_s']:.0f} audio-s/s on {v['cores']} threads")
