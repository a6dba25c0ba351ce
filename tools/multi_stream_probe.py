"""Development aid: time per step of the captured front end with 1..3 steps in flight, with the streams in phase or
offset by half a step (stream i starts after i extra featurizer passes).  python tools/multi_stream_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, telugu_asr_b200 as tasr
dev = torch.device("cuda:0")
wav_np, lens_np = bench.make_batch(0, 256)
fe = tasr.FrontEnd(math="tf32"); fe.set_weights(bench.make_weights(), dev)
wav = torch.from_numpy(wav_np).to(dev); lens = torch.from_numpy(lens_np).to(dev)
caps = []
for i in range(3):
    c = tasr.CapturedFrontEnd(fe, 256, wav.shape[1], dev); c.load(wav, lens); caps.append(c)
torch.cuda.synchronize()
def run(nstreams, offset, K=200):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    for i in range(10):
        with torch.cuda.stream(streams[i % nstreams]): caps[i % nstreams].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_event(e0)
    if offset:
        for i in range(1, nstreams):
            with torch.cuda.stream(streams[i]):
                for _ in range(i): fe.featurizer.featurize_batch(wav, lens)     # ~half a step of delay per stream index
    for i in range(K):
        with torch.cuda.stream(streams[i % nstreams]): caps[i % nstreams].replay()
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
for n in (1, 2, 3):
    print(n, "streams: in phase %.1f us/step, offset %.1f us/step" % (run(n, False) * 1e3, run(n, True) * 1e3))
