import os, sys
sys.path.insert(0, "/root/repo")
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
def run(nstreams, K=200):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    for i in range(10):
        with torch.cuda.stream(streams[i % nstreams]): caps[i % nstreams].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_event(e0)
    for i in range(K):
        with torch.cuda.stream(streams[i % nstreams]): caps[i % nstreams].replay()
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
for n in (1, 2, 3):
    print(n, "streams:", "%.1f us/step" % (run(n) * 1e3))
