"""Per-stage device times of the hot path on the bench workload (CUDA events, one stage at a time,
inputs larger than L2).  Development aid, run under gpurun:  python tools/time_stages.py [tf32|fp32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import bench, telugu_asr_b200 as tasr
from telugu_asr_b200 import _native

math = sys.argv[1] if len(sys.argv) > 1 else "tf32"
dev = torch.device("cuda:0")
wav_np, lens_np = bench.make_batch(0, 256)
fe = tasr.FrontEnd(math=math); fe.set_weights(bench.make_weights(), dev)
wav = torch.from_numpy(wav_np).to(dev); lens = torch.from_numpy(lens_np).to(dev)
max_len = int(lens_np.max())
feat = fe.featurizer; sub = fe.subsampling
L = _native.lib()

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

t_max = feat.get_nframes(max_len)
feats, nf = feat.featurize_batch(wav, lens, t_max=t_max)
print("featurize_batch (absmax+logmel): med %.1f us  min %.1f us" % timeit(lambda: feat.featurize_batch(wav, lens, t_max=t_max)))
if math == "tf32": sub._ensure_plans()
h = feats.reshape(256, t_max, 80)
t_in = t_max
for i in range(3):
    t_out = (t_in - 9) // 2 + 1
    y = torch.empty((256, t_out, sub.filters[i]), dtype=torch.float32, device=dev)
    st = _native.stream_ptr()
    if math == "tf32" and "dense" not in sys.argv:
        f = lambda h=h, y=y, t_in=t_in, t_out=t_out, i=i: _native.check(L.tasr_sepconv1d_tf32_ragged(sub._plans[i], h.data_ptr(), nf.data_ptr(), i, 256, t_in, y.data_ptr(), t_out, st))
    elif math == "tf32":
        f = lambda h=h, y=y, t_in=t_in, t_out=t_out, i=i: _native.check(L.tasr_sepconv1d_tf32(sub._plans[i], h.data_ptr(), 256, t_in, y.data_ptr(), t_out, st))
    else:
        ls = sub._layer_struct(i)
        f = lambda h=h, y=y, t_in=t_in, t_out=t_out, ls=ls: _native.check(L.tasr_sepconv1d_f32(h.data_ptr(), 256, t_in, C.byref(ls), y.data_ptr(), t_out, st))
    med, mn = timeit(f)
    byts = (h.numel() + y.numel()) * 4
    print(f"sepconv layer {i+1} [{t_in}x{h.shape[-1]} -> {t_out}x{sub.filters[i]}]: med {med:.1f} us  min {mn:.1f} us   {byts/med/1e3:.0f} GB/s of x+y")
    h, t_in = y, t_out
print("full step: med %.1f us  min %.1f us" % timeit(lambda: fe(wav, lens, max_length=max_len)))
