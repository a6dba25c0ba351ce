"""Development aid: timeline of CTA 0 of the persistent separable-conv kernel (roles: dw warp 0/7, MMA issuer,
epilogue warp 0/7).  python tools/ws_trace.py [layer 1..3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import bench, telugu_asr_b200 as tasr
from telugu_asr_b200 import _native
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
wav_np, lens_np = bench.make_batch(0, 256)
fe = tasr.FrontEnd(math="tf32"); fe.set_weights(bench.make_weights(), dev)
wav = torch.from_numpy(wav_np).to(dev); lens = torch.from_numpy(lens_np).to(dev)
L = _native.lib()
L.tasr_debug_ws_trace.argtypes = [C.c_void_p]; L.tasr_debug_ws_trace.restype = None
feats, nf = fe.featurizer.featurize_batch(wav, lens)
sub = fe.subsampling; sub._ensure_plans()
h = feats.reshape(256, -1, 80); t_in = h.shape[1]
for i in range(3):
    t_out = (t_in - 9) // 2 + 1
    y = torch.empty((256, t_out, sub.filters[i]), device=dev)
    st = _native.stream_ptr()
    call = lambda: _native.check(L.tasr_sepconv1d_tf32_ragged(sub._plans[i], h.data_ptr(), nf.data_ptr(), i, 256, t_in, y.data_ptr(), t_out, st))
    for _ in range(3): call()
    if i == layer - 1:
        buf = torch.zeros(6 * 512, dtype=torch.int64, device=dev)
        L.tasr_debug_ws_trace(buf.data_ptr()); call(); torch.cuda.synchronize(); L.tasr_debug_ws_trace(None)
        t = buf.cpu().numpy().reshape(6, 256, 2)
        t0 = t[5, 0, 1]
        print(f"layer {layer}: n_c={t[5,1,0]} n_f={t[5,1,1]}  prologue {(t0 - t[5,2,0]) / 1e3:.2f} us, CTA lifetime {(t[5,2,1] - t[5,2,0]) / 1e3:.2f} us (times in us after the prologue)")
        names = ["dw0", "xld", "mma", "ep0", "ep7"]
        ev = []
        for r in range(5):
            for tag, ts in t[r]:
                if tag: ev.append(((ts - t0) / 1e3, names[r], int(tag)))
        for ts, n, tag in sorted(ev): print(f"{ts:9.2f}  {n:4s} {tag}")
    h, t_in = y, t_out
