"""Development aid: timeline of CTA 0 of logmel_tc_kernel (roles: stager warp 0, MMA issuer, producer warp 0,
consumer warp 0).  TASR_LOGMEL_TC=1 python tools/tc_trace.py [n_events]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import oracle, telugu_asr_b200 as tasr
from telugu_asr_b200 import _native
dev = torch.device("cuda:0")
feat = tasr.SpeechFeaturizer(**tasr.REFERENCE_SPEECH_CONFIG)
lens = tasr.synth.draw_lengths(256, 16000, 240000, seed=2)
wav, ln = oracle.make_waveforms(lens, seed=2, dist="tilt")
w, l = torch.from_numpy(wav).to(dev), torch.from_numpy(ln).to(dev)
L = _native.lib()
L.tasr_debug_tc_trace.argtypes = [C.c_void_p]; L.tasr_debug_tc_trace.restype = None
for _ in range(3): feat(w, l)
buf = torch.zeros(6 * 512, dtype=torch.int64, device=dev)
L.tasr_debug_tc_trace(buf.data_ptr()); feat(w, l); torch.cuda.synchronize(); L.tasr_debug_tc_trace(None)
t = buf.cpu().numpy().reshape(6, 256, 2)
t0 = t[5, 0, 1]
print(f"n_v={t[5,0,0]}")
names = ["stg", "mma", "prd", "con"]
what = {("stg", 1): "start", ("stg", 2): "done", ("mma", 1): "afull", ("mma", 2): "issued", ("prd", 1): "wfull", ("prd", 2): "aempty", ("prd", 3): "done",
        ("con", 1): "accf", ("con", 2): "P done", ("con", 3): "mel done", ("con", 4): "stored", ("con", 9): "end"}
ev = []
for r in range(4):
    for tag, ts in t[r]:
        if tag: ev.append(((ts - t0) / 1e3, names[r], int(tag)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
for ts, nm, tag in sorted(ev)[:n]: print(f"{ts:9.2f}  {nm:4s} tile {tag % 1000:3d} {what.get((nm, tag // 1000), tag)}")
print("...")
for ts, nm, tag in sorted(ev)[-12:]: print(f"{ts:9.2f}  {nm:4s} tile {tag % 1000:3d} {what.get((nm, tag // 1000), tag)}")
