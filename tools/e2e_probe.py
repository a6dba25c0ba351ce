"""Development aid: e2e pipeline step time with the D2H leg on / off, to see how much the two PCIe directions
interfere.  python tools/e2e_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, telugu_asr_b200 as tasr
from telugu_asr_b200.synth import to_pcm16
dev = torch.device("cuda:0")
wav_np, lens_np = bench.make_batch(0, 256)
fe = tasr.FrontEnd(math="tf32"); fe.set_weights(bench.make_weights(), dev)
utts = [to_pcm16(wav_np[b, : lens_np[b]]) for b in range(256)]
pipe = tasr.FrontEndPipeline(fe, 256, wav_np.shape[1], dev, pcm16=True, slots=2, graph=False)
for s in range(2): pipe.stage(s, utts)
def run(K=100):
    for i in range(4): tk = pipe.submit(i % 2)
    tk.wait(); torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(K): tk = pipe.submit(i % 2)
    pipe.s_out.synchronize(); g1.record(); torch.cuda.synchronize()
    return g0.elapsed_time(g1) / K
print("with D2H: %.3f ms/step" % run())
# monkeypatch: skip the D2H copies
import telugu_asr_b200.pipeline as P
orig = torch.Tensor.copy_
class NoCopy:
    pass
def submit_nod2h(self, slot):
    pb = self.staging[slot]
    with torch.cuda.device(self.device):
        comp = torch.cuda.current_stream()
        if self._used[slot]: self.s_in.wait_event(self.ev_free[slot])
        with torch.cuda.stream(self.s_in):
            pb.to_device(non_blocking=True); self.ev_ready[slot].record(self.s_in)
        comp.wait_event(self.ev_ready[slot])
        wav, lens = pb.unpack(); self.ev_free[slot].record(comp)
        out, mask, len3 = self.fe(wav, lens, max_length=pb.max_len)
        self.ev_comp[slot].record(comp)
        self.s_out.wait_event(self.ev_comp[slot])
        with torch.cuda.stream(self.s_out): self.ev_done[slot].record(self.s_out)
    self._used[slot] = True
    return P.Ticket(slot, self.ev_done[slot], None, None, None)
P.FrontEndPipeline.submit = submit_nod2h
print("without D2H: %.3f ms/step" % run())
