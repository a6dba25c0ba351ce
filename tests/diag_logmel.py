"""GPU diagnostic: where does the log-mel kernel deviate most from the float64 oracle on the
bench workload?  (development aid, run under gpurun)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle, telugu_asr_b200 as tasr
from telugu_asr_b200.synth import draw_lengths

dev = torch.device("cuda:0")
lens = draw_lengths(256, 16000, 240000, seed=2)
wav, ln = oracle.make_waveforms(lens, seed=2, dist="tilt")
feat = tasr.SpeechFeaturizer(**tasr.REFERENCE_SPEECH_CONFIG)
out, nf = feat(torch.from_numpy(wav).to(dev), torch.from_numpy(ln).to(dev))
out2, _ = feat(torch.from_numpy(wav).to(dev), torch.from_numpy(ln).to(dev))
print("run-to-run identical:", torch.equal(out, out2))
o = out.cpu().numpy()[..., 0]
rows = []
for b in range(256):
    r64 = oracle.logmel_ref(wav[b, :ln[b]], dtype=np.float64)
    r32 = oracle.logmel_ref(wav[b, :ln[b]], dtype=np.float32)
    T = r64.shape[0]
    e = np.abs(o[b, :T] - r64)
    t, m = np.unravel_index(e.argmax(), e.shape)
    rows.append((e.max(), b, int(ln[b]), T, int(t), int(m), o[b, t, m], r64[t, m], r32[t, m], np.abs(r32 - r64).max()))
rows.sort(reverse=True)
for r in rows[:12]:
    print("err %.3e b=%d len=%d T=%d t=%d m=%d gpu=%.6f ref64=%.6f ref32=%.6f | f32-oracle band %.2e" % r)
print("median per-utt max err %.3e" % np.median([r[0] for r in rows]))
# alone vs in batch
b = rows[0][1]
alone = feat(torch.from_numpy(wav[b, :ln[b]].copy()).to(dev)).cpu().numpy()
print("alone == batch:", np.array_equal(alone, o[b, :alone.shape[0]]), np.abs(alone - o[b, :alone.shape[0]]).max())
