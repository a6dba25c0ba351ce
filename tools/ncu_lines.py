"""Join an ncu SASS source-page CSV with nvdisasm -g line info: per CUDA source line, the share
of executed warp instructions and of stall samples.  Development aid.

  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > sass.csv
  cuobjdump -xelf all lib.so; nvdisasm -g -c file.cubin > disasm.txt
  python tools/ncu_lines.py sass.csv disasm.txt KERNEL_SUBSTR source.cu
"""
import csv, re, sys, collections

sass_csv, disasm, kname, srcfile = sys.argv[1:5]
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
addrs = [int(r[ix['Address']], 16) if r[ix['Address']].startswith('0x') else int(r[ix['Address']]) for r in data]
base = min(addrs)
per_off = {a - base: (int(r[ix['Instructions Executed']] or 0), int(r[ix['# Samples']] or 0), r[ix['Source']]) for a, r in zip(addrs, data)}

# parse disasm: track current line within the kernel section
line_of = {}
cur = None; inside = False
for l in open(disasm):
    if l.startswith('//---') :
        inside = kname in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0])
for off, (n, s, src) in per_off.items():
    k = line_of.get(off)
    agg[k][0] += n; agg[k][1] += s
tot_n = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
src = open(srcfile).read().splitlines()
print(f"total warp-inst {tot_n}  samples {tot_s}")
for k, (n, s) in sorted(agg.items(), key=lambda kv: (kv[0] is None, kv[0])):
    if n < 0.003 * tot_n and s < 0.003 * tot_s: continue
    text = ''
    if k and k[0] == srcfile.split('/')[-1] and k[1] - 1 < len(src): text = src[k[1] - 1].strip()[:90]
    print(f"{str(k):28s} inst {100*n/tot_n:5.1f}%  smp {100*s/tot_s:5.1f}% | {text}")
