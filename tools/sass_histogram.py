"""Per-kernel SASS opcode histogram of the in-tree library (cuobjdump -sass), written as tracked evidence:
  profiles/<name>_sass.md   one table per kernel: registers / shared / spills from -res-usage, the Blackwell-native
                            opcodes (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor
                            load/store, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier, FFMA2/FADD2/FMUL2 =
                            packed FP32x2) and the 12 most frequent opcodes.
Usage: python tools/sass_histogram.py NAME   (no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "telugu_asr_b200", "libtasr_b200.so")
name = sys.argv[1] if len(sys.argv) > 1 else "sass"
NATIVE = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "USETMAXREG",
          "FFMA2", "FADD2", "FMUL2", "MUFU", "REDUX", "SHFL", "LDGSTS", "HMMA", "F2FP"]


def demangle(sym):
    try:
        return subprocess.run(["cu++filt", sym], capture_output=True, text=True).stdout.strip() or sym
    except OSError:
        return sym


def short(k):
    k = k.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    k = re.sub(r"\((int|bool|unsigned int)\)", "", k)
    return k.split("(")[0]


sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True, check=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in line:
        usage[cur] = line.strip()
        cur = None

kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Za-z0-9_]+)*)", line)
    if m:
        kernels[cur][m.group(1)] += 1
        if m.group(1) in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "MUFU"):
            kernels[cur][m.group(1) + m.group(2)] += 0   # keep the plain key only; variants listed below
            kernels[cur]["~" + m.group(1) + m.group(2)] += 1

out = [f"# {name} — SASS opcode histograms of `telugu_asr_b200/libtasr_b200.so` (sm_100a)",
       "",
       "`python tools/sass_histogram.py " + name + "` = `cuobjdump -sass` + `cuobjdump -res-usage` of the in-tree library, counted per kernel.",
       "Static instruction counts (one per SASS line), not executed counts.  Legend: UTCHMMA = `tcgen05.mma`, LDTM / STTM = `tcgen05.ld / st`,",
       "UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = `cp.async.bulk`, UTCBAR = `tcgen05.commit`, SYNCS = mbarrier ops,",
       "USETMAXREG = `setmaxnreg`, FFMA2 / FADD2 / FMUL2 = packed FP32x2.", ""]
out.append("| kernel | resources | Blackwell-native / notable opcodes | most frequent opcodes |")
out.append("|---|---|---|---|")
for sym, cnt in kernels.items():
    total = sum(v for k, v in cnt.items() if not k.startswith("~"))
    nat = ", ".join(f"{k} {cnt[k]}" for k in NATIVE if cnt.get(k))
    var = ", ".join(f"{k[1:]} {v}" for k, v in sorted(cnt.items()) if k.startswith("~") and not k.startswith("~MUFU"))
    mufu = ", ".join(f"{k[1:]} {v}" for k, v in sorted(cnt.items()) if k.startswith("~MUFU"))
    top = ", ".join(f"{k} {v}" for k, v in cnt.most_common(40) if not k.startswith("~"))
    top = ", ".join(top.split(", ")[:12])
    u = usage.get(sym, "")
    u = re.sub(r"\s+", " ", u)
    u = re.sub(r" (STACK:0|LOCAL:0|TEXTURE:0|SURFACE:0|SAMPLER:0)", "", u)
    extra = "; ".join(x for x in (var, mufu) if x)
    out.append(f"| `{short(demangle(sym))}` ({total} instr.) | {u} | {nat or '—'}{' — ' + extra if extra else ''} | {top} |")
path = os.path.join(ROOT, "profiles", f"{name}_sass.md")
open(path, "w").write("\n".join(out) + "\n")
print(path, len(kernels), "kernels")
