#!/bin/bash
# Development aid: one short bench run, printed as a single brief line (value, step time, stage times).
python bench.py --steps ${STEPS:-100} --warmup 10 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c '
import json,sys
d=json.loads(sys.stdin.read())
st=d.get("stages",{})
print("value %.3f M  step %.1f us  e2e %.3f M | " % (d["value"]/1e6, d["ms_per_step"]*1e3, d["e2e"]["value"]/1e6) + "  ".join("%s %.1f" % (k.replace("sepconv_layer","L").replace("_kernel",""), v["us"]) for k,v in st.items()))
'
