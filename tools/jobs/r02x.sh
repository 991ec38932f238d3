#!/bin/bash
for v in 0 1; do
TCB_SUFFIX=_l$v XW_TC_FLUSH=4 timeout 300 python tools/tc_prof.py 20 20 > gpurun_out/r02x_l$v.json 2>&1
python - $v <<'PY'
import json, sys
t=open('gpurun_out/r02x_l%s.json' % sys.argv[1]).read()
try:
    d=json.loads(t[t.index('{'):])
    print("late", sys.argv[1], d['ms_per_call'], d['ms_per_call_instrumented'])
    for k,v in d['roles'].items():
        if "warp" in k: print(k, v['cycles_per_tile'], {a[:18]:round(b*100,1) for a,b in v['wait_share'].items()})
except Exception as e: print("ERR", e, t[-1500:])
PY
done
