#!/bin/bash
for fl in 4; do
XW_TC_FLUSH=$fl timeout 300 python tools/tc_prof.py 20 20 > gpurun_out/r02x_tc_prof_fl$fl.json 2>&1
python - $fl <<'PY'
import json, sys
t=open('gpurun_out/r02x_tc_prof_fl%s.json' % sys.argv[1]).read()
try:
    d=json.loads(t[t.index('{'):])
    print("flush", sys.argv[1], d['ms_per_call'], d['ms_per_call_instrumented'])
    for k,v in d['roles'].items(): print(k, v['cycles_per_tile'], {a[:18]:round(b*100,1) for a,b in v['wait_share'].items()})
except Exception as e: print("ERR", e, t[-2000:])
PY
done
timeout 300 python tools/tc_prof.py acc 18 20 2>&1 | grep -A8 "flush_every_4"
