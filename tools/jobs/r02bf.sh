#!/bin/bash
for L2N in 17 20; do for F in 0 1; do for G in 0 1; do
XW_CONCURRENT_BOUNDARY=$F timeout 300 python bench.py --log2n $L2N --steps 10 --warmup 5 --no-cpu --no-ttt --nlc-max-gb 0 --graph $G > gpurun_out/r02bf_tmp.json 2> gpurun_out/r02bf_tmp.err
python - $L2N $F $G <<'PY'
import json, sys
t=open("gpurun_out/r02bf_tmp.json").read(); j=json.loads(t[t.index('{"metric'):]); print("log2n", sys.argv[1], "concurrent", sys.argv[2], "graph", sys.argv[3], round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3))
PY
done; done; done
