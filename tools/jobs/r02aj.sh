#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu --no-ttt > gpurun_out/r02aj_bench.json 2> gpurun_out/r02aj_bench.err; echo rc=$?
python - <<'PY'
import json
t=open("gpurun_out/r02aj_bench.json").read(); j=json.loads(t[t.index('{"metric'):]); print(round(j["ms_per_step"],2), "%.4g"%j["value"], j["e2e"]["ms_per_step"], {k:round(v,2) for k,v in j["kernels_ms_per_call"].items()}, j["variants"])
PY
tail -3 gpurun_out/r02aj_bench.err | cut -c1-300
