#!/bin/bash
timeout 800 python bench.py --config xl --steps 2 --warmup 3 --no-ttt --no-cpu > gpurun_out/r02ai_bench_xl.json 2> gpurun_out/r02ai_bench_xl.err; echo rc=$?
python - <<'PY'
import json
t=open("gpurun_out/r02ai_bench_xl.json").read(); j=json.loads(t[t.index('{"metric'):]); print(round(j["ms_per_step"],2), "%.4g"%j["value"], j["e2e"]["ms_per_step"], {k:round(v,2) for k,v in j["kernels_ms_per_call"].items()}, j["kernels_launched"])
PY
tail -3 gpurun_out/r02ai_bench_xl.err | cut -c1-300
