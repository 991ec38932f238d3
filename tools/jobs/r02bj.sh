#!/bin/bash
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/r02bj_bench.json 2> gpurun_out/r02bj_bench.err; echo "bench rc=$? $(( $(date +%s)-S )) s"
python - <<'PY'
import json
t=open("gpurun_out/r02bj_bench.json").read(); j=json.loads(t[t.index('{"metric'):])
print(round(j["ms_per_step"],2), "%.4g"%j["value"], "e2e", j["e2e"]["ms_per_step"], "frac", j["roofline"]["frac"], j["clocks"], j["gpu_launches"])
print(j["variants"])
print({k:v for k,v in j["time_to_target"].items() if k in ("sub_iters","seconds","steady_ms_per_outer_iter")}, j["cpu_baseline"]["value"])
PY
tail -2 gpurun_out/r02bj_bench.err | cut -c1-200
