#!/bin/bash
timeout 900 python bench.py --config xl --steps 2 --warmup 3 --no-ttt --no-cpu > gpurun_out/r02ay_bench_xl.json 2> gpurun_out/r02ay_bench_xl.err; echo rc=$?
python - <<'PY'
import json
t=open("gpurun_out/r02ay_bench_xl.json").read(); j=json.loads(t[t.index('{"metric'):]); print(j["ms_per_step"], "%.4g"%j["value"], j["e2e"]["ms_per_step"], j["kernels_ms_per_call"], j["variants"])
PY
tail -3 gpurun_out/r02ay_bench_xl.err | cut -c1-300
