#!/bin/bash
timeout 300 python tools/small_n_probe.py > gpurun_out/r02an_probe.json 2> gpurun_out/r02an_probe.err; echo rc=$?
cat gpurun_out/r02an_probe.json
grep -v "^$" gpurun_out/r02an_probe.err | head -60 | cut -c1-200
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-ttt --nlc-max-gb 0 > gpurun_out/r02an_bench.json 2>gpurun_out/r02an_bench.err; echo rc=$?
python - <<'PY'
import json
t=open("gpurun_out/r02an_bench.json").read(); j=json.loads(t[t.index('{"metric'):]); print(json.dumps(j["phases"]))
PY
