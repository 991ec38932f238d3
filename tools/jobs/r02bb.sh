#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
PROBE_SYNC=0 timeout 300 python tools/ttt_probe.py 2>&1 | tail -1
XW_DEFER_BOUNDARY_JOIN=0 PROBE_SYNC=0 timeout 300 python tools/ttt_probe.py 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --nlc-max-gb 0 > gpurun_out/r02bb_bench.json 2>gpurun_out/r02bb_bench.err; echo rc=$?
python - <<'PY'
import json
t=open("gpurun_out/r02bb_bench.json").read(); j=json.loads(t[t.index('{"metric'):]); print(j["ms_per_step"], j["e2e"]["ms_per_step"], j["gpu_launches"], {k:v for k,v in j["time_to_target"].items() if k in ("sub_iters","seconds","ms_per_sub_iter","steady_ms_per_outer_iter")})
PY
tail -2 gpurun_out/r02bb_bench.err | cut -c1-300
