#!/bin/bash
# re-entry check of the restored tree: full GPU suite + the bench line
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tail -9
timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo rc=$?
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02v_bench.json")); print(round(j["ms_per_step"],2), "%.4g"%j["value"], j["e2e"]["ms_per_step"], {k:round(v,2) for k,v in j["kernels_ms_per_call"].items()})
PY
