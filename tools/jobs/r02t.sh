#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_capi.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-ttt --no-cpu > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo rc=$?
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02t_bench.json")); print(round(j["ms_per_step"],2), "%.4g"%j["value"], {k:round(v,2) for k,v in j["kernels_ms_per_call"].items()})
PY
