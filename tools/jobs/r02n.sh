#!/bin/bash
echo "== mid-size tests (wide input path)"
timeout 400 python -m pytest tests/test_gpu_api.py -x -q -k "mid_size" 2>&1 | grep -E "^E|passed|failed" | cut -c1-600 | head -12
echo "== xl bench"
timeout 600 python bench.py --config xl --steps 2 --warmup 3 --no-ttt > gpurun_out/r02n_bench_xl.json 2> gpurun_out/r02n_bench_xl.err; echo rc=$?; tail -c 500 gpurun_out/r02n_bench_xl.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/r02n_bench_xl.json"))
    print({k:j[k] for k in ("value","ms_per_step","kernels_launched","kernels_ms_per_call")}); print(j["e2e"]); print(j["cpu_baseline"])
    r=j["roofline"]; print(r["kernel"], r["frac"]); [print(k, round(v["ms_per_step"],2), v.get("frac"), v.get("frac_executed")) for k,v in r["kernels"].items()]
except Exception as e: print("xl ERR", e)
PY
