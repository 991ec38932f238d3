#!/bin/bash
echo "== guard"
timeout 180 python -m pytest tests/test_gpu_capi.py -x -q -k "cube_d5_shipped_small or cube_d3_rk4 or cube_d3_euler" 2>&1 | tail -3 || exit 1
echo "== full gpu suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== bench m"
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo rc=$?; tail -c 300 gpurun_out/r02j_bench.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/r02j_bench.json")); print(round(j["ms_per_step"],2), "%.4g"%j["value"], {k:round(v,2) for k,v in j["kernels_ms_per_call"].items()}); print(j["e2e"]); print(j["time_to_target"]["sub_iters"], j["time_to_target"]["seconds"], j["time_to_target"]["ms_per_sub_iter"])
    r=j["roofline"]; print(r["kernel"], r["frac"]); [print(k, round(v["ms_per_step"],2), v.get("frac")) for k,v in r["kernels"].items()]
except Exception as e: print("ERR", e)
PY
