#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_api.py -x -q -k "graph or realloc or prefetch" 2>&1 | tail -3
timeout 300 python tools/small_n_probe.py > gpurun_out/r02u_small_n.json 2> gpurun_out/r02u_small_n.prof; cat gpurun_out/r02u_small_n.json; grep -A16 "cumulative" gpurun_out/r02u_small_n.prof | cut -c1-140 | head -20
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo rc=$?
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02u_bench.json")); t=j["time_to_target"]; print(round(j["ms_per_step"],2), t["sub_iters"], t["seconds"], t["ms_per_sub_iter"])
PY
