#!/bin/bash
for g in 0 1; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-ttt --no-cpu --graph $g > gpurun_out/r02q_bench_g$g.json 2> gpurun_out/r02q_bench_g$g.err; echo rc=$?; tail -c 300 gpurun_out/r02q_bench_g$g.err
done
python - <<'PY'
import json
for g in (0,1):
    try:
        j=json.load(open("gpurun_out/r02q_bench_g%d.json"%g)); print(g, round(j["ms_per_step"],2), "%.4g"%j["value"], j["e2e"]["ms_per_step"], j["gpu_launches"], {k:round(v,2) for k,v in j["kernels_ms_per_call"].items()})
    except Exception as e: print(g, "ERR", e)
PY
