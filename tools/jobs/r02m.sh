#!/bin/bash
echo "== full gpu suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench m"
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; echo rc=$?; tail -c 300 gpurun_out/r02m_bench.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/r02m_bench.json")); print(round(j["ms_per_step"],2), "%.4g"%j["value"], {k:round(v,2) for k,v in j["kernels_ms_per_call"].items()}); print(j["e2e"]); print(j["e2e_nlc"]); print(j["kernels_launched"])
except Exception as e: print("ERR", e)
PY
echo "== launch list"
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu --no-ttt > gpurun_out/r02m_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 500 --csv --log-file gpurun_out/r02m_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-ttt > gpurun_out/r02m_ncu.log 2>&1; wc -l gpurun_out/r02m_launches.csv
