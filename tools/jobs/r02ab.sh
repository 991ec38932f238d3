#!/bin/bash
# ABI v3 (general coefficients) + hourglass bound_pad: full GPU suite, bench
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu --no-ttt > gpurun_out/r02ab_bench.json 2> gpurun_out/r02ab_bench.err; echo rc=$?
python - <<'PY'
import json
for f in ("r02ab_bench",):
    try:
        j=json.load(open("gpurun_out/%s.json" % f)); print(f, round(j["ms_per_step"],2), "%.4g"%j["value"], j["e2e"]["ms_per_step"], {k:round(v,2) for k,v in j["kernels_ms_per_call"].items()})
    except Exception as e: print(f, "ERR", e)
PY
