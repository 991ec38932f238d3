#!/bin/bash
# three-warp MMA issue in the forward tcgen05 kernel (default): full GPU suite in both modes, bench
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
XW_TC_SPLIT=0 timeout 600 python -m pytest tests/test_gpu_capi.py tests/test_gpu_tc.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -3
timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu --no-ttt > gpurun_out/r02ae_bench.json 2> gpurun_out/r02ae_bench.err; echo rc=$?
python - <<'PY'
import json
for f in ("r02ae_bench",):
    try:
        j=json.load(open("gpurun_out/%s.json" % f)); print(f, round(j["ms_per_step"],2), "%.4g"%j["value"], j["e2e"]["ms_per_step"], {k:round(v,2) for k,v in j["kernels_ms_per_call"].items()})
    except Exception as e: print(f, "ERR", e)
PY
