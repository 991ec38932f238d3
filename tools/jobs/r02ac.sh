#!/bin/bash
# forward tcgen05 kernel: one issuing thread vs three issuing warps
XW_TC_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_capi.py tests/test_gpu_tc.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3
for sp in 0 1; do
XW_TC_SPLIT=$sp timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu --no-ttt > gpurun_out/r02ac_bench_sp$sp.json 2> gpurun_out/r02ac_bench_sp$sp.err; echo rc=$?
done
python - <<'PY'
import json
for f in ("r02ac_bench_sp0","r02ac_bench_sp1"):
    try:
        j=json.load(open("gpurun_out/%s.json" % f)); print(f, round(j["ms_per_step"],2), "%.4g"%j["value"], j["e2e"]["ms_per_step"], {k:round(v,2) for k,v in j["kernels_ms_per_call"].items()})
    except Exception as e: print(f, "ERR", e)
PY
