#!/bin/bash
# stacked P-op of k_vnet_tc_bwd3: parity (both tensor-memory layouts) + bench
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
XW_TC_TMEM_PACKED=1 timeout 600 python -m pytest tests/test_gpu_capi.py tests/test_gpu_tc.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -5
timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu --no-ttt > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err; echo rc=$?
XW_TC_TMEM_PACKED=1 timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu --no-ttt > gpurun_out/r02w_bench_packed.json 2> gpurun_out/r02w_bench_packed.err; echo rc=$?
python - <<'PY'
import json
for f in ("r02w_bench", "r02w_bench_packed"):
    try:
        j=json.load(open("gpurun_out/%s.json" % f)); print(f, round(j["ms_per_step"],2), "%.4g"%j["value"], j["e2e"]["ms_per_step"], {k:round(v,2) for k,v in j["kernels_ms_per_call"].items()})
    except Exception as e: print(f, "ERR", e)
PY
