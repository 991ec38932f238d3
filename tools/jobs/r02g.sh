#!/bin/bash
# GPU job r02g: guarded first run of the 3-role XNODE backward, then the full gpu suite, benches and probes.
# Every command runs under its own `timeout`; the job stops at the first failure of the guard steps.
set -o pipefail
echo "== guard: one small golden case through the C ABI (3-role backward)"
timeout 180 python -m pytest tests/test_gpu_capi.py -x -q -k "cube_d5_shipped_small or cube_d3_rk4" 2>&1 | tail -5 || { echo "GUARD FAILED"; exit 1; }
echo "== full gpu suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== bench m (3-role)"
timeout 400 python bench.py --steps 3 --warmup 3 --no-ttt --no-cpu > gpurun_out/r02g_bench_v3.json 2> gpurun_out/r02g_bench_v3.err; echo rc=$?
echo "== bench m (2-role, A/B)"
XW_XNODE_BWD=v2 timeout 400 python bench.py --steps 3 --warmup 3 --no-ttt --no-cpu > gpurun_out/r02g_bench_v2.json 2> gpurun_out/r02g_bench_v2.err; echo rc=$?
python - <<'PY'
import json
for f in ("gpurun_out/r02g_bench_v3.json","gpurun_out/r02g_bench_v2.json"):
    try:
        j=json.load(open(f)); print(f, round(j["ms_per_step"],2), "%.4g"%j["value"], {k:round(v,2) for k,v in j["kernels_ms_per_call"].items()}, j["kernels_launched"]["xnode_generation"])
    except Exception as e: print(f, "ERR", e)
PY
echo "== small-N probe"
timeout 300 python tools/small_n_probe.py > gpurun_out/r02g_small_n.json 2> gpurun_out/r02g_small_n.prof; cat gpurun_out/r02g_small_n.json; grep -A34 "cumulative" gpurun_out/r02g_small_n.prof | cut -c1-140 | head -40
echo "== DRAM traffic of the v-net kernels (L2 hints)"
timeout 300 python bench.py --steps 1 --warmup 1 --no-ttt --no-cpu > gpurun_out/r02g_plain.json 2>gpurun_out/r02g_plain.err && timeout 600 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_vnet_tc -c 6 --csv --log-file gpurun_out/r02g_vnet_dram.csv python bench.py --steps 1 --warmup 1 --no-ttt --no-cpu > gpurun_out/r02g_ncu.log 2>&1
grep -v "^==" gpurun_out/r02g_vnet_dram.csv | cut -d, -f5,13- | head -30
echo "== xl bench"
timeout 600 python bench.py --config xl --steps 2 --warmup 3 --no-ttt > gpurun_out/r02g_bench_xl.json 2> gpurun_out/r02g_bench_xl.err; echo rc=$?
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/r02g_bench_xl.json"))
    print({k:j[k] for k in ("value","ms_per_step","kernels_launched","kernels_ms_per_call")}); print(j["e2e"]); print(j["cpu_baseline"])
except Exception as e: print("xl ERR", e)
PY
