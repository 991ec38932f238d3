#!/bin/bash
# GPU job r02e: gpu tests, xl bench at N=1, DRAM traffic of the tcgen05 v-net kernels (L2 hints)
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --config xl --steps 2 --warmup 3 --no-ttt > gpurun_out/r02e_bench_xl.json 2> gpurun_out/r02e_bench_xl.err; echo rc=$?
tail -c 600 gpurun_out/r02e_bench_xl.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02e_bench_xl.json"))
print({k:j[k] for k in ("value","ms_per_step","e2e","e2e_nlc","kernels_launched","kernels_ms_per_call","cpu_baseline")})
r=j["roofline"]
print(r["kernel"], r["frac"], r["share_of_step"])
for k,v in r["kernels"].items(): print(k, v["ms_per_step"], v.get("frac"), v.get("frac_executed"))
PY
python bench.py --steps 1 --warmup 1 --no-ttt --no-cpu > gpurun_out/r02e_plain.json 2>gpurun_out/r02e_plain.err && ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_vnet_tc -c 6 --csv --log-file gpurun_out/r02e_vnet_dram.csv python bench.py --steps 1 --warmup 1 --no-ttt --no-cpu > gpurun_out/r02e_ncu.log 2>&1
python -c "
import json; j=json.load(open('gpurun_out/r02e_plain.json')); print(j['ms_per_step'], j['kernels_ms_per_call'])"
grep -v "^==" gpurun_out/r02e_vnet_dram.csv | cut -d, -f5,13- | head -40
