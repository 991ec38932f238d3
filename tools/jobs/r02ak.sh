#!/bin/bash
# exactly what the driver runs at round end: smoke, the default bench line, the reference arm
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4
( time timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02ak_bench.json 2> gpurun_out/r02ak_bench.err ) 2>&1 | tail -3; echo rc=$?
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02ak_ref.json 2> gpurun_out/r02ak_ref.err ) 2>&1 | tail -3; echo rc=$?
python - <<'PY'
import json
t=open("gpurun_out/r02ak_bench.json").read(); j=json.loads(t[t.index('{"metric'):])
print(round(j["ms_per_step"],2), "%.4g"%j["value"], "e2e", j["e2e"]["ms_per_step"], "ttt", j.get("time_to_target",{}).get("sub_iters"), j.get("time_to_target",{}).get("seconds"), "cpu", j.get("cpu_baseline"))
print("roofline", {k: j["roofline"][k] for k in ("bound","kernel","achieved","peak","frac","traffic")})
t=open("gpurun_out/r02ak_ref.json").read(); r=json.loads(t[t.index('{'):]); print("ref", {k: r.get(k) for k in ("impl","value","unit","ms_per_step","cpu_baseline","e2e")})
print("e2e ratio", j["e2e"]["value"]/r["value"])
PY
