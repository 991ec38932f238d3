#!/bin/bash
# final-state check of round 2: GPU tests, smoke, default bench (both arms), each timed
S=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4; echo "pytest $(( $(date +%s)-S )) s"
S=$(date +%s)
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2; echo "smoke $(( $(date +%s)-S )) s"
S=$(date +%s)
timeout 600 python bench.py --impl reference > gpurun_out/r02am_ref.json 2> gpurun_out/r02am_ref.err; echo "ref rc=$? $(( $(date +%s)-S )) s"
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/r02am_bench.json 2> gpurun_out/r02am_bench.err; echo "bench rc=$? $(( $(date +%s)-S )) s"
python - <<'PY'
import json
t=open("gpurun_out/r02am_bench.json").read(); j=json.loads(t[t.index('{"metric'):])
print(round(j["ms_per_step"],2), "%.4g"%j["value"], "e2e", j["e2e"]["ms_per_step"], "frac", j["roofline"]["frac"], j["clocks"], j.get("cpu_baseline"), j.get("variants"))
print({k:v for k,v in j.get("time_to_target",{}).items() if k!="runs"})
t=open("gpurun_out/r02am_ref.json").read(); r=json.loads(t[t.index('{'):]); print("ref", r.get("value"), r.get("cpu_baseline"))
PY
tail -3 gpurun_out/r02am_bench.err | cut -c1-300
