#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --config xl --steps 2 --warmup 3 --no-ttt --no-cpu --graph 1 > gpurun_out/r02bg_xl_graph.json 2> gpurun_out/r02bg_xl_graph.err; echo xl rc=$?
python - <<'PY'
import json
t=open("gpurun_out/r02bg_xl_graph.json").read(); j=json.loads(t[t.index('{"metric'):]); print("xl graph", j["ms_per_step"], "%.4g"%j["value"], j["e2e"]["ms_per_step"])
PY
timeout 600 python bench.py --impl reference > gpurun_out/r02bg_ref.json 2> gpurun_out/r02bg_ref.err; echo ref rc=$?
timeout 900 python bench.py > gpurun_out/r02bg_bench.json 2> gpurun_out/r02bg_bench.err; echo bench rc=$?
python - <<'PY'
import json
t=open("gpurun_out/r02bg_bench.json").read(); j=json.loads(t[t.index('{"metric'):])
print(round(j["ms_per_step"],2), "%.4g"%j["value"], "e2e", j["e2e"]["ms_per_step"], "frac", j["roofline"]["frac"], j["clocks"], j["gpu_launches"])
print({k:v for k,v in j["time_to_target"].items() if k in ("sub_iters","seconds","steady_ms_per_outer_iter")})
print({k:(round(v["ms"],2), "%.3g"%v["value"]) for k,v in j["phases"].items() if isinstance(v, dict)})
PY
