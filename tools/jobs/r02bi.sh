#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 900 python bench.py --config xl --steps 2 --warmup 3 --no-ttt --no-cpu --graph 1 > gpurun_out/r02bi_xl_graph.json 2> gpurun_out/r02bi_xl_graph.err; echo xl rc=$?
timeout 300 python bench.py --log2n 20 --steps 10 --warmup 5 --no-cpu --no-ttt --nlc-max-gb 0 --graph 1 > gpurun_out/r02bi_g20.json 2> gpurun_out/r02bi_g20.err; echo rc=$?
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --nlc-max-gb 0 > gpurun_out/r02bi_bench.json 2> gpurun_out/r02bi_bench.err; echo rc=$?
python - <<'PY'
import json
for f in ("xl_graph","g20","bench"):
    try:
        t=open("gpurun_out/r02bi_%s.json"%f).read(); j=json.loads(t[t.index('{"metric'):]); print(f, round(j["ms_per_step"],3), "%.4g"%j["value"], "e2e", round(j["e2e"]["ms_per_step"],3), {k:v for k,v in j.get("time_to_target",{}).items() if k in ("sub_iters","seconds","steady_ms_per_outer_iter")})
    except Exception as e: print(f, "failed", e)
PY
tail -2 gpurun_out/r02bi_xl_graph.err | cut -c1-300
