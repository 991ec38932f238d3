#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 900 python bench.py --config xl --steps 2 --warmup 3 --no-ttt --no-cpu --graph 1 > gpurun_out/r02bh_xl_graph.json 2> gpurun_out/r02bh_xl_graph.err; echo xl rc=$?
for L2N in 17 20; do
timeout 300 python bench.py --log2n $L2N --steps 10 --warmup 5 --no-cpu --no-ttt --nlc-max-gb 0 --graph 1 > gpurun_out/r02bh_g$L2N.json 2> gpurun_out/r02bh_g$L2N.err; echo rc=$?
done
python - <<'PY'
import json
for f in ("xl_graph","g17","g20"):
    try:
        t=open("gpurun_out/r02bh_%s.json"%f).read(); j=json.loads(t[t.index('{"metric'):]); print(f, round(j["ms_per_step"],3), "%.4g"%j["value"], "e2e", round(j["e2e"]["ms_per_step"],3))
    except Exception as e: print(f, "failed", e)
PY
tail -2 gpurun_out/r02bh_xl_graph.err | cut -c1-300
