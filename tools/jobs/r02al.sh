#!/bin/bash
timeout 300 python bench.py --log2n 17 --steps 10 --warmup 5 --no-cpu --no-ttt --graph 1 > gpurun_out/r02al_g1.json 2> gpurun_out/r02al_g1.err; echo rc=$?
python - 1 <<'PY'
import json, sys
t=open("gpurun_out/r02al_g%s.json" % sys.argv[1]).read(); j=json.loads(t[t.index('{"metric'):]); print("graph", sys.argv[1], round(j["ms_per_step"],3), "%.4g"%j["value"], "e2e", round(j["e2e"]["ms_per_step"],3), "launches", j["gpu_launches"], {k:round(v,3) for k,v in j["kernels_ms_per_step"].items()})
PY
tail -2 gpurun_out/r02al_g1.err | cut -c1-300
CMD="python bench.py --log2n 17 --steps 2 --warmup 3 --no-cpu --no-ttt --nlc-max-gb 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r02al_launches_log2n17.csv $CMD > gpurun_out/r02al_ncu.log 2>&1; echo list rc=$?
