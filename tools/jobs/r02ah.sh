#!/bin/bash
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02ah_bench_n$N.json 2> gpurun_out/r02ah_bench_n$N.err; echo rc=$?
grep -v "^\*\|OMP_NUM" gpurun_out/r02ah_bench_n$N.err | tail -30 | cut -c1-300
python - $N <<'PY'
import json, sys
try:
    j=json.load(open("gpurun_out/r02ah_bench_n%s.json" % sys.argv[1])); print("n", j["n_gpus"], round(j["ms_per_step"],2), "%.4g"%j["value"], "e2e", j["e2e"]["ms_per_step"], "strong", j["strong"], "shard", j["shard_check"])
except Exception as e: print("no json:", e)
PY
