#!/bin/bash
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02o_bench_n8.json 2> gpurun_out/r02o_bench_n8.err; echo rc=$?; tail -c 600 gpurun_out/r02o_bench_n8.err
python - <<'PY'
import json
try:
    j=json.loads(open("gpurun_out/r02o_bench_n8.json").read().strip().split("\n")[-1]); print(round(j["ms_per_step"],2), "%.4g"%j["value"], j["strong"], j["shard_check"]); print(j["e2e"])
except Exception as e: print("ERR", e)
PY
