#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02be_bench_n8.json 2> gpurun_out/r02be_bench_n8.err; echo rc=$?
python - <<'PY'
import json
t=open("gpurun_out/r02be_bench_n8.json").read(); j=json.loads(t[t.index('{"metric'):]); print(j["ms_per_step"], "%.4g"%j["value"], j["e2e"]["ms_per_step"], j["strong"], j["shard_check"])
PY
tail -3 gpurun_out/r02be_bench_n8.err | cut -c1-300
