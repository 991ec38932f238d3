#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02ax_bench_n2.json 2> gpurun_out/r02ax_bench_n2.err; echo rc=$?
python - <<'PY'
import json
t=open("gpurun_out/r02ax_bench_n2.json").read(); j=json.loads(t[t.index('{"metric'):]); print(j["ms_per_step"], "%.4g"%j["value"], j["e2e"]["ms_per_step"], j["strong"], j["shard_check"])
PY
tail -3 gpurun_out/r02ax_bench_n2.err | cut -c1-300
