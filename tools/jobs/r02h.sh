#!/bin/bash
echo "== graph test detail"
timeout 300 python -m pytest tests/test_gpu_api.py -x -q -k "graph" 2>&1 | grep -E "^E|passed|failed" | cut -c1-1800 | head -20
echo "== ncu 3-role kernel"
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --no-ttt --log2n 17 > gpurun_out/r02h_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_xnode3 -c 2 -o gpurun_out/r02h_xnode3 python bench.py --steps 1 --warmup 1 --no-cpu --no-ttt --log2n 17 > gpurun_out/r02h_ncu.log 2>&1; tail -2 gpurun_out/r02h_ncu.log
