#!/bin/bash
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --no-ttt --log2n 17 > gpurun_out/r02s_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_xnode3 -c 2 -o gpurun_out/r02s_xnode3 python bench.py --steps 1 --warmup 1 --no-cpu --no-ttt --log2n 17 > gpurun_out/r02s_ncu.log 2>&1; tail -2 gpurun_out/r02s_ncu.log
