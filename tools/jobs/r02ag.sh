#!/bin/bash
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vnet_tc_bwd3" -s 1 -c 1 -o gpurun_out/r02ag_bwd3 python bench.py --steps 1 --warmup 3 --no-cpu --no-ttt --nlc-max-gb 0 > gpurun_out/r02ag_ncu.log 2>&1; echo full rc=$?
ls -la gpurun_out/*.ncu-rep
