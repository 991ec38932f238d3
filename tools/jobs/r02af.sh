#!/bin/bash
# final-state profiles: plain run first, then the ncu launch list of the same command and one --set full capture
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-ttt --nlc-max-gb 0"
timeout 300 $CMD > gpurun_out/r02af_plain.json 2> gpurun_out/r02af_plain.err; echo plain rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/r02af_launches.csv $CMD > gpurun_out/r02af_ncu_list.log 2>&1; echo list rc=$?
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_vnet_tc|k_xnode3_bwd|k_xnode2_fwd" -s 12 -c 9 -o gpurun_out/r02af_full python bench.py --steps 1 --warmup 3 --no-cpu --no-ttt --nlc-max-gb 0 > gpurun_out/r02af_ncu_full.log 2>&1; echo full rc=$?
ls -la gpurun_out/ | tail -8
