#!/bin/bash
# final state of the third session: plain run first, then the ncu launch list of the same command
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-ttt --nlc-max-gb 0"
timeout 300 $CMD > gpurun_out/r02az_plain.json 2> gpurun_out/r02az_plain.err; echo plain rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/r02az_launches.csv $CMD > gpurun_out/r02az_ncu_list.log 2>&1; echo list rc=$?
wc -l gpurun_out/r02az_launches.csv
