#!/bin/bash
timeout 120 python tools/small_launches.py | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02au_launches_shipped.csv python tools/small_launches.py > gpurun_out/r02au_ncu.log 2>&1; echo rc=$?
wc -l gpurun_out/r02au_launches_shipped.csv
