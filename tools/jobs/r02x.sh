#!/bin/bash
timeout 300 python tools/tc_prof.py 20 20 2>&1 | tee gpurun_out/r02x_tc_prof.json | tail -60
