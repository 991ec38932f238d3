#!/bin/bash
timeout 200 python tools/mma_bench.py 2>&1 | tee gpurun_out/r02y_mma_bench.txt | tail -50
