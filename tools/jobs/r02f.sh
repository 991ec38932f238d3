#!/bin/bash
python tools/small_n_probe.py > gpurun_out/r02f_small_n.json 2> gpurun_out/r02f_small_n.prof; cat gpurun_out/r02f_small_n.json; grep -A40 "cumulative" gpurun_out/r02f_small_n.prof | cut -c1-150 | head -45
