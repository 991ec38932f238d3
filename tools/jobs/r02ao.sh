#!/bin/bash
timeout 300 python tools/small_n_probe.py > gpurun_out/r02ao_probe.json 2> gpurun_out/r02ao_probe.err; echo rc=$?
cat gpurun_out/r02ao_probe.json
grep -v "^$" gpurun_out/r02ao_probe.err | head -60 | cut -c1-180
