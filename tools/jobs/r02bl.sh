#!/bin/bash
ITERS=2 timeout 240 compute-sanitizer --tool initcheck --print-limit 30 python tools/small_launches.py > gpurun_out/r02bl_initcheck.log 2>&1; echo rc=$?
grep -c "Uninitialized" gpurun_out/r02bl_initcheck.log
grep -A6 "Uninitialized" gpurun_out/r02bl_initcheck.log | grep -E "Uninitialized|at .*\(|in " | head -30
tail -4 gpurun_out/r02bl_initcheck.log
