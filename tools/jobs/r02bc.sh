#!/bin/bash
timeout 120 python tools/malloc_probe.py 2>&1 | tail -2
PROBE_COLD=1 PROBE_SYNC=0 timeout 300 python tools/ttt_probe.py 2>&1 | tail -1 | cut -c1-2500
