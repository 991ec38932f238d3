#!/bin/bash
PROBE_SYNC=1 timeout 300 python tools/ttt_probe.py 2>&1 | tail -1
PROBE_SYNC=0 timeout 300 python tools/ttt_probe.py 2>&1 | tail -1
