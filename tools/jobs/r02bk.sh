#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
PROBE_SYNC=0 timeout 300 python tools/ttt_probe.py 2>&1 | tail -1 | cut -c1-120
