#!/bin/bash
echo base;  PROBE_COLD=1 PROBE_SYNC=0 timeout 300 python tools/ttt_probe.py 2>&1 | tail -1 | python -c "import sys,json; [print(c['seed'], c['wall_s'], c['events_over_10ms']) for c in json.loads(sys.stdin.read())['cold']]"
echo keep;  PROBE_KEEP=1 PROBE_COLD=1 PROBE_SYNC=0 timeout 300 python tools/ttt_probe.py 2>&1 | tail -1 | python -c "import sys,json; [print(c['seed'], c['wall_s'], c['events_over_10ms']) for c in json.loads(sys.stdin.read())['cold']]"
echo empty; PROBE_EMPTY=1 PROBE_COLD=1 PROBE_SYNC=0 timeout 300 python tools/ttt_probe.py 2>&1 | tail -1 | python -c "import sys,json; [print(c['seed'], c['wall_s'], c['events_over_10ms']) for c in json.loads(sys.stdin.read())['cold']]"
