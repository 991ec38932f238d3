#!/bin/bash
for b in 0 1; do echo "--- bwd split=$b d=100"; XW_TC_SPLIT_BWD=$b timeout 400 python tools/dbg_mid.py 100 307 203 3 4 5 6 7 8 2>&1 | grep "grad rel"; done
for b in 0 1; do echo "--- bwd split=$b d=20"; XW_TC_SPLIT_BWD=$b timeout 400 python tools/dbg_mid.py 20 700 300 3 4 5 6 7 8 2>&1 | grep "grad rel"; done
