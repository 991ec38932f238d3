#!/bin/bash
timeout 200 python - <<'PY' 2>/dev/null | tail -3
import sys, json
sys.path.insert(0, ".")
import torch, bench
import xnode_wan_b200 as xw
out = bench.time_to_target(xw, "cuda:0", [0, 0, 0, 1, 0])
print(out["seeds"], out["sub_iters"], out["final_rel_l2"])
PY
