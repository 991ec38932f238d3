"""how long fresh device allocations take on this box: every CUDA-graph capture of a new solver allocates its private
pool with cudaMalloc (tools/ttt_probe.py: 12 per solver).  Times raw cudaMalloc / cudaFree through the runtime."""
import ctypes as C
import json
import time

import torch

torch.cuda.init()
x = torch.zeros(1, device="cuda:0")
rt = C.CDLL("libcudart.so.12")
out = {}
for mb in (2, 20, 64):
    ts, fs = [], []
    for _ in range(12):
        p = C.c_void_p()
        t0 = time.perf_counter()
        rc = rt.cudaMalloc(C.byref(p), C.c_size_t(mb << 20))
        t1 = time.perf_counter()
        assert rc == 0
        ts.append(round(1e3 * (t1 - t0), 3))
        t0 = time.perf_counter()
        rt.cudaFree(p)
        fs.append(round(1e3 * (time.perf_counter() - t0), 3))
    out["%d MB" % mb] = {"cudaMalloc_ms": ts, "cudaFree_ms": fs}
print(json.dumps(out))
