"""where do the rows of an M=64 tcgen05.mma (cta_group::1) land in tensor memory?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xnode_wan_b200 as xw
from xnode_wan_b200 import _lib
lib = _lib.get(); dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
K, N = 16, 16
A = torch.randn(128, K, generator=g); B = torch.randn(N, K, generator=g)
ref = A.double().numpy() @ B.double().numpy().T
D = torch.full((128, N), 777.0, device=dev); err = torch.zeros(1, dtype=torch.int32, device=dev)
# pre-fill tensor memory lanes with a marker by a plain M=128 run of zeros first is not possible; just read
Ad, Bd = A.to(dev), B.to(dev)
lib.call("xw_umma_probe", Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), K, N, 3, err.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
d = D.cpu().double().numpy()
for lane in range(128):
    errs = np.abs(ref - d[lane][None, :]).max(axis=1)
    r = int(errs.argmin())
    print("lane %3d <- row %3d (err %.1e)" % (lane, r, errs[r]) if errs[r] < 1e-3 else "lane %3d <- (no row)" % lane)
