"""tcgen05 probe check: D = A B^T on the tensor cores vs fp64 numpy (plain TF32 and 3xTF32)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xnode_wan_b200 as xw
from xnode_wan_b200 import _lib

lib = _lib.get()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
for K, N in ((8, 16), (56, 64), (64, 64), (56, 48), (56, 56), (24, 56)):
    A = torch.randn(128, K, generator=g); B = torch.randn(N, K, generator=g)
    ref = A.double().numpy() @ B.double().numpy().T
    for terms in (1, 3, 11, 13):
        Ad, Bd = A.to(dev), B.to(dev)
        D = torch.full((128, N), float("nan"), device=dev); err = torch.zeros(1, dtype=torch.int32, device=dev)
        lib.call("xw_umma_probe", Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), K, N, terms, err.data_ptr(),
                 torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        d = D.cpu().double().numpy()
        e = np.abs(d - ref).max() / np.abs(ref).max()
        print("K=%d N=%d terms=%d  err_flag=%d  max_rel_err=%.3e  nan=%d" % (K, N, terms, int(err.item()), e, int(np.isnan(d).sum())), flush=True)

# transposed product (weight-gradient shape): out[o][i] = sum_r P[r][o] Q[r][i], MN-major reads of the same images
for MO, NI in ((56, 56), (52, 24), (8, 8), (128, 64)):
    P = torch.randn(128, MO, generator=g); Q = torch.randn(128, NI, generator=g)
    ref = P.double().numpy().T @ Q.double().numpy()
    for terms in (1, 3):
        D = torch.full((128, NI), float("nan"), device=dev); err = torch.zeros(1, dtype=torch.int32, device=dev)
        Pd, Qd = P.to(dev), Q.to(dev)
        lib.call("xw_umma_probe", Pd.data_ptr(), Qd.data_ptr(), D.data_ptr(), MO, NI, -terms, err.data_ptr(),
                 torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        d = D.cpu().double().numpy()
        e = np.abs(d[:MO] - ref).max() / np.abs(ref).max()
        z = np.abs(d[MO:]).max() if MO < 128 else 0.0
        print("T: MO=%d NI=%d terms=%d  err_flag=%d  max_rel_err=%.3e  pad_rows_max=%.1e nan=%d" % (MO, NI, terms, int(err.item()), e, z, int(np.isnan(d).sum())), flush=True)
