import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xnode_wan_b200 as xw
from xnode_wan_b200 import _lib
lib = _lib.get(); dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
MO, NI = 128, 64
P = torch.randn(128, MO, generator=g); Q = torch.randn(128, NI, generator=g)
ref = P.double().numpy().T @ Q.double().numpy()
D = torch.full((128, NI), float("nan"), device=dev); err = torch.zeros(1, dtype=torch.int32, device=dev)
Pd, Qd = P.to(dev), Q.to(dev)
lib.call("xw_umma_probe", Pd.data_ptr(), Qd.data_ptr(), D.data_ptr(), MO, NI, -1, err.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
d = D.cpu().double().numpy()
print("variant", os.environ.get("XW_UMMA_VARIANT"), "err", int(err.item()), "absmax", np.abs(d).max(), "relerr", np.abs(d - ref).max() / np.abs(ref).max())
print(d[:2, :4]); print(ref[:2, :4])
