"""compares the time-row-0 gradient cache and the sums of xw_interior_forward between the kernel variants"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xnode_wan_b200 as xw
L_ = xw._lib
lib = L_.get(); dev = torch.device("cuda:0"); torch.manual_seed(0)
N = 5000; d, L, H, hh, nu, Hv, nv = 20, 20, 20, 10, 8, 50, 9
dims = L_.Dims(d, H, hh, nu, Hv, nv, 1)
pu, pv = lib.theta_sizes(dims)
thu = (torch.randn(pu, device=dev) * 0.2); thv = (torch.randn(pv, device=dev) * 0.2)
x = torch.rand(N, d, device=dev) * 2 - 1; xv = torch.rand(N, d, device=dev) * 2 - 1
times = torch.linspace(0, 1, L, device=dev)
h = torch.randn(N, device=dev); gh = torch.randn(N, d, device=dev); f = torch.randn(N * L, device=dev)
wsb = lib.workspace_bytes(dims, N, L); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
dom = L_.Domain(0, -1.0, 1.0, 0.0); coef = L_.Coef(0.0, 1.0, None, None)
pts = L_.Points(times.data_ptr(), 0, 1, xv.data_ptr(), d, 0)
st = torch.cuda.current_stream().cuda_stream
out = {}
for impl in ("tc", "tile"):
    os.environ["XW_VNET_IMPL"] = impl
    sums = torch.zeros(8, dtype=torch.float64, device=dev)
    cu = torch.empty(N * L, device=dev); cv = torch.empty(N * L, device=dev)
    vc = torch.zeros(int(lib.cdll.xw_vcache_floats(C.byref(dims), N, L)), device=dev)
    lib.call("xw_interior_forward", C.byref(dims), C.byref(dom), C.byref(coef), thu.data_ptr(), thv.data_ptr(), x.data_ptr(), d,
             times.data_ptr(), L, C.byref(pts), h.data_ptr(), gh.data_ptr(), f.data_ptr(), N, sums.data_ptr(), cu.data_ptr(),
             cv.data_ptr(), None, ws.data_ptr(), wsb, st, None, vc.data_ptr(), 1, None)
    torch.cuda.synchronize()
    out[impl] = (sums.cpu().numpy().copy(), vc[4 * N * L:].cpu().numpy().copy(), vc[:4 * N * L].cpu().numpy().copy())
sa, ga, va = out["tc"]; sb, gb, vb = out["tile"]
print("sums tc  ", sa[:4]); print("sums tile", sb[:4])
print("gcache rel-L2 diff %.3e  max abs %.3e (scale %.3e)" % (np.linalg.norm(ga - gb) / np.linalg.norm(gb), np.abs(ga - gb).max(), np.abs(gb).max()))
print("vcache rel-L2 diff %.3e" % (np.linalg.norm(va - vb) / np.linalg.norm(vb)))
