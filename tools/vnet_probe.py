"""times only the test-function entry points (interior forward + v backward) -- short, for ncu.
usage: python tools/vnet_probe.py [log2N] [reps]"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xnode_wan_b200 as xw
L_ = xw._lib
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 15
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib = L_.get(); dev = torch.device("cuda:0"); torch.manual_seed(0)
N = 1 << log2n; d, L, H, hh, nu, Hv, nv = 20, 20, 20, 10, 8, 50, 9
dims = L_.Dims(d, H, hh, nu, Hv, nv, 1)
pu, pv = lib.theta_sizes(dims)
thu = (torch.randn(pu, device=dev) * 0.2); thv = (torch.randn(pv, device=dev) * 0.2)
x = torch.rand(N, d, device=dev) * 2 - 1; xv = torch.rand(N, d, device=dev) * 2 - 1
times = torch.linspace(0, 1, L, device=dev)
h = torch.randn(N, device=dev); gh = torch.randn(N, d, device=dev); f = torch.randn(N * L, device=dev)
sums = torch.zeros(8, dtype=torch.float64, device=dev)
cu = torch.empty(N * L, device=dev); cv = torch.empty(N * L, device=dev)
wsb = lib.workspace_bytes(dims, N, L); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
dom = L_.Domain(0, -1.0, 1.0, 0.0); coef = L_.Coef(0.0, 1.0, None, None)
pts = L_.Points(times.data_ptr(), 0, 1, xv.data_ptr(), d, 0)
st = torch.cuda.current_stream().cuda_stream
coefs = torch.tensor([1e-3, 1e-3, 1.0], dtype=torch.float64, device=dev)
gv = torch.zeros(pv, device=dev)
def fwd():
    lib.call("xw_interior_forward", C.byref(dims), C.byref(dom), C.byref(coef), thu.data_ptr(), thv.data_ptr(), x.data_ptr(), d,
             times.data_ptr(), L, C.byref(pts), h.data_ptr(), gh.data_ptr(), f.data_ptr(), N, sums.data_ptr(), cu.data_ptr(),
             cv.data_ptr(), None, ws.data_ptr(), wsb, st, None, None, 0, None)
def bwd():
    lib.call("xw_interior_backward_v", C.byref(dims), C.byref(dom), thv.data_ptr(), C.byref(pts), cv.data_ptr(), N, L,
             coefs.data_ptr(), gv.data_ptr(), 0, ws.data_ptr(), wsb, st)
for name, fn in (("interior_forward", fwd), ("interior_backward_v", bwd)):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print("%s N=2^%d: %.3f ms  (%.1f Mpts/s)" % (name, log2n, best, N * L / best / 1e3), flush=True)
print("grad_v norm %.6e  sums %s" % (float(gv.norm()), sums.cpu().numpy()[:4]))
