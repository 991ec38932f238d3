"""times every C-ABI entry point on synthetic cube data (CUDA events) and the FP32-FMA probe.
usage: python tools/perf_probe.py [log2N] [d]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xnode_wan_b200 as xw  # noqa: E402

L_ = xw._lib


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), ts


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 17
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    lib = L_.XwLib(os.environ.get("XW_LIB", L_.LIB_PATH))
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    N = Nb = 1 << log2n
    L, H, hh, nu, Hv, nv = 20, 20, 10, 8, 50, 9
    dims = L_.Dims(d, H, hh, nu, Hv, nv, 1)
    dom = L_.Domain(0, -1.0, 1.0, 0.0)
    coef = L_.Coef(0.0, -1.0, None, None)
    Pu, Pv = lib.theta_sizes(dims)
    thu = (torch.rand(Pu, device=dev) - 0.5) * 0.6
    thv = (torch.rand(Pv, device=dev) - 0.5) * 0.4
    times = torch.sort(torch.rand(L, device=dev))[0]
    times[0], times[-1] = 0.0, 1.0
    x = torch.rand(N, d, device=dev) * 2 - 1
    xv = torch.rand(N, d, device=dev) * 2 - 1
    xb = torch.rand(Nb, d, device=dev) * 2 - 1
    h = torch.sin(x[:, 0]).contiguous()
    gh = torch.zeros(N, d, device=dev)
    gh[:, 0] = torch.cos(x[:, 0])
    f = torch.randn(N, L, device=dev)
    g = torch.randn(Nb, L, device=dev)
    sb = torch.sin(xb[:, 0]).contiguous()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    pts = L_.Points(times.data_ptr(), 0, 1, xv.data_ptr(), d, 0)     # collapsed layout
    wsb = lib.workspace_bytes(dims, N, L)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    sums = torch.zeros(8, dtype=torch.float64, device=dev)
    cot_u = torch.empty(N * L, device=dev)
    cot_v = torch.empty(N * L, device=dev)
    u_out = torch.empty(N * L, device=dev)
    v_out = torch.empty(N * L, device=dev)
    gu = torch.zeros(Pu, device=dev)
    yh = torch.empty(int(lib.cdll.xw_yhist_floats(C.byref(dims), N, L)), device=dev)
    gv = torch.zeros(Pv, device=dev)
    ku = torch.tensor([0.3, 0.1, 1.0], dtype=torch.float64, device=dev)
    res = {"N": N, "d": d, "L": L}

    def fma(variant):
        fl = C.c_double(0)
        def run():
            lib.call("xw_fma_probe", variant, 4096, C.byref(fl), st)
        t, _ = ev_time(run, reps=5, warm=2)
        return fl.value / (t * 1e-3) / 1e12
    res["fma_tflops_ffma"] = fma(0)
    res["fma_tflops_ffma2"] = fma(1)
    res["mma_sync_tf32_tflops"] = fma(2)

    def xeval():
        lib.call("xw_xnode_eval", C.byref(dims), p(thu), p(x), d, p(times), L, p(h), N, p(u_out), st)
    def veval():
        lib.call("xw_vnet_eval", C.byref(dims), p(thv), C.byref(pts), N, L, p(v_out), st)
    def ifwd():
        lib.call("xw_interior_forward", C.byref(dims), C.byref(dom), C.byref(coef), p(thu), p(thv), p(x), d, p(times), L,
                 C.byref(pts), p(h), p(gh), p(f), N, p(sums), p(cot_u), p(cot_v), p(u_out), p(ws), wsb, st, None, None, 0, p(yh))
    def bdry():
        lib.call("xw_boundary_u", C.byref(dims), p(thu), p(xb), d, p(times), L, p(sb), p(g), Nb, 1e-3, p(sums), p(gu), 0,
                 p(ws), wsb, st)
    def bwdu():
        lib.call("xw_interior_backward_u", C.byref(dims), p(thu), p(x), d, p(times), L, p(h), p(cot_u), N, p(ku), p(gu), 1,
                 p(ws), wsb, st, None, p(yh))
    def bwdv():
        lib.call("xw_interior_backward_v", C.byref(dims), C.byref(dom), p(thv), C.byref(pts), p(cot_v), N, L, p(ku), p(gv),
                 0, p(ws), wsb, st)
    F = (H + d + 1) * hh + (nu - 1) * hh * hh + hh * H
    U = 2 * (L - 1) / L * F + H + (H + 2 * H * H) / L
    Vm = (d + 1) * Hv + nv * Hv * Hv + Hv
    alg = {"xnode_eval": 2 * U, "vnet_eval": 2 * Vm, "interior_forward": 2 * (2 * U + 2 * Vm), "boundary_u": 2 * 3 * U,
           "interior_backward_u": 2 * 2 * U, "interior_backward_v": 2 * 2 * Vm}
    for name, fn in (("xnode_eval", xeval), ("vnet_eval", veval), ("interior_forward", ifwd), ("boundary_u", bdry),
                     ("interior_backward_u", bwdu), ("interior_backward_v", bwdv)):
        t, ts = ev_time(fn)
        res[name + "_ms"] = t
        res[name + "_Mpts_per_s"] = N * L / t / 1e3
        res[name + "_alg_tflops"] = alg[name] * N * L / (t * 1e-3) / 1e12
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
