"""drives the C ABI (include/xnode_wan_b200.h) end to end on one batch with either numpy arrays
(CPU emulation build, kernel-logic tests) or torch CUDA tensors (the real library, -m gpu tests)."""
import ctypes as C

import numpy as np

import xnode_wan_b200 as xw

L_ = xw._lib


class NumpyBackend:
    def arr(self, a, dtype=np.float32):
        return np.ascontiguousarray(np.asarray(a, dtype=dtype))

    def zeros(self, n, dtype=np.float32):
        return np.zeros(int(n), dtype=dtype)

    def ptr(self, a):
        return C.c_void_p(a.ctypes.data) if a is not None else None

    def ptr_off(self, a, off_elems):
        return C.c_void_p(a.ctypes.data + off_elems * a.itemsize)

    def host(self, a):
        return np.asarray(a)

    stream = None

    def sync(self):
        pass


class TorchBackend:
    def __init__(self):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda:0")
        self.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def arr(self, a, dtype=np.float32):
        return self.torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=dtype))).to(self.dev)

    def zeros(self, n, dtype=np.float32):
        td = {np.float32: self.torch.float32, np.float64: self.torch.float64, np.uint8: self.torch.uint8}[dtype]
        return self.torch.zeros(int(n), dtype=td, device=self.dev)

    def ptr(self, a):
        return C.c_void_p(a.data_ptr()) if a is not None else None

    def ptr_off(self, a, off_elems):
        return C.c_void_p(a.data_ptr() + off_elems * a.element_size())

    def host(self, a):
        return a.detach().cpu().numpy()

    def sync(self):
        self.torch.cuda.synchronize()


def flat_theta(lst):
    return np.concatenate([np.asarray(p, np.float64).reshape(-1) for p in lst]).astype(np.float32)


def split_like(flat, lst):
    out, o = [], 0
    for p in lst:
        n = int(np.prod(p.shape))
        out.append(np.asarray(flat[o:o + n]).reshape(p.shape))
        o += n
    assert o == len(flat)
    return out


def make_dims(case):
    p = case["params"]
    return L_.Dims(p["dim"], p["u_hidden_dim"], p["u_hidden_hidden_dim"], p["u_layers"], p["v_hidden_dim"],
                   p["v_layers"], L_.SOLVERS[p["solver"]])


def make_domain(dom):
    kind = L_.DOMAINS[dom[0]]
    ps = [float(x) for x in dom[1:]] + [0.0, 0.0, 0.0]
    return L_.Domain(kind, ps[0], ps[1], ps[2])


def run_case(lib, be, case, coef_a=None, coef_b=None, use_yhist=True):
    """full u-phase and v-phase evaluation of one golden case through the C ABI.
    returns dict(I,S,init,bdry,loss_u,loss_v,grads_u,grads_v,sums,u)"""
    z, p = case["z"], case["params"]
    if coef_a is None and "coef_a" in z.files:
        coef_a = z["coef_a"]
    dims = make_dims(case)
    dom = make_domain(case["meta"]["domain"])
    X, XV, BX = z["X"], z["XV"], z["BX"]
    N, L, Cc = X.shape
    Nb, Lb, _ = BX.shape
    d = Cc - 1
    alpha, V = float(p["alpha"]), float(case["meta"]["V"])
    Xd, XVd, BXd = be.arr(X), be.arr(XV), be.arr(BX)
    times = be.arr(X[0, :, 0])
    times_b = be.arr(BX[0, :, 0])
    thu = be.arr(flat_theta(case["thu_list"]))
    thv = be.arr(flat_theta(case["thv_list"]))
    Pu, Pv = lib.theta_sizes(dims)
    assert Pu == thu.shape[0] and Pv == thv.shape[0]
    h, gh, f = be.arr(z["h"]), be.arr(z["grad_h"]), be.arr(z["f"])
    s0 = be.arr(z["s0"]) if "s0" in z.files else None
    g, sb = be.arr(z["g"]), be.arr(z["sb"])
    a_dev = be.arr(coef_a) if coef_a is not None else None
    b_dev = be.arr(coef_b) if coef_b is not None else None
    coef = L_.Coef(case["meta"]["c0"], case["meta"]["c1"], be.ptr(a_dev).value if a_dev is not None else None,
                   be.ptr(b_dev).value if b_dev is not None else None)
    pts = L_.Points(be.ptr(XVd).value, L * Cc, Cc, be.ptr_off(XVd, 1).value, L * Cc, Cc)
    wsb = max(lib.workspace_bytes(dims, N, L), lib.workspace_bytes(dims, Nb, Lb))
    ws = be.zeros(wsb, np.uint8)
    sums = be.zeros(L_.NSUMS, np.float64)
    cot_u, cot_v, u_out = be.zeros(N * L), be.zeros(N * L), be.zeros(N * L)
    vcache = be.zeros(lib.cdll.xw_vcache_floats(C.byref(dims), N, L))
    yhist = be.zeros(lib.cdll.xw_yhist_floats(C.byref(dims), N, L)) if use_yhist else None
    lib.call("xw_interior_forward", C.byref(dims), C.byref(dom), C.byref(coef), be.ptr(thu), be.ptr(thv),
             be.ptr_off(Xd, 1), L * Cc, be.ptr(times), L, C.byref(pts), be.ptr(h), be.ptr(gh), be.ptr(f), N,
             be.ptr(sums), be.ptr(cot_u), be.ptr(cot_v), be.ptr(u_out), be.ptr(ws), wsb, be.stream, be.ptr(s0),
             be.ptr(vcache), 1, be.ptr(yhist), len(vcache), len(yhist) if yhist is not None else 0)
    gu = be.zeros(Pu)
    lib.call("xw_boundary_u", C.byref(dims), be.ptr(thu), be.ptr_off(BXd, 1), Lb * Cc, be.ptr(times_b), Lb,
             be.ptr(sb), be.ptr(g), Nb, alpha / (Nb * Lb), be.ptr(sums), be.ptr(gu), 0, be.ptr(ws), wsb, be.stream)
    be.sync()
    # second evaluation from the test-function cache (mode 2) must reproduce the sums and the seeds
    sums2 = be.zeros(L_.NSUMS, np.float64)
    cu2, cv2 = be.zeros(N * L), be.zeros(N * L)
    lib.call("xw_interior_forward", C.byref(dims), C.byref(dom), C.byref(coef), be.ptr(thu), be.ptr(thv),
             be.ptr_off(Xd, 1), L * Cc, be.ptr(times), L, C.byref(pts), be.ptr(h), be.ptr(gh), be.ptr(f), N,
             be.ptr(sums2), be.ptr(cu2), be.ptr(cv2), None, be.ptr(ws), wsb, be.stream, be.ptr(s0), be.ptr(vcache), 2, None,
             len(vcache), 0)
    be.sync()
    s_a, s_b = be.host(sums)[:5].copy(), be.host(sums2)[:5].copy()
    assert np.allclose(s_a, s_b, rtol=1e-6, atol=1e-6 * np.abs(s_a).max()), (s_a, s_b)
    assert np.array_equal(be.host(cot_u), be.host(cu2)) and np.array_equal(be.host(cot_v), be.host(cv2))
    s = be.host(sums).copy()
    I = V / N * s[0] - V / (N * L) * (s[1] - s[2])
    S = V * s[3] / (N * L)
    init, bdry = s[4] / N, s[5] / (Nb * Lb)
    integ = np.log(I * I) - np.log(S)
    out = dict(I=I, S=S, init=init, bdry=bdry, loss_u=integ + alpha * (init + bdry), loss_v=-integ, sums=s,
               u=be.host(u_out).reshape(N, L).copy())
    ku = be.arr(np.array([(2.0 / I) * V / (N * L), 2.0 * alpha / N, 1.0]), np.float64)
    lib.call("xw_interior_backward_u", C.byref(dims), be.ptr(thu), be.ptr_off(Xd, 1), L * Cc, be.ptr(times), L,
             be.ptr(h), be.ptr(cot_u), N, be.ptr(ku), be.ptr(gu), 1, be.ptr(ws), wsb, be.stream, be.ptr(s0), be.ptr(yhist))
    kv = be.arr(np.array([-(2.0 / I) * V / (N * L), 2.0 / s[3], 1.0]), np.float64)
    gv = be.zeros(Pv)
    lib.call("xw_interior_backward_v", C.byref(dims), C.byref(dom), be.ptr(thv), C.byref(pts), be.ptr(cot_v), N, L,
             be.ptr(kv), be.ptr(gv), 0, be.ptr(ws), wsb, be.stream)
    be.sync()
    out["grads_u"] = split_like(be.host(gu), case["thu_list"])
    out["grads_v"] = split_like(be.host(gv), case["thv_list"])
    out["cot_u"] = be.host(cot_u).reshape(N, L).copy()
    out["cot_v"] = be.host(cot_v).reshape(N, L).copy()
    return out
