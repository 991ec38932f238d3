"""GPU (-m gpu): the tcgen05 (5th-generation tensor core) path of the test-function net.
* the 3xTF32 building block (C-ABI `xw_umma_probe`) against fp64 numpy: plain TF32 is NOT accurate enough
  for the 1e-4 / 1e-3 tolerances of the path, the error-compensated form is;
* the kernel variants (tensor-core pipeline, serial tensor-core backward, FP32 tile engine, one thread per
  point) all reproduce the golden vectors of the unmodified reference and agree with each other."""
import os

import numpy as np
import pytest
import torch

import xnode_wan_b200 as xw
from tests import _golden as G
from tests import _lowlevel as LL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return xw._lib.get()


def _probe(lib, A, B, K, N, terms):
    dev = torch.device("cuda:0")
    D = torch.full((128, N), float("nan"), device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    Ad, Bd = A.to(dev), B.to(dev)            # (kept alive: a temporary would be recycled before the launch)
    lib.call("xw_umma_probe", Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), K, N, terms, err.data_ptr(),
             torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "a bounded mbarrier wait expired"
    return D.cpu().double().numpy()


@pytest.mark.parametrize("K,N", [(8, 16), (24, 56), (56, 56), (64, 64)])
def test_umma_probe_3xtf32_accuracy(lib, K, N):
    g = torch.Generator().manual_seed(K * 100 + N)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ref = A.double().numpy() @ B.double().numpy().T
    scale = np.abs(ref).max()
    for ts in (0, 10):          # A from shared memory / A from tensor memory
        e1 = np.abs(_probe(lib, A, B, K, N, 1 + ts) - ref).max() / scale
        e3 = np.abs(_probe(lib, A, B, K, N, 3 + ts) - ref).max() / scale
        assert 1e-5 < e1 < 5e-3, e1          # one TF32 MMA: 10-bit mantissas
        assert e3 < 5e-6, e3                 # 3xTF32: fp32 level


def test_umma_probe_rejects_bad_shapes(lib):
    z = torch.zeros(128 * 64, device="cuda:0")
    err = torch.zeros(1, dtype=torch.int32, device="cuda:0")
    with pytest.raises(xw._lib.XwError):
        lib.call("xw_umma_probe", z.data_ptr(), z.data_ptr(), z.data_ptr(), 12, 16, 3, err.data_ptr(), None)
    with pytest.raises(xw._lib.XwError):
        lib.call("xw_umma_probe", z.data_ptr(), z.data_ptr(), z.data_ptr(), 8, 16, 2, err.data_ptr(), None)


VARIANTS = [("tc", None), ("tile", None)]


@pytest.mark.parametrize("name", ["cube_d20_ex41", "cube_d3_rk4", "cone_d5_g2", "hourglass_d5_g4_reentry"])
def test_kernel_variants_agree_and_match_golden(lib, name):
    c = G.load(name)
    z = c["z"]
    out = {}
    old = {k: os.environ.get(k) for k in ("XW_VNET_IMPL", "XW_VNET_BWD")}
    try:
        for impl, bwd in VARIANTS:
            os.environ["XW_VNET_IMPL"] = impl
            if bwd:
                os.environ["XW_VNET_BWD"] = bwd
            else:
                os.environ.pop("XW_VNET_BWD", None)
            out[(impl, bwd)] = LL.run_case(lib, LL.TorchBackend(), c)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for key, r in out.items():
        for k in ("I", "S"):
            assert abs(r[k] - float(z[k])) <= 1e-4 * abs(float(z[k])) + 1e-9, (key, k)
        for i, (a, b) in enumerate(zip(r["grads_v"], c["gv"])):
            assert G.rel(a, b) < 1e-3, (key, "grad_v", i, G.rel(a, b))
    base = out[("tile", None)]
    for key, r in out.items():
        # (I is a difference of Monte-Carlo sums: on the small sphere groups it cancels to ~1e-2 of its terms)
        assert abs(r["I"] - base["I"]) <= 2e-4 * abs(base["I"]) + 1e-9, key
        assert abs(r["loss_v"] - base["loss_v"]) <= 2e-4 * abs(base["loss_v"]) + 1e-6, key
        for a, b in zip(r["grads_v"], base["grads_v"]):
            assert G.rel(a, b) < 3e-4, (key, G.rel(a, b))


def test_many_iterations_ragged_sizes_stay_finite_and_deterministic():
    """soak of the warp-specialised pipeline: 150 training iterations (2 u-steps + 1 v-step each, CUDA-graph
    replay) at a size that is not a multiple of any tile (tail tiles, partially filled last CTA): every loss stays
    finite, no bounded wait expires (a trap would surface as a CUDA error), and two runs from the same seed agree"""
    def run():
        torch.manual_seed(7)
        N = (1 << 14) + 37
        p = xw.problems.cube_params(dim=20, N_r=N, N_b=N // 2 + 5, alpha=10.0, shape_param=[-1.0, 1.0])
        prob = xw.problems.ex4_1()
        s = xw.NODE_WAN_solver(p, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g, "cuda:0",
                               "./", func_u_sol=prob.func_u_sol, p=2, log_json=False)
        dom = s.new_domain(sample_device=torch.device("cuda:0"), collapsed=True)
        pts = xw.Comb_loader(N, N // 2 + 5, dom, torch.device("cuda:0"))
        pts[0]
        out = []
        for it in range(150):
            lu, lv = s.train_iteration(dom, pts)
            if it % 30 == 29:
                out.append((float(lu), float(lv)))
        torch.cuda.synchronize()
        return out
    a = run()
    b = run()
    assert all(np.isfinite(x) and np.isfinite(y) for x, y in a), a
    for (x0, y0), (x1, y1) in zip(a, b):      # fp32 atomics in the reductions: not bit-exact, but close
        assert abs(x0 - x1) <= 1e-3 * abs(x0) + 1e-6 and abs(y0 - y1) <= 1e-3 * abs(y0) + 1e-6, (a, b)


@pytest.mark.parametrize("d", [20, 50, 100])
def test_forward_values_agree_point_by_point_with_the_fp32_kernels(lib, d):
    """Every ROW of the tensor-core forward -- v and dv/dt of every point (test-function cache) and both cotangent seeds
    -- against the FP32 tile engine on the same 2^15+13 paths, several launches.  Sums and gradients (the golden tests)
    average over the points; this pins each of them: a tile whose MMAs raced (three issuing warps accumulate into one
    tensor-memory accumulator, xw_vnet_tc.cuh issue_3xtf32_split) or a k-step that got lost would show up here as a row
    that is off by far more than fp32 rounding.  d = 50 is the widest direct input (kin = 56), d = 100 runs on the virtual
    net of input width Hv."""
    import ctypes as C
    L_ = xw._lib
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(100 + d)
    n, L = (1 << 15) + 13, 20
    dims = L_.Dims(d, 20, 10, 8, 50, 9, 1)
    dom = L_.Domain(0, -1.0, 1.0, 0.0)
    Pu, Pv = lib.theta_sizes(dims)
    thu = (torch.rand(Pu, device=dev, generator=g) - 0.5) * 0.5
    thv = (torch.rand(Pv, device=dev, generator=g) - 0.5) * (0.5 if d <= 50 else 0.3)
    x = torch.rand(n, d, device=dev, generator=g) * 2 - 1
    xv = torch.rand(n, d, device=dev, generator=g) * 2 - 1
    times = torch.sort(torch.rand(L, device=dev, generator=g))[0].contiguous()
    times[0] = 0.0
    h = torch.randn(n, device=dev, generator=g)
    gh = torch.zeros(n, d, device=dev)
    f = torch.randn(n * L, device=dev, generator=g)
    coef = L_.Coef(0.0, -1.0, None, None)
    pts = L_.Points(times.data_ptr(), 0, 1, xv.data_ptr(), d, 0)
    wsb = lib.workspace_bytes(dims, n, L)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    nvc = lib.cdll.xw_vcache_floats(C.byref(dims), n, L)
    st = torch.cuda.current_stream().cuda_stream

    def run(impl, split="0"):
        old, old_s = os.environ.get("XW_VNET_IMPL"), os.environ.get("XW_TC_SPLIT")
        os.environ["XW_VNET_IMPL"] = impl
        os.environ["XW_TC_SPLIT"] = split
        try:
            sums = torch.zeros(L_.NSUMS, dtype=torch.float64, device=dev)
            cu, cv = torch.zeros(n * L, device=dev), torch.zeros(n * L, device=dev)
            vc = torch.zeros(nvc, device=dev)
            lib.call("xw_interior_forward", C.byref(dims), C.byref(dom), C.byref(coef), thu.data_ptr(), thv.data_ptr(),
                     x.data_ptr(), d, times.data_ptr(), L, C.byref(pts), h.data_ptr(), gh.data_ptr(), f.data_ptr(), n,
                     sums.data_ptr(), cu.data_ptr(), cv.data_ptr(), None, ws.data_ptr(), wsb, st, None, vc.data_ptr(), 1, None,
                     nvc, 0)
            torch.cuda.synchronize()
            return vc[:4 * n * L].view(n * L, 4).clone(), cu, cv, sums.cpu().numpy()
        finally:
            for k, v in (("XW_VNET_IMPL", old), ("XW_TC_SPLIT", old_s)):
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    ref_vc, ref_cu, ref_cv, ref_s = run("tile")
    vs, ts = float(ref_vc[:, 0].abs().max()), float(ref_vc[:, 1].abs().max())
    assert vs > 1e-3 and ts > 1e-5
    for rep in range(6):           # the default single-issuer kernel twice, the three-issuer variant (XW_TC_SPLIT=1) four times
        vc, cu, cv, s = run("tc", "0" if rep < 2 else "1")
        assert (lib.cdll.xw_last_vnet_impl() & 15) == 3
        # v is continuous in the pre-activations: EVERY point agrees to fp32 rounding (an MMA covers all 128 rows of a
        # tile, value and tangent rows alike: a lost or raced update would move v)
        assert float((vc[:, 0] - ref_vc[:, 0]).abs().max()) <= 2e-5 * vs, (d, rep)
        assert float((cu - ref_cu).abs().max()) <= 5e-5 * float(ref_cu.abs().max()), (d, rep)
        assert float((cv - ref_cv).abs().max()) <= 5e-5 * float(ref_cv.abs().max()), (d, rep)
        # dv/dt is NOT continuous: a pre-activation within rounding of 0 flips its relu mask and moves the tangent of that
        # point by O(weight) -- a few points in 10^5 between any two fp32 evaluation orders (3e8 pre-activations here)
        off = ((vc[:, 1] - ref_vc[:, 1]).abs() > 5e-5 * ts).float().mean().item()
        assert off <= 5e-4, (d, rep, off)
        assert np.allclose(s[:5], ref_s[:5], rtol=1e-4, atol=1e-4 * np.abs(ref_s[:5]).max())
