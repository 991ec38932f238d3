"""GPU (-m gpu): the real sm_100a library through the C ABI against the golden vectors of the
unmodified reference (bit-for-bit inputs, fp32 kernels vs the reference's fp64 PyTorch path).
Tolerances are BASELINE.json's: rel 1e-4 on loss values, 1e-3 (rel-L2 per tensor) on gradients."""
import numpy as np
import pytest

import xnode_wan_b200 as xw
from oracle import closed_form as cf
from tests import _golden as G
from tests import _lowlevel as LL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return xw._lib.get()


@pytest.mark.parametrize("name", G.names())
def test_capi_matches_reference_golden(lib, name):
    c = G.load(name)
    z = c["z"]
    r = LL.run_case(lib, LL.TorchBackend(), c)
    assert np.abs(r["u"] - z["u"]).max() < 2e-5
    for k in ("I", "S", "init", "bdry"):
        assert abs(r[k] - float(z[k])) <= 1e-4 * abs(float(z[k])) + 1e-9, (k, r[k], float(z[k]))
    for k in ("loss_u", "loss_v"):
        assert abs(r[k] - float(z[k])) <= 1e-4 * abs(float(z[k])) + 1e-6, (k, r[k], float(z[k]))
    for i, (a, b) in enumerate(zip(r["grads_u"], c["gu"])):
        assert G.rel(a, b) < 1e-3, ("grad_u", i, G.rel(a, b))
    for i, (a, b) in enumerate(zip(r["grads_v"], c["gv"])):
        assert G.rel(a, b) < 1e-3, ("grad_v", i, G.rel(a, b))


@pytest.mark.parametrize("name", G.big_names())
def test_capi_matches_full_size_reference_golden(lib, name):
    """BASELINE configs[0]/[1] at the shipped size (d=5, N_r=N_b=4000, seeds 0/0: SURVEY.md Appendix A.6 anchors) and
    d=20 with N = 4096 / 8192 paths: outputs of the UNMODIFIED reference (tests/golden/big), inputs re-created with the
    bit-identical sampler and SHA-1 checked.  Whole waves of tensor-core tiles, not a handful."""
    c = G.load_big(name)
    z = c["z"]
    r = LL.run_case(lib, LL.TorchBackend(), c)
    assert np.abs(r["u"][:z["u_head"].shape[0]] - z["u_head"]).max() < 2e-5
    for k in ("I", "S", "init", "bdry"):
        assert abs(r[k] - float(z[k])) <= 1e-4 * abs(float(z[k])) + 1e-9, (k, r[k], float(z[k]))
    for k in ("loss_u", "loss_v"):
        assert abs(r[k] - float(z[k])) <= 1e-4 * abs(float(z[k])) + 1e-6, (k, r[k], float(z[k]))
    for i, (a, b) in enumerate(zip(r["grads_u"], c["gu"])):
        assert G.rel(a, b) < 1e-3, ("grad_u", i, G.rel(a, b))
    for i, (a, b) in enumerate(zip(r["grads_v"], c["gv"])):
        assert G.rel(a, b) < 1e-3, ("grad_v", i, G.rel(a, b))
    assert (lib.cdll.xw_last_xnode_impl() >> 4) & 15 >= 2          # generation 2 / 3 XNODE kernels, not the generation-1 fallback


def test_capi_dense_a_b_matches_oracle(lib):
    c = G.load("cube_d3_small_nets")
    rng = np.random.default_rng(0)
    a = np.eye(3) + 0.3 * rng.standard_normal((3, 3))
    b = rng.standard_normal(3)
    coef = dict(c["coef"], a=a, b=b)
    z = c["z"]
    ru = cf.weak_form(c["thu"], c["thv"], z["X"], z["XV"], z["BX"], coef, c["cfg"], "u")
    rv = cf.weak_form(c["thu"], c["thv"], z["X"], z["XV"], z["BX"], coef, c["cfg"], "v")
    r = LL.run_case(lib, LL.TorchBackend(), c, coef_a=a, coef_b=b)
    assert abs(r["I"] - ru["I"]) <= 1e-4 * abs(ru["I"])
    for x, y in zip(r["grads_u"], ru["grads"]):
        assert G.rel(x, y) < 1e-3
    for x, y in zip(r["grads_v"], rv["grads"]):
        assert G.rel(x, y) < 1e-3


def test_loss_scalars_kernel_against_fp64_arithmetic(lib):
    """xw_loss_scalars on the device against the same formulas in numpy fp64 (reference src/loss.py:64-96; the cotangent
    coefficients of the backward entries), both phases, with and without a boundary batch"""
    import ctypes as C
    import torch
    rng = np.random.default_rng(3)
    for phase, nb in ((0, 4000.0), (1, 0.0), (0, 0.0)):
        s = rng.random(8) + 0.1
        s[0] -= 0.6
        V, n, L, Lb, alpha, side = 2.0 ** 20, 1048576.0, 20.0, 20.0, 1e8, 1.0
        sums = torch.tensor(s, dtype=torch.float64, device="cuda:0")
        out = torch.zeros(8, dtype=torch.float64, device="cuda:0")
        lib.call("xw_loss_scalars", C.c_void_p(sums.data_ptr()), phase, V, n, L, nb, Lb, alpha, side, C.c_void_p(out.data_ptr()),
                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
        got = out.cpu().numpy()
        I = (V / n) * s[0] - (V / (n * L)) * (s[1] - s[2])
        S = V * s[3] / (n * L)
        init, bdry = s[4] / n, (s[5] / (nb * Lb) if nb else 0.0)
        integ = np.log(I * I) - np.log(S)
        want = [integ + alpha * (init + bdry), I, S, init, bdry, (2 / I) * (V / (n * L)), 2 * alpha / n, side] if phase == 0 else \
               [-integ, I, S, init, bdry, -(2 / I) * (V / (n * L)), 2 / s[3], side]
        for a, w in zip(got, want):
            assert abs(a - w) <= 1e-13 * abs(w) + 1e-300, (phase, nb, a, w)
