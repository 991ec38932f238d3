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
