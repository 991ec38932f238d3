"""CPU: the closed-form oracle (oracle/closed_form.py) against the golden vectors produced by the
unmodified reference, plus known-answer tests pinning the restated torchdiffeq fixed-grid scheme."""
import numpy as np
import pytest

from oracle import closed_form as cf
from tests import _golden as G


@pytest.mark.parametrize("name", G.names())
def test_closed_form_matches_reference_golden(name):
    c = G.load(name)
    z = c["z"]
    for phase, gold_loss, gold_g in (("u", float(z["loss_u"]), c["gu"]), ("v", float(z["loss_v"]), c["gv"])):
        r = cf.weak_form(c["thu"], c["thv"], z["X"], z["XV"], z["BX"], c["coef"], c["cfg"], phase)
        # the reference itself rounds du/dphi to fp32 (X.grad dtype) => ~1e-7 on I
        assert abs(r["I"] - float(z["I"])) <= 2e-6 * abs(float(z["I"])) + 1e-12
        assert abs(r["S"] - float(z["S"])) <= 1e-12 * abs(float(z["S"]))
        assert abs(r["init"] - float(z["init"])) <= 1e-12 * abs(float(z["init"])) + 1e-15
        assert abs(r["bdry"] - float(z["bdry"])) <= 1e-10 * abs(float(z["bdry"])) + 1e-15
        assert abs(r["loss_" + phase] - gold_loss) <= 1e-6 * abs(gold_loss) + 1e-6
        assert np.abs(r["u"] - z["u"]).max() < 1e-12
        assert np.abs(r["v"] - z["v"]).max() < 1e-12
        assert np.abs(r["du"] - z["du"][:, 0, 1:]).max() <= 1e-6 * np.abs(z["du"]).max() + 1e-9
        assert np.abs(r["dphi"] - z["dphi"]).max() <= 1e-6 * np.abs(z["dphi"]).max() + 1e-9
        for a, b in zip(r["grads"], gold_g):
            assert G.rel(a, b) < 5e-6, (phase, G.rel(a, b))


def test_reference_du_structure_in_golden():
    """X.grad is zero for time rows l>=1 (x is read from row 0 only)"""
    z = G.load("cube_d5_shipped_small")["z"]
    assert np.abs(z["du"][:, 1:, 1:]).max() == 0.0


@pytest.mark.parametrize("solver,order", [("euler", 1), ("midpoint", 2), ("rk4", 4)])
def test_fixed_grid_scheme_known_answer_and_order(solver, order):
    """y' = lam*y through the field machinery (nu=1, tanh linearised by tiny weights is awkward, so
    test the step functions directly with a linear field)"""
    lam = -0.7

    def run(n):
        t = np.linspace(0.0, 1.0, n + 1)
        y = 1.0
        for i in range(n):
            dt = t[i + 1] - t[i]
            f = lambda tt, yy: lam * yy
            if solver == "euler":
                y = y + dt * f(t[i], y)
            elif solver == "midpoint":
                ym = y + f(t[i], y) * dt / 2
                y = y + dt * f(t[i] + dt / 2, ym)
            else:
                k1 = f(t[i], y)
                k2 = f(t[i] + dt / 3, y + dt * k1 / 3)
                k3 = f(t[i] + dt * 2 / 3, y + dt * (k2 - k1 / 3))
                k4 = f(t[i] + dt, y + dt * (k1 - k2 + k3))
                y = y + dt * (k1 + 3 * (k2 + k3) + k4) / 8
        return y
    # closed form of one midpoint step: y1 = y0 (1 + z + z^2/2)
    if solver == "midpoint":
        z = lam * 1.0
        assert abs(run(1) - (1 + z + z * z / 2)) < 1e-15
    e1, e2 = abs(run(20) - np.exp(lam)), abs(run(40) - np.exp(lam))
    assert abs(np.log2(e1 / e2) - order) < 0.15


def test_shim_odeint_matches_closed_form_steps():
    """the torchdiffeq stand-in used to make the golden vectors == the oracle's own stepping"""
    import os
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shims"))
    import torchdiffeq
    rng = np.random.default_rng(0)
    thu, _ = cf.xavier_theta(3, H=6, hh=5, seed=1)
    for k in thu:
        thu[k] = thu[k] + 0.1 * rng.standard_normal(thu[k].shape)
    x = rng.uniform(-1, 1, (4, 3))
    times = np.sort(rng.uniform(0, 1, 6))
    s = rng.standard_normal(4)
    for solver in ("euler", "midpoint", "rk4"):
        u, cache = cf.xnode_forward(thu, x, times, s, nu=3, solver=solver)
        ax = torch.tensor(x @ thu["Wa"][:, :3].T + thu["ba"])

        def field(t, y):
            out, _ = cf.field_fwd(thu, ax.numpy(), float(t), y.numpy(), 3)
            return torch.tensor(out)
        ys = torchdiffeq.odeint(field, torch.tensor(cache["Y"][:, 0]), torch.tensor(times), method=solver)
        assert np.abs(ys.numpy().transpose(1, 0, 2) - cache["Y"]).max() < 1e-13


def test_xnode_vjp_matches_finite_differences():
    rng = np.random.default_rng(3)
    thu, _ = cf.xavier_theta(2, H=5, hh=4, seed=2)
    for k in thu:
        thu[k] = thu[k] + 0.2 * rng.standard_normal(thu[k].shape)
    x = rng.uniform(-1, 1, (3, 2))
    times = np.array([0.0, 0.3, 0.55, 1.0])
    s = rng.standard_normal(3)
    Gc = rng.standard_normal((3, 4))
    for solver in ("euler", "midpoint", "rk4"):
        u, cache = cf.xnode_forward(thu, x, times, s, 3, solver)
        g, gx, gs = cf.xnode_vjp(thu, cache, Gc)
        for key in ("Ws", "Wa", "W1", "bf", "W0"):
            idx = tuple(rng.integers(0, n) for n in thu[key].shape)
            eps = 1e-6
            tp = {k: v.copy() for k, v in thu.items()}
            tp[key][idx] += eps
            up, _ = cf.xnode_forward(tp, x, times, s, 3, solver)
            tp[key][idx] -= 2 * eps
            um, _ = cf.xnode_forward(tp, x, times, s, 3, solver)
            fd = ((up - um) * Gc).sum() / (2 * eps)
            assert abs(fd - g[key][idx]) < 1e-6 * max(1.0, abs(fd)), (solver, key)
        xp = x.copy(); xp[1, 0] += 1e-6
        xm = x.copy(); xm[1, 0] -= 1e-6
        fd = ((cf.xnode_forward(thu, xp, times, s, 3, solver)[0] - cf.xnode_forward(thu, xm, times, s, 3, solver)[0]) * Gc).sum() / 2e-6
        assert abs(fd - gx[1, 0]) < 1e-6 * max(1.0, abs(fd))


def test_domain_w_gradients_fd():
    rng = np.random.default_rng(5)
    P = np.concatenate([rng.uniform(0.05, 0.45, (6, 4, 1)), rng.uniform(-0.3, 0.3, (6, 4, 3))], -1)
    for dom in (("cube", -1.0, 1.0), ("cone", 1.0), ("hourglass", 1.0, 0.0, 1.0)):
        w, dw = cf.domain_w(dom, P)
        for c in range(4):
            Pp = P.copy(); Pp[..., c] += 1e-7
            Pm = P.copy(); Pm[..., c] -= 1e-7
            fd = (cf.domain_w(dom, Pp)[0] - cf.domain_w(dom, Pm)[0]) / 2e-7
            assert np.abs(fd - dw[..., c]).max() < 1e-6


@pytest.mark.parametrize("name", ["cube_d5_alpha1_randbias", "cube_d4_ex43", "cube_d3_rk4", "cube_d2_L2"])
def test_torch_port_matches_reference_golden(name):
    """the PyTorch/CPU port used as cpu_baseline and `--impl reference` stand-in reproduces the
    unmodified reference's loss values and parameter gradients (same ops, float64)"""
    import torch
    import xnode_wan_b200 as xw
    from oracle import torch_port as tp
    c = G.load(name)
    z, p = c["z"], c["params"]
    prob = xw.problems.by_name(c["meta"]["funcs"], p["dim"])
    pu = [torch.tensor(np.asarray(w), dtype=torch.float64, requires_grad=True) for w in c["thu_list"]]
    pv = [torch.tensor(np.asarray(w), dtype=torch.float64, requires_grad=True) for w in c["thv_list"]]
    cfg = dict(nu=p["u_layers"], nv=p["v_layers"], solver=p["solver"], alpha=p["alpha"], bot=c["meta"]["domain"][1],
               top=c["meta"]["domain"][2], V=c["meta"]["V"])
    X, XV, BX = (torch.from_numpy(z[k]) for k in ("X", "XV", "BX"))
    lu, comps = tp.step("u", pu, pv, X, XV, BX, prob, cfg)
    assert abs(lu - float(z["loss_u"])) <= 1e-6 * abs(float(z["loss_u"])) + 1e-6
    assert abs(comps["I"] - float(z["I"])) <= 1e-5 * abs(float(z["I"]))
    for q, gref in zip(pu, c["gu"]):
        assert G.rel(q.grad.numpy(), gref) < 1e-5
    lv, _ = tp.step("v", pu, pv, X, XV, BX, prob, cfg)
    assert abs(lv - float(z["loss_v"])) <= 1e-5 * abs(float(z["loss_v"])) + 1e-6
    for q, gref in zip(pv, c["gv"]):
        assert G.rel(q.grad.numpy(), gref) < 1e-5


@pytest.mark.parametrize("name", G.big_names())
def test_closed_form_matches_full_size_reference_golden(name):
    """the oracle pinned at the BASELINE sizes (shipped d=5 N_r=N_b=4000 with seeds 0/0 -- the SURVEY.md Appendix A.6
    anchors -- and d=20 with 4096 / 8192 paths): outputs of the UNMODIFIED reference in tests/golden/big"""
    c = G.load_big(name)
    z = c["z"]
    K = z["u_head"].shape[0]
    for phase, gold_loss, gold_g in (("u", float(z["loss_u"]), c["gu"]), ("v", float(z["loss_v"]), c["gv"])):
        r = cf.weak_form(c["thu"], c["thv"], z["X"], z["XV"], z["BX"], c["coef"], c["cfg"], phase)
        assert abs(r["I"] - float(z["I"])) <= 2e-6 * abs(float(z["I"])) + 1e-12
        assert abs(r["S"] - float(z["S"])) <= 1e-10 * abs(float(z["S"]))
        assert abs(r["init"] - float(z["init"])) <= 1e-10 * abs(float(z["init"]))
        assert abs(r["bdry"] - float(z["bdry"])) <= 1e-8 * abs(float(z["bdry"]))
        assert abs(r["loss_" + phase] - gold_loss) <= 1e-6 * abs(gold_loss) + 1e-6
        assert np.abs(r["u"][:K] - z["u_head"]).max() < 1e-11
        assert np.abs(r["v"][:K] - z["v_head"]).max() < 1e-11
        assert np.abs(r["du"][:K] - z["du_head"]).max() <= 1e-6 * np.abs(z["du_head"]).max() + 1e-9
        for a, b in zip(r["grads"], gold_g):
            assert G.rel(a, b) < 5e-6, (phase, G.rel(a, b))


def test_shipped_full_golden_is_the_survey_anchor():
    """seeds 0/0, cube_pde.yaml N_r=N_b=4000, Ex4_1: the values SURVEY.md Appendix A.6 recorded from the reference"""
    z = np.load(G.os.path.join(G.BIG_DIR, "cube_d5_shipped_full.npz"))
    assert abs(float(z["I"]) - 0.0726511181) < 1e-9
    assert abs(float(z["S"]) - 0.00103209205) < 1e-10
    assert abs(float(z["init"]) - 0.981500739) < 1e-8
    assert abs(float(z["bdry"]) - 0.500196223) < 1e-8
    assert abs(float(z["loss_u"]) - 1.4816969779e8) < 1.0
    assert abs(float(z["loss_v"]) + 1.631994) < 1e-5
