"""TEST INFRASTRUCTURE ONLY -- PyTorch/CPU float64 port of the reference's hot path, performing the
SAME sequence of tensor operations the reference performs (autograd included), so that it can stand
in for the reference's CPU implementation where /root/reference does not exist (the GPU box):

  * bench.py `cpu_baseline` and `--impl reference` time it on the host cores;
  * tests pin it against the golden vectors of the unmodified reference (tests/test_oracle.py).

Restated from (all under /root/reference): src/model.py:37-47 (test-function net), :87-112,
:133-141,:153-156 (XNODE), torchdiffeq 0.1.1 fixed-grid midpoint/euler/rk4 (third party; see
oracle/shims/torchdiffeq), src/training.py:25-41 (dense coefficient tensors, d^2 Python loop),
src/loss.py:46-96 (weak form with its two helper backward calls and the in-place zeroing of the
input gradients), src/training.py:129-137 / :153-161 (one u-step / v-step).
Never imported by the product package.
"""
import itertools
import math

import torch
import torch.nn.functional as F


def make_params(d, H=20, hh=10, Hv=50, seed=0, dtype=torch.float64):
    """xavier-uniform weights, zero biases, in the reference's parameter order (14 + 6 tensors)"""
    g = torch.Generator().manual_seed(seed)

    def xav(o, i):
        b = math.sqrt(6.0 / (i + o))
        return ((torch.rand(o, i, generator=g, dtype=dtype) * 2 - 1) * b).requires_grad_(True)

    def zb(n):
        return torch.zeros(n, dtype=dtype, requires_grad=True)

    pu = [xav(H, 1), zb(H), xav(H, H), zb(H), xav(H, H), zb(H), xav(hh, H + d + 1), zb(hh), xav(hh, hh), zb(hh),
          xav(H, hh), zb(H), xav(1, H), zb(1)]
    pv = [xav(Hv, d + 1), zb(Hv), xav(Hv, Hv), zb(Hv), xav(1, Hv), zb(1)]
    return pu, pv


def v_net(pv, XV, nv):
    Wi, bi, Wh, bh, Wz, bz = pv
    a = F.linear(XV.double(), Wi, bi)
    for _ in range(nv):
        a = F.linear(torch.relu(a), Wh, bh)
    return F.linear(torch.tanh(a), Wz, bz)


def _field(pu, x, nu):
    Wa, ba, Ws, bs, Wf, bf = pu[6:12]

    def rhs(t, y):
        z = torch.cat((x, t.repeat(y.shape[0], 1), y), dim=1)
        a = F.linear(z, Wa, ba)
        for _ in range(nu - 1):
            a = F.linear(torch.relu(a), Ws, bs)
        return F.linear(torch.tanh(a), Wf, bf)
    return rhs


def _odeint(rhs, y0, t, method):
    t = t.type_as(y0)
    ys, y = [y0], y0
    for i in range(t.numel() - 1):
        t0, dt = t[i], t[i + 1] - t[i]
        if method == "euler":
            dy = dt * rhs(t0, y)
        elif method == "midpoint":
            ym = y + rhs(t0, y) * dt / 2
            dy = dt * rhs(t0 + dt / 2, ym)
        elif method == "rk4":
            k1 = rhs(t0, y)
            k2 = rhs(t0 + dt / 3, y + dt * k1 / 3)
            k3 = rhs(t0 + dt * 2 / 3, y + dt * (k2 - k1 / 3))
            k4 = rhs(t0 + dt, y + dt * (k1 - k2 + k3))
            dy = dt * (k1 + 3 * (k2 + k3) + k4) / 8
        else:
            raise ValueError(method)
        y = y + dy
        ys.append(y)
    return torch.stack(ys, 0)


def u_net(pu, X, func_h, nu, solver="midpoint"):
    """batch starting at T0 (the cube's interior and boundary batches)"""
    W0, b0, W1, b1, W2, b2 = pu[:6]
    Wo, bo = pu[12:14]
    s = func_h(X[:, 0, :]).unsqueeze(1).double()
    y0 = F.linear(torch.relu(F.linear(torch.relu(F.linear(s, W0, b0)), W1, b1)), W2, b2)
    out = _odeint(_field(pu, X[:, 0, 1:], nu), y0, X[0, :, 0], solver).transpose(0, 1)
    return F.linear(out, Wo, bo)


def cube_w(X, bot, top):
    s = X[:, :, 1:]
    return torch.minimum(torch.min(torch.abs(top - s), dim=2).values, torch.min(torch.abs(bot - s), dim=2).values)


def step(phase, pu, pv, X, XV, BX, prob, cfg):
    """one u-step or v-step on fresh leaf copies; fills .grad of the phase's parameters.
    cfg: dict(nu, nv, solver, alpha, bot, top, V).  Returns (loss value, components)."""
    d = X.shape[2] - 1
    X = X.clone().detach().requires_grad_(True)
    XV = XV.clone().detach().requires_grad_(True)
    BX = BX.clone().detach().requires_grad_(True)
    for p in pu + pv:
        p.grad = None
    pred_v = v_net(pv, XV, cfg["nv"])
    pred_u = u_net(pu, X, prob.func_h, cfg["nu"], cfg["solver"])
    Xd, BXd = X.clone().detach(), BX.clone().detach()
    h, f, g = prob.func_h(Xd[:, 0, :]), prob.func_f(Xd), prob.func_g(BXd)
    c = prob.func_c(Xd, pred_u)
    a = torch.empty(d, d, X.shape[0], X.shape[1])
    for i, j in itertools.product(range(d), repeat=2):
        a[i, j] = prob.func_a(Xd, i, j)
    b = torch.empty(d, X.shape[0], X.shape[1])
    for i in range(d):
        b[i] = prob.func_b(Xd, i)
    V, alpha = cfg["V"], cfg["alpha"]
    N, Nt = pred_u.shape[0], pred_u.shape[1]
    w = cube_w(XV, cfg["bot"], cfg["top"]).unsqueeze(2)
    phi = pred_v * w
    pred_u.backward(torch.ones_like(pred_u), retain_graph=True)
    du = [X.grad[:, :, i] for i in range(d + 1)]
    phi.backward(torch.ones_like(phi), retain_graph=True)
    dphi = [XV.grad[:, :, i] for i in range(d + 1)]
    us, vs, ph = pred_u.squeeze(), pred_v.squeeze(), phi.squeeze()
    s1 = V * (pred_u[:, -1].squeeze() * pred_v[:, -1].squeeze() - h * pred_v[:, 0].squeeze()) / N
    s2 = V * (us * dphi[0]) / N / Nt
    s31 = torch.stack([a[i, j] * dphi[i + 1] * du[j + 1] for i, j in itertools.product(range(d), repeat=2)], 0).sum(0)
    s32 = torch.stack([b[i] * ph * du[i + 1] for i in range(d)], 0).sum(0)
    s3 = (V / N / Nt) * (s31 + s32 + c.squeeze() * us * ph + f * ph)
    I = torch.sum(s1 - torch.sum(s2 - s3, 1), 0)
    X.grad.data.zero_()
    XV.grad.data.zero_()
    S = V * torch.sum(pred_v ** 2) / (N * Nt)
    integ = torch.log(I ** 2) - torch.log(S)
    comps = {"I": float(I.detach()), "S": float(S.detach())}
    if phase == "u":
        init = torch.mean((pred_u[:, 0] - h.unsqueeze(1)) ** 2)
        bdry = torch.mean((u_net(pu, BX, prob.func_h, cfg["nu"], cfg["solver"]) - g.unsqueeze(2)) ** 2)
        val = integ + alpha * (init + bdry)
        comps.update(init=float(init.detach()), bdry=float(bdry.detach()))
    else:
        val = -integ
    val.backward(retain_graph=True)
    return float(val.detach()), comps
