"""TEST INFRASTRUCTURE ONLY -- numpy/float64 closed-form restatement of the XNODE-WAN hot path.

This is the CHECKER for the CUDA path (tests/, __graft_entry__.smoke(), bench.py cpu_baseline).
It is never imported by the product package.  No autograd: forward passes and hand-derived
reverse sweeps, i.e. exactly the arithmetic the kernels implement, so that every kernel stage has
a line here to be compared with.

Reference lines restated (all under /root/reference):
  * XNODE primal net            src/model.py:87-112 (forward), :133-141 (_ODEField), :153-156 (_F)
  * fixed-grid ODE scheme       torchdiffeq==0.1.1 (third party, not vendored; see
                                oracle/shims/torchdiffeq/__init__.py) -- call site src/model.py:103-106
  * test-function net           src/model.py:37-47
  * domain weight w             src/dataset.py:278-282 (cube), :216-218 (cone), :119-125 (hourglass)
  * weak form I, init, bdry, int, u, v        src/loss.py:46-96
  * what the optimisers see     src/training.py:127,137-138,152,161-162  

Pinned against golden vectors produced by the unmodified reference (tests/golden/make_golden.py,
tests/test_oracle.py).  The torchdiffeq boundary itself is "parity unpinned" (no reference test
pins it); it is pinned by closed-form known answers only.

Parameter containers are plain dicts of float64 arrays:
  theta_u: W0[H,1] b0[H] W1[H,H] b1[H] W2[H,H] b2[H]      (initial_layers.0/.2/.4)
           Wa[hh,d+1+H] ba[hh]                               (ODE_rhs.net.0; input order (x, t, y))
           Ws[hh,hh] bs[hh]                                  (ODE_rhs.net.2, shared (nu-1) times)
           Wf[H,hh] bf[H]                                    (ODE_rhs.net.<2*nu>)
           Wo[1,H] bo[1]                                     (final_linear)
  theta_v: Wi[Hv,C] bi[Hv] Wh[Hv,Hv] bh[Hv] Wz[1,Hv] bz[1]  (input, hidden shared nv times, output)
"""
import numpy as np

U_KEYS = ("W0", "b0", "W1", "b1", "W2", "b2", "Wa", "ba", "Ws", "bs", "Wf", "bf", "Wo", "bo")
V_KEYS = ("Wi", "bi", "Wh", "bh", "Wz", "bz")


# ----------------------------------------------------------------------------- XNODE vector field
def field_fwd(th, ax, t, y, nu):
    """F(t, y) = net(cat(x, t, y)); ax = Wa[:, :d] @ x + ba is the path-constant part.
    returns (out[N,H], cache)"""
    d1 = th["Wa"].shape[1] - y.shape[1]          # d + 1
    wt = th["Wa"][:, d1 - 1]
    Wy = th["Wa"][:, d1:]
    a = ax + np.outer(t, wt) if np.ndim(t) else ax + t * wt
    a = a + y @ Wy.T
    acts = [a]
    for _ in range(nu - 1):
        a = np.maximum(a, 0.0) @ th["Ws"].T + th["bs"]
        acts.append(a)
    tau = np.tanh(a)
    out = tau @ th["Wf"].T + th["bf"]
    return out, (acts, tau, t, y)


def field_vjp(th, cache, gout, g, nu):
    """reverse sweep of one field evaluation.
    gout[N,H] cotangent of F; accumulates parameter grads into g (dict) when g is not None.
    returns (gy[N,H], ga0[N,hh], gt[N])  -- ga0 is the cotangent of the first pre-activation
    (its sum over evaluations gives the x-gradient: Wa[:, :d]^T ga0)."""
    acts, tau, t, y = cache
    H = y.shape[1]
    d1 = th["Wa"].shape[1] - H
    if g is not None:
        g["Wf"] += gout.T @ tau
        g["bf"] += gout.sum(0)
    da = (gout @ th["Wf"]) * (1.0 - tau * tau)
    for k in range(nu - 1, 0, -1):
        r = np.maximum(acts[k - 1], 0.0)
        if g is not None:
            g["Ws"] += da.T @ r
            g["bs"] += da.sum(0)
        da = (da @ th["Ws"]) * (acts[k - 1] > 0)
    if g is not None:
        g["ba"] += da.sum(0)
        g["Wa"][:, d1:] += da.T @ y
        tt = np.broadcast_to(t, (y.shape[0],))
        g["Wa"][:, d1 - 1] += da.T @ tt
    gy = da @ th["Wa"][:, d1:]
    gt = da @ th["Wa"][:, d1 - 1]
    return gy, da, gt


_STAGES = {"euler": 1, "midpoint": 2, "rk4": 4}


def _step_fwd(th, ax, t0, dt, y, nu, solver):
    """one fixed-grid step; returns (y1, caches)"""
    if solver == "euler":
        k1, c1 = field_fwd(th, ax, t0, y, nu)
        return y + dt * k1, (c1,)
    if solver == "midpoint":
        k1, c1 = field_fwd(th, ax, t0, y, nu)
        ym = y + k1 * dt / 2
        k2, c2 = field_fwd(th, ax, t0 + dt / 2, ym, nu)
        return y + dt * k2, (c1, c2)
    if solver == "rk4":
        k1, c1 = field_fwd(th, ax, t0, y, nu)
        k2, c2 = field_fwd(th, ax, t0 + dt / 3, y + dt * k1 / 3, nu)
        k3, c3 = field_fwd(th, ax, t0 + dt * 2 / 3, y + dt * (k2 - k1 / 3), nu)
        k4, c4 = field_fwd(th, ax, t0 + dt, y + dt * (k1 - k2 + k3), nu)
        return y + dt * (k1 + 3 * (k2 + k3) + k4) / 8, (c1, c2, c3, c4)
    raise ValueError(solver)


def _step_vjp(th, caches, dt, lam, g, nu, solver):
    """reverse of _step_fwd.  lam = cotangent of y1.  returns (cotangent of y0, sum of ga0)"""
    if solver == "euler":
        gy, ga0, _ = field_vjp(th, caches[0], dt * lam, g, nu)
        return lam + gy, ga0
    if solver == "midpoint":
        gym, ga2, _ = field_vjp(th, caches[1], dt * lam, g, nu)
        gy, ga1, _ = field_vjp(th, caches[0], gym * (dt / 2), g, nu)
        return lam + gym + gy, ga1 + ga2
    if solver == "rk4":
        k1b = dt * lam / 8
        k2b = 3 * dt * lam / 8
        k3b = 3 * dt * lam / 8
        k4b = dt * lam / 8
        y0b = lam.copy()
        g4, a4, _ = field_vjp(th, caches[3], k4b, g, nu)       # input y + dt*(k1 - k2 + k3)
        y0b += g4
        k1b = k1b + dt * g4
        k2b = k2b - dt * g4
        k3b = k3b + dt * g4
        g3, a3, _ = field_vjp(th, caches[2], k3b, g, nu)       # input y + dt*(k2 - k1/3)
        y0b += g3
        k2b = k2b + dt * g3
        k1b = k1b - dt * g3 / 3
        g2, a2, _ = field_vjp(th, caches[1], k2b, g, nu)       # input y + dt*k1/3
        y0b += g2
        k1b = k1b + dt * g2 / 3
        g1, a1, _ = field_vjp(th, caches[0], k1b, g, nu)
        y0b += g1
        return y0b, a1 + a2 + a3 + a4
    raise ValueError(solver)


# ----------------------------------------------------------------------------------- XNODE
def xnode_forward(th, x, times, s, nu, solver="midpoint"):
    """src/model.py:87-112 for a batch whose time grid is `times` (path 0's times).
    x[N,d] spatial coords of time-row 0, s[N] initial scalar (h(x) or g(x)).
    returns (u[N,L], cache)"""
    x = np.asarray(x, np.float64)
    times = np.asarray(times, np.float64)
    s = np.asarray(s, np.float64)
    d = x.shape[1]
    p1 = np.outer(s, th["W0"][:, 0]) + th["b0"]
    z1 = np.maximum(p1, 0.0)
    p2 = z1 @ th["W1"].T + th["b1"]
    z2 = np.maximum(p2, 0.0)
    y = z2 @ th["W2"].T + th["b2"]
    ax = x @ th["Wa"][:, :d].T + th["ba"]
    ys = [y]
    steps = []
    for l in range(len(times) - 1):
        dt = times[l + 1] - times[l]
        y, caches = _step_fwd(th, ax, times[l], dt, y, nu, solver)
        ys.append(y)
        steps.append((dt, caches))
    Y = np.stack(ys, 1)                                     # [N,L,H]
    u = Y @ th["Wo"][0] + th["bo"][0]
    return u, dict(x=x, s=s, p1=p1, z1=z1, p2=p2, z2=z2, Y=Y, steps=steps, nu=nu, solver=solver)


def xnode_vjp(th, cache, G, want_param_grads=True):
    """VJP of u[N,L] at cotangent G[N,L].
    returns (grads dict or None, gx[N,d] = dSum(G*u)/dx through the field only, gs[N] = d/ds)"""
    nu, solver = cache["nu"], cache["solver"]
    Y = cache["Y"]
    N, L, H = Y.shape
    d = cache["x"].shape[1]
    g = {k: np.zeros_like(th[k]) for k in U_KEYS} if want_param_grads else None
    if g is not None:
        g["Wo"][0] += np.einsum("nl,nlh->h", G, Y)
        g["bo"][0] += G.sum()
    lam = np.outer(G[:, L - 1], th["Wo"][0])
    A0 = np.zeros((N, th["Wa"].shape[0]))
    for l in range(L - 2, -1, -1):
        dt, caches = cache["steps"][l]
        lam, ga0 = _step_vjp(th, caches, dt, lam, g, nu, solver)
        A0 += ga0
        lam = lam + np.outer(G[:, l], th["Wo"][0])
    if g is not None:
        g["Wa"][:, :d] += A0.T @ cache["x"]
    gx = A0 @ th["Wa"][:, :d]
    # lift 1 -> H -> H -> H
    if g is not None:
        g["W2"] += lam.T @ cache["z2"]
        g["b2"] += lam.sum(0)
    dz2 = (lam @ th["W2"]) * (cache["p2"] > 0)
    if g is not None:
        g["W1"] += dz2.T @ cache["z1"]
        g["b1"] += dz2.sum(0)
    dz1 = (dz2 @ th["W1"]) * (cache["p1"] > 0)
    if g is not None:
        g["W0"][:, 0] += dz1.T @ cache["s"]
        g["b0"] += dz1.sum(0)
    gs = dz1 @ th["W0"][:, 0]
    return g, gx, gs


# ------------------------------------------------------------------------------ test-function net
def vnet_forward(th, P, nv):
    """src/model.py:37-47 pointwise on P[..., C]; returns (v[...], cache)"""
    P = np.asarray(P, np.float64)
    a = P @ th["Wi"].T + th["bi"]
    acts = [a]
    for _ in range(nv):
        a = np.maximum(a, 0.0) @ th["Wh"].T + th["bh"]
        acts.append(a)
    tau = np.tanh(a)
    v = tau @ th["Wz"][0] + th["bz"][0]
    return v, (P, acts, tau)


def vnet_vjp(th, cache, G, nv, want_param_grads=True):
    """VJP of v at cotangent G[...]; returns (grads or None, dP[..., C])"""
    P, acts, tau = cache
    C = P.shape[-1]
    Pf = P.reshape(-1, C)
    Gf = np.asarray(G, np.float64).reshape(-1)
    tauf = tau.reshape(Pf.shape[0], -1)
    g = {k: np.zeros_like(th[k]) for k in V_KEYS} if want_param_grads else None
    if g is not None:
        g["Wz"][0] += Gf @ tauf
        g["bz"][0] += Gf.sum()
    da = np.outer(Gf, th["Wz"][0]) * (1.0 - tauf * tauf)
    for k in range(nv, 0, -1):
        pre = acts[k - 1].reshape(Pf.shape[0], -1)
        if g is not None:
            g["Wh"] += da.T @ np.maximum(pre, 0.0)
            g["bh"] += da.sum(0)
        da = (da @ th["Wh"]) * (pre > 0)
    if g is not None:
        g["Wi"] += da.T @ Pf
        g["bi"] += da.sum(0)
    dP = (da @ th["Wi"]).reshape(P.shape)
    return g, dP


# --------------------------------------------------------------------------------- domain weights
def domain_w(domain, P):
    """w[...] and dw/dP[..., C] (time in channel 0) for
    domain = ("cube", bot, top) | ("cone", r) | ("hourglass", r, T0, T)."""
    P = np.asarray(P, np.float64)
    x = P[..., 1:]
    t = P[..., 0]
    dw = np.zeros_like(P)
    kind = domain[0]
    if kind == "cube":
        bot, top = float(domain[1]), float(domain[2])
        dtop = np.abs(top - x)
        dbot = np.abs(bot - x)
        itop = dtop.argmin(-1)
        ibot = dbot.argmin(-1)
        mtop = np.take_along_axis(dtop, itop[..., None], -1)[..., 0]
        mbot = np.take_along_axis(dbot, ibot[..., None], -1)[..., 0]
        use_top = mtop <= mbot
        w = np.where(use_top, mtop, mbot)
        idx = np.where(use_top, itop, ibot)
        xc = np.take_along_axis(x, idx[..., None], -1)[..., 0]
        gsel = np.where(use_top, -np.sign(top - xc), -np.sign(bot - xc))
        np.put_along_axis(dw[..., 1:], idx[..., None], gsel[..., None], -1)
        return w, dw
    r = float(domain[1])
    nrm = np.sqrt((x * x).sum(-1))
    if kind == "cone":
        w = r * (1.0 - t) - nrm
        dw[..., 0] = -r
    elif kind == "hourglass":
        T0, T = float(domain[2]), float(domain[3])
        first = t <= (T - T0) / 2
        w = np.where(first, r * ((T - T0) - t) - nrm, r * t - nrm)
        dw[..., 0] = np.where(first, -r, r)
    else:
        raise ValueError(kind)
    dw[..., 1:] = -x / nrm[..., None]
    return w, dw


# ------------------------------------------------------------------------------------ weak form
def weak_form(thu, thv, X, XV, BX, coef, cfg, phase=None):
    """Everything src/training.py:129-137 / :153-161 produces for one batch.

    X, XV [N,L,C], BX [Nb,Lb,C] (time in channel 0)
    coef : dict(h[N], grad_h[N,d] (d s0/dx at X[:,0,1:]), f[N,L], g[Nb,Lb], sb[Nb] initial scalar
                of the boundary paths, optional s0[N] initial scalar of the interior paths (default h),
                a (None=identity | [d,d] constant), b (None | [d] constant), c0, c1 (c(X,u) = c0 + c1*u))
    cfg  : dict(nu, nv, solver, alpha, V, domain)
    returns dict with u, v, ub, w, du[N,d], dphi[N,L,C], I, S, init, bdry, loss_u, loss_v and,
    for phase in ('u','v'), `grads` = list in the reference's parameter order.
    """
    X = np.asarray(X, np.float64)
    XV = np.asarray(XV, np.float64)
    BX = np.asarray(BX, np.float64)
    N, L, C = X.shape
    d = C - 1
    nu, nv, solver = cfg["nu"], cfg["nv"], cfg.get("solver", "midpoint")
    alpha, V = float(cfg["alpha"]), float(cfg["V"])
    h, f, g_b = coef["h"], coef["f"], coef["g"]
    c0, c1 = float(coef.get("c0", 0.0)), float(coef.get("c1", 0.0))

    s0 = coef.get("s0", h)      # initial scalar of the interior paths: h(x) at T0, else g at the entry point (src/model.py:95-96)
    u, cu = xnode_forward(thu, X[:, 0, 1:], X[0, :, 0], s0, nu, solver)
    v, cv = vnet_forward(thv, XV, nv)
    w, dw = domain_w(cfg["domain"], XV)
    phi = v * w
    # du = grad_X sum(u): only time-row 0 is non-zero (src/model.py:99 reads x from row 0)
    _, gx, gs = xnode_vjp(thu, cu, np.ones_like(u), want_param_grads=False)
    du = gx + gs[:, None] * coef["grad_h"]
    _, dv = vnet_vjp(thv, cv, np.ones_like(v), nv, want_param_grads=False)
    dphi = w[..., None] * dv + v[..., None] * dw

    s1 = V / N * (u[:, -1] * v[:, -1] - h * v[:, 0])
    s2 = V / N / L * u * dphi[..., 0]
    a = coef.get("a")
    if a is None:
        s31_0 = (dphi[:, 0, 1:] * du).sum(-1)
    else:
        s31_0 = np.einsum("ij,ni,nj->n", np.asarray(a, np.float64), dphi[:, 0, 1:], du)
    s3f = (c0 + c1 * u) * u * phi + f * phi
    s3f[:, 0] += s31_0
    if coef.get("b") is not None:
        s3f[:, 0] += phi[:, 0] * (du @ np.asarray(coef["b"], np.float64))
    s3 = V / N / L * s3f
    I = float((s1 - (s2 - s3).sum(1)).sum())
    S = float(V * (v * v).sum() / (N * L))
    init = float(((u[:, 0] - h) ** 2).mean())
    ub, cb = xnode_forward(thu, BX[:, 0, 1:], BX[0, :, 0], coef["sb"], nu, solver)
    bdry = float(((ub - g_b) ** 2).mean())
    integ = np.log(I * I) - np.log(S)
    out = dict(u=u, v=v, ub=ub, w=w, phi=phi, du=du, dphi=dphi, I=I, S=S, init=init, bdry=bdry,
               loss_u=float(integ + alpha * (init + bdry)), loss_v=float(-integ))
    Aq = (c0 + c1 * u) * u
    Ap = c0 + 2.0 * c1 * u
    if phase == "u":
        Gu = (2.0 / I) * (V / N / L) * Ap * phi + 1.0
        Gu[:, -1] += (2.0 / I) * (V / N) * v[:, -1]
        Gu[:, 0] += alpha * (2.0 / N) * (u[:, 0] - h)
        gi, _, _ = xnode_vjp(thu, cu, Gu)
        Gb = alpha * 2.0 * (ub - g_b) / ub.size
        gb, _, _ = xnode_vjp(thu, cb, Gb)
        out["grads"] = [gi[k] + gb[k] for k in U_KEYS]
        out["G"] = Gu
    elif phase == "v":
        Pv = (V / N / L) * w * (Aq + f)
        Pv[:, -1] += (V / N) * u[:, -1]
        Pv[:, 0] -= (V / N) * h
        Gv = -(2.0 / I) * Pv + (1.0 / S) * (V / N / L) * 2.0 * v + w
        gv, _ = vnet_vjp(thv, cv, Gv, nv)
        out["grads"] = [gv[k] for k in V_KEYS]
        out["G"] = Gv
    return out


# ------------------------------------------------------------------- helpers for tests / fixtures
def theta_from_state(u_params, v_params):
    """lists of arrays in the reference's named_parameters() order -> (theta_u, theta_v)"""
    thu = {k: np.asarray(p, np.float64) for k, p in zip(U_KEYS, u_params)}
    thv = {k: np.asarray(p, np.float64) for k, p in zip(V_KEYS, v_params)}
    return thu, thv


def xavier_theta(d, H=20, hh=10, Hv=50, seed=0):
    """xavier-uniform weights / zero biases (src/training.py:46-49) with numpy's RNG"""
    rng = np.random.default_rng(seed)

    def xav(o, i):
        b = np.sqrt(6.0 / (i + o))
        return rng.uniform(-b, b, size=(o, i))

    thu = dict(W0=xav(H, 1), b0=np.zeros(H), W1=xav(H, H), b1=np.zeros(H), W2=xav(H, H), b2=np.zeros(H),
               Wa=xav(hh, d + 1 + H), ba=np.zeros(hh), Ws=xav(hh, hh), bs=np.zeros(hh),
               Wf=xav(H, hh), bf=np.zeros(H), Wo=xav(1, H), bo=np.zeros(1))
    thv = dict(Wi=xav(Hv, d + 1), bi=np.zeros(Hv), Wh=xav(Hv, Hv), bh=np.zeros(Hv),
               Wz=xav(1, Hv), bz=np.zeros(1))
    return thu, thv
