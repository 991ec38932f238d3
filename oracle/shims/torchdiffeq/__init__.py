"""TEST INFRASTRUCTURE ONLY -- stand-in for the third-party package `torchdiffeq==0.1.1`.

The reference pins torchdiffeq 0.1.1 (/root/reference/requirements.txt:7) and calls
`odeint(func=field, y0=h0, t=timesteps, method=self.solver)` at
/root/reference/src/model.py:104-106.  The package is not vendored in the reference and is not
installable here (no network), so its *published* fixed-grid algorithm is restated:

  * no `step_size` option => integration grid == the requested output times `t`;
  * `t` is cast to the dtype of `y0`; it must be strictly monotone;
  * per consecutive pair (t0, t1): dy = step_func(func, t0, t1 - t0, y0); y1 = y0 + dy;
  * outputs at grid points are y1 exactly (linear interpolation only happens off-grid);
  * euler:    dy = dt * f(t0, y0)
  * midpoint: y_mid = y0 + f(t0, y0) * dt / 2 ; dy = dt * f(t0 + dt / 2, y_mid)
  * rk4 (3/8 rule, as torchdiffeq's `rk4_alt_step_func`):
        k1 = f(t0, y0); k2 = f(t0 + dt/3, y0 + dt*k1/3);
        k3 = f(t0 + 2dt/3, y0 + dt*(k2 - k1/3)); k4 = f(t1, y0 + dt*(k1 - k2 + k3));
        dy = dt * (k1 + 3(k2 + k3) + k4) / 8
  * `len(t) == 1` returns `[y0]`.

PARITY UNPINNED at this boundary: the reference ships no test that pins torchdiffeq's output.  It
is pinned here only by closed-form known answers (tests/test_oracle.py::test_midpoint_*).
"""
import torch

__version__ = "0.1.1-shim"


def _step(method, func, t0, dt, y):
    if method == "euler":
        return dt * func(t0, y)
    if method == "midpoint":
        y_mid = y + func(t0, y) * dt / 2
        return dt * func(t0 + dt / 2, y_mid)
    if method == "rk4":
        k1 = func(t0, y)
        k2 = func(t0 + dt / 3, y + dt * k1 / 3)
        k3 = func(t0 + dt * 2 / 3, y + dt * (k2 - k1 / 3))
        k4 = func(t0 + dt, y + dt * (k1 - k2 + k3))
        return dt * (k1 + 3 * (k2 + k3) + k4) / 8
    raise ValueError("fixed-grid shim supports euler/midpoint/rk4, got %r" % (method,))


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None):
    assert torch.is_floating_point(t), "t must be floating point"
    t = t.type_as(y0)
    if t.numel() > 1:
        d = t[1:] - t[:-1]
        assert bool((d > 0).all()) or bool((d < 0).all()), "t must be strictly increasing or decreasing"
    method = method or "dopri5"
    ys = [y0]
    y = y0
    for i in range(t.numel() - 1):
        t0, t1 = t[i], t[i + 1]
        y = y + _step(method, func, t0, t1 - t0, y)
        ys.append(y)
    return torch.stack(ys, 0)


odeint_adjoint = odeint
