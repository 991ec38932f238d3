"""error norms against a known solution (reference /root/reference/utils/auxillary_funcs.py:7-30)

The reference evaluates `stop()` -> `rel_err` after EVERY u sub-iteration on the training sample (src/training.py:142):
two `L_norm` calls, each moving the sample to the device and evaluating `func_u_sol` on it.  At the shipped size a
sub-iteration of this package is ~0.7 ms, so that host work decides the time to target.  Two things that do not depend on
the parameters are therefore kept with the sample: its device copy (the one `Comb_loader` already made, tagged with the
branch `NeuralODE.forward` takes, so no device->host probe of `X[0,0,0] == T0`) and the values of `func_u_sol` on it with
their norm.  Same numbers as before; `x._version` guards against a sample that was modified in place."""
import torch


def _version(x):
    if hasattr(x, "times"):                  # CollapsedPaths
        return (x.times._version, x.x._version)
    return x._version


def _on_device(x, dev):
    """the sample on `dev`: its loader's own device copy when there is one (Comb_loader._dev leaves it on the host tensor)"""
    kept = getattr(x, "_xw_dev", None)
    if kept is not None and kept[0] == _version(x) and kept[1].device == dev:
        return kept[1]
    return x.to(dev)


def _solution_on(x, xd, func_u_sol, p):
    """(func_u_sol on the sample, mean |func_u_sol|^p), evaluated once per sample and callable"""
    kept = getattr(x, "_xw_sol", None)
    if kept is not None and kept[0] is func_u_sol and kept[1] == _version(x) and kept[2] == p and kept[3].device == xd.device:
        return kept[3], kept[4]
    sol = func_u_sol(xd)
    mp = torch.mean(torch.abs(sol) ** p)
    try:
        x._xw_sol = (func_u_sol, _version(x), p, sol, mp)
    except AttributeError:
        pass
    return sol, mp


def L_norm(X, u_net, p: float, func_u_sol, volume: float, N_r: int, error=True):
    """(volume * mean |u_sol - u_net|^p)^(1/p) over the interior sample (error=False: norm of u_sol)"""
    dev = next(u_net.parameters()).device
    if isinstance(u_net, torch.nn.DataParallel) and len(u_net.device_ids) <= 1:
        # one device per process: DataParallel.forward would only walk all parameters to check their placement (~0.2 ms of
        # Python per call, as much as the evaluation kernel itself at the shipped size)
        u_net = u_net.module

    def mean_p(x):
        xd = _on_device(x, dev)
        sol, sol_mp = _solution_on(x, xd, func_u_sol, p)
        if not error:
            return sol_mp
        pred = u_net(xd)
        pred = pred.materialize() if hasattr(pred, "materialize") else pred
        return torch.mean(torch.abs(sol - pred.squeeze()) ** p)
    if not isinstance(X, list):
        return (volume * mean_p(X)) ** (1 / p)
    acc = 0
    for x in X:
        acc = acc + x.shape[0] / N_r * mean_p(x)
    return (volume * acc) ** (1 / p)


def rel_err(X, predu, func_u_sol, p: float, volume: float, N_r: int):
    return L_norm(X, predu, p, func_u_sol, volume, N_r) / L_norm(X, predu, p, func_u_sol, volume, N_r, error=False)
