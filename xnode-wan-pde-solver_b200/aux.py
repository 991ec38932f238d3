"""error norms against a known solution (reference /root/reference/utils/auxillary_funcs.py:7-30)"""
import torch


def L_norm(X, u_net, p: float, func_u_sol, volume: float, N_r: int, error=True):
    """(volume * mean |u_sol - u_net|^p)^(1/p) over the interior sample (error=False: norm of u_sol)"""
    def resid(x):
        dev = next(u_net.parameters()).device
        x = x.to(dev)
        sol = func_u_sol(x)
        if not error:
            return sol
        pred = u_net(x)
        pred = pred.materialize() if hasattr(pred, "materialize") else pred
        return sol - pred.squeeze()
    if not isinstance(X, list):
        return (volume * torch.mean(torch.abs(resid(X)) ** p)) ** (1 / p)
    acc = 0
    for x in X:
        acc = acc + x.shape[0] / N_r * torch.mean(torch.abs(resid(x)) ** p)
    return (volume * acc) ** (1 / p)


def rel_err(X, predu, func_u_sol, p: float, volume: float, N_r: int):
    return L_norm(X, predu, p, func_u_sol, volume, N_r) / L_norm(X, predu, p, func_u_sol, volume, N_r, error=False)
