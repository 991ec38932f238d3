"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference (/root/reference) on CPU to produce
golden vectors.  Works only where /root/reference exists (the build container); never imported by
the product, by the `-m gpu` tests, by smoke() or by bench.py.

Recipe = SURVEY.md Appendix A:
  * PYTHONPATH gets oracle/shims (torchdiffeq / matplotlib / NODE_GAN stand-ins) + /root/reference;
  * `np` inside module `src.loss` is replaced by a proxy whose sum(list_of_tensors) is an
    elementwise torch sum (/root/reference/src/loss.py:69 crashes on numpy>=1.20 otherwise);
  * the solver is built the way example.ipynb does, from the yaml dict with overrides in place
    (keeps the positional key order /root/reference/src/training.py:80-83 depends on);
  * every evaluation uses FRESH LEAF copies of (X, XV, BX).
"""
import importlib
import os
import sys

import numpy as _np
import torch

REF = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))


class _NpProxy:
    def __getattr__(self, k):
        return getattr(_np, k)

    @staticmethod
    def sum(x, *a, **k):
        if isinstance(x, (list, tuple)) and len(x) and torch.is_tensor(x[0]):
            return torch.stack(list(x), 0).sum(0)
        return _np.sum(x, *a, **k)


def available():
    return os.path.isdir(os.path.join(REF, "src"))


def load_reference(dim_for_ex43=None):
    """import the reference package with the shims in place; returns the module namespace dict"""
    assert available(), "reference tree not present on this machine"
    shim = os.path.join(_HERE, "shims")
    for p in (REF, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    import NODE_GAN.main as ng
    if dim_for_ex43 is not None:
        ng.params['dim'] = dim_for_ex43
    import src  # noqa: F401  (star-imports; `src.loss` attribute becomes the class)
    sys.modules['src.loss'].np = _NpProxy()
    training = sys.modules['src.training']
    dataset = sys.modules['src.dataset']
    return {"training": training, "dataset": dataset, "loss_mod": sys.modules['src.loss'],
            "model": sys.modules['src.model'],
            "aux": importlib.import_module("utils.auxillary_funcs")}


def load_funcs(name, dim=5):
    load_reference(dim)
    import NODE_GAN.main as ng
    ng.params['dim'] = dim
    mod = importlib.import_module("configs." + name)
    return mod


def base_params():
    import yaml
    with open(os.path.join(REF, "configs", "cube_pde.yaml")) as f:
        return yaml.safe_load(f)


def build(params_override, funcs_name="Ex4_1_funcs", seed=0):
    """-> (solver, funcs module, params)"""
    ref = load_reference(params_override.get('dim', 5))
    funcs = load_funcs(funcs_name, params_override.get('dim', 5))
    params = base_params()
    for k, v in params_override.items():
        assert k in params, k
        params[k] = v
    torch.manual_seed(seed)
    _np.random.seed(seed)
    solver = ref["training"].NODE_WAN_solver(
        params, funcs.func_a, funcs.func_b, funcs.func_c, funcs.func_h, funcs.func_f, funcs.func_g,
        'cpu', './', func_u_sol=funcs.func_u_sol, p=2)
    return solver, funcs, params


def sample(solver):
    """-> (domain, list of (X, XV, BX) batches) exactly as Comb_loader iterates them"""
    ref = load_reference()
    s = solver.setup
    domain = solver.domain(s['shape_param'], s['dim'], s['T0'], s['T'], s['N_t'])
    points = ref["dataset"].Comb_loader(s['N_r'], s['N_b'], domain, 'cpu')
    batches = [tuple(t.clone().detach() for t in b) for b in points]
    return domain, batches


def evaluate(solver, domain, batch, phase):
    """one fresh-leaf evaluation of /root/reference/src/training.py:129-137 (phase 'u') or
    :153-161 (phase 'v').  Returns dict of numpy arrays / floats."""
    ref = load_reference()
    tr, Loss = ref["training"], ref["loss_mod"].loss
    X, XV, BX = [t.clone().detach().requires_grad_(True) for t in batch]
    solver.optimizer_u.zero_grad()
    solver.optimizer_v.zero_grad()
    for p in list(solver.u_net.parameters()) + list(solver.v_net.parameters()):
        p.grad = None
    pv = solver.v_net(XV)
    pu = solver.u_net(X)
    h, f, g, a, b, c = tr.func_eval(X.clone().detach(), BX.clone().detach(), solver.setup, pu,
                                    solver.func_a, solver.func_b, solver.func_c, solver.func_h,
                                    solver.func_f, solver.func_g)
    L = Loss(solver.config['alpha'], a, b, c, h, f, g, solver.setup, domain, 'cpu')
    out = {}
    if phase == 'u':
        l = L.u(pu, pv, solver.u_net, X, XV, BX)
        l.backward(retain_graph=True)
        out['grads'] = [(p.grad if p.grad is not None else torch.zeros_like(p)).detach().clone().numpy()
                        for p in solver.u_net.parameters()]
    else:
        l = L.v(pu, pv, X, XV)
        l.backward(retain_graph=True)
        out['grads'] = [(p.grad if p.grad is not None else torch.zeros_like(p)).detach().clone().numpy()
                        for p in solver.v_net.parameters()]
    out['loss'] = float(l.item())
    out['u'] = pu.detach().numpy().copy()
    out['v'] = pv.detach().numpy().copy()
    out['h'] = h.detach().numpy().copy()
    out['f'] = f.detach().numpy().copy()
    out['g'] = g.detach().numpy().copy()
    return out


def components(solver, domain, batch):
    """I, S, init, bdry, du, dphi on a second fresh-leaf evaluation"""
    ref = load_reference()
    tr, Loss = ref["training"], ref["loss_mod"].loss
    X, XV, BX = [t.clone().detach().requires_grad_(True) for t in batch]
    pv = solver.v_net(XV)
    pu = solver.u_net(X)
    h, f, g, a, b, c = tr.func_eval(X.clone().detach(), BX.clone().detach(), solver.setup, pu,
                                    solver.func_a, solver.func_b, solver.func_c, solver.func_h,
                                    solver.func_f, solver.func_g)
    L = Loss(solver.config['alpha'], a, b, c, h, f, g, solver.setup, domain, 'cpu')
    # capture du / dphi before loss.I zeroes them: replicate the two helper backward calls
    pu.backward(torch.ones_like(pu), retain_graph=True)
    du = X.grad.detach().clone().numpy()
    X.grad.data.zero_()
    w = domain.func_w(XV).unsqueeze(2)
    phi = pv * w
    phi.backward(torch.ones_like(phi), retain_graph=True)
    dphi = XV.grad.detach().clone().numpy()
    XV.grad.data.zero_()
    I = L.I(pu, pv, X, XV)
    N = pv.shape[0] * pv.shape[1]
    S = L.V * torch.sum(pv ** 2) / N
    init = L.init(pu)
    bd = L.bdry(solver.u_net, BX)
    return {"I": float(I.item()), "S": float(S.item()), "init": float(init.item()),
            "bdry": float(bd.item()), "du": du, "dphi": dphi,
            "w": w.detach().numpy().copy(), "V": float(L.V)}
