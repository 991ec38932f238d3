"""Regenerates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU
through oracle/ref_runner.py (shims: torchdiffeq restatement, empty matplotlib, numpy-sum proxy).

Run in the build container only:   python tests/golden/make_golden.py
The reference sets no seeds; every case here is seeded with torch.manual_seed / np.random.seed.
Each fixture holds the inputs (samples, weights, coefficient values the reference evaluated) and
the reference's outputs (loss_u, loss_v, I, S, init, bdry, u, v, du, dphi, parameter grads).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_runner as rr  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (params override, funcs module, seed, randomise biases?)
    "cube_d5_shipped_small": ({'N_r': 48, 'N_b': 40, 'dim': 5}, "Ex4_1_funcs", 0, False),
    "cube_d5_alpha1_randbias": ({'N_r': 64, 'N_b': 50, 'dim': 5, 'alpha': 1}, "Ex4_1_funcs", 1, True),
    "cube_d3_small_nets": ({'N_r': 32, 'N_b': 30, 'dim': 3, 'N_t': 7, 'u_layers': 3, 'v_layers': 2,
                            'u_hidden_dim': 8, 'u_hidden_hidden_dim': 6, 'v_hidden_dim': 12, 'alpha': 10},
                           "Ex4_1_funcs", 2, True),
    "cube_d4_ex43": ({'N_r': 40, 'N_b': 48, 'dim': 4, 'alpha': 100}, "Ex4_3_funcs", 3, True),
    "cube_d20_ex41": ({'N_r': 24, 'N_b': 40, 'dim': 20, 'alpha': 1000}, "Ex4_1_funcs", 4, True),
    "cube_d3_euler": ({'N_r': 16, 'N_b': 12, 'dim': 3, 'N_t': 5, 'solver': 'euler', 'alpha': 1}, "Ex4_1_funcs", 5, True),
    "cube_d3_rk4": ({'N_r': 16, 'N_b': 12, 'dim': 3, 'N_t': 5, 'solver': 'rk4', 'alpha': 1}, "Ex4_1_funcs", 6, True),
    "cube_d2_L2": ({'N_r': 8, 'N_b': 8, 'dim': 2, 'N_t': 2, 'alpha': 1}, "Ex4_1_funcs", 7, True),
    # constant a != I, b = 0 (no shipped config has it; the reference CAN run it: src/training.py:32-35, src/loss.py:66-68)
    "cube_d4_aconst": ({'N_r': 72, 'N_b': 40, 'dim': 4, 'alpha': 10}, "Ex4_3_funcs", 8, True),
}
A_CONST = {"cube_d4_aconst": 9}      # case -> seed of the constant matrix a = I + 0.3 * randn(d, d)
# a_ij(X) varying over the sample + c(X, u) non-affine (tests/_general_coef.py); stored under extra/ (the generic parity
# tests glob tests/golden/*.npz and assume structured coefficients)
GENERAL = {"general_coef_cube_d3": ({'N_r': 56, 'N_b': 40, 'dim': 3, 'alpha': 10}, "Ex4_3_funcs", 12, True)}


SPHERE_CASES = {
    # name: (domain, group index, seed)
    "cone_d5_g2": ("NSphere_TCone", 2, 3),
    "cone_d5_g4": ("NSphere_TCone", 4, 3),
    "hourglass_d5_g2_reentry": ("NSphere_THourglass", 2, 3),
    "hourglass_d5_g4_reentry": ("NSphere_THourglass", 4, 3),
    "hourglass_d5_g18": ("NSphere_THourglass", 18, 3),
    # group 0: single time point at T0 -> the reference's rank-2 shortcut and its [n, n] broadcasts
    "cone_d5_g0_single_time": ("NSphere_TCone", 0, 3),
}


def run_sphere_case(name, dom, group, seed):
    """variable-length float64 groups of the time-varying domains (src/dataset.py:48-229); group 0
    (single time point, rank-2 shortcut of src/model.py:89-91) is not a supported batch"""
    over = {'N_r': 300, 'N_b': 300, 'dim': 5, 'domain': dom, 'shape_param': 1.0, 'alpha': 10}
    solver, funcs, params = rr.build(over, "Ex4_3_funcs", seed)
    g_ = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        for p in list(solver.u_net.parameters()) + list(solver.v_net.parameters()):
            p.add_(0.1 * torch.randn(p.shape, generator=g_, dtype=p.dtype))
    domain, batches = rr.sample(solver)
    b = batches[group]
    ou = rr.evaluate(solver, domain, b, 'u')
    ov = rr.evaluate(solver, domain, b, 'v')
    c = rr.components(solver, domain, b)
    X, XV, BX = b
    T0 = params['T0']

    def scalar0(P):          # src/model.py:95-96
        x0 = P[:, 0, :].clone().detach().double().requires_grad_(True)
        val = funcs.func_h(x0) if float(P[0, 0, 0]) == T0 else funcs.func_g(x0.unsqueeze(1)).reshape(-1)
        gr, = torch.autograd.grad(val.sum(), x0)
        return val.detach().numpy(), gr[:, 1:].numpy()
    s0, gs0 = scalar0(X)
    sb, _ = scalar0(BX)
    domspec = ["cone", 1.0] if dom == "NSphere_TCone" else ["hourglass", 1.0, float(params['T0']), float(params['T'])]
    meta = dict(params={k: v for k, v in params.items() if k != 'domain'}, funcs="Ex4_3_funcs", seed=seed,
                domain=domspec, V=c['V'], c0=0.0, c1=-1.0, group=group, domain_class=dom)
    arrays = dict(X=X.numpy(), XV=XV.numpy(), BX=BX.numpy(), h=ou['h'], f=ou['f'], g=ou['g'], sb=sb, s0=s0, grad_h=gs0,
                  loss_u=np.float64(ou['loss']), loss_v=np.float64(ov['loss']), I=np.float64(c['I']), S=np.float64(c['S']),
                  init=np.float64(c['init']), bdry=np.float64(c['bdry']), u=ou['u'][..., 0], v=ou['v'][..., 0],
                  du=c['du'], dphi=c['dphi'], w=c['w'][..., 0], meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    for i, p in enumerate(solver.u_net.parameters()):
        arrays["thu_%02d" % i] = p.detach().numpy()
        arrays["gu_%02d" % i] = ou['grads'][i]
    for i, p in enumerate(solver.v_net.parameters()):
        arrays["thv_%02d" % i] = p.detach().numpy()
        arrays["gv_%02d" % i] = ov['grads'][i]
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-28s loss_u=%.10e loss_v=%.10e I=%.6e shapes %s %s (%d KB)" % (
        name, ou['loss'], ov['loss'], c['I'], tuple(X.shape), tuple(BX.shape), os.path.getsize(path) // 1024))


def grad_h(func_h, X0):
    x = X0.clone().detach().double().requires_grad_(True)
    func_h(x).sum().backward()
    return x.grad[:, 1:].numpy()


def run_case(name, over, funcs_name, seed, rand_bias):
    solver, funcs, params = rr.build(over, funcs_name, seed)
    if rand_bias:   # zero biases (xavier init) hide bias-path bugs: perturb every parameter a bit
        g = torch.Generator().manual_seed(seed + 100)
        with torch.no_grad():
            for p in list(solver.u_net.parameters()) + list(solver.v_net.parameters()):
                p.add_(0.1 * torch.randn(p.shape, generator=g, dtype=p.dtype))
    coef_a = None
    if name in A_CONST:
        d_ = params['dim']
        coef_a = np.eye(d_) + 0.3 * np.random.default_rng(A_CONST[name]).standard_normal((d_, d_))
        solver.func_a = lambda X_, i, j: torch.full(X_.shape[:-1], float(np.float32(coef_a[i, j])))
    if name in GENERAL:
        from tests import _general_coef as GC
        solver.func_a, solver.func_c = GC.func_a, GC.func_c
    domain, batches = rr.sample(solver)
    assert len(batches) == 1
    b = batches[0]
    ou = rr.evaluate(solver, domain, b, 'u')
    ov = rr.evaluate(solver, domain, b, 'v')
    c = rr.components(solver, domain, b)
    X, XV, BX = b
    sp = params['shape_param']
    meta = dict(params={k: (v if not callable(v) else str(v)) for k, v in params.items() if k != 'domain'},
                funcs=funcs_name, seed=seed, domain=["cube", float(sp[0]), float(sp[1])], V=c['V'],
                c0=0.0, c1=-1.0, general_coef=name in GENERAL)
    arrays = dict(
        X=X.numpy(), XV=XV.numpy(), BX=BX.numpy(),
        h=ou['h'], f=ou['f'], g=ou['g'],
        sb=funcs.func_h(BX[:, 0, :]).numpy(),              # src/model.py:95 (boundary starts at T0)
        grad_h=grad_h(funcs.func_h, X[:, 0, :]),
        loss_u=np.float64(ou['loss']), loss_v=np.float64(ov['loss']),
        I=np.float64(c['I']), S=np.float64(c['S']), init=np.float64(c['init']), bdry=np.float64(c['bdry']),
        u=ou['u'][..., 0], v=ou['v'][..., 0], du=c['du'], dphi=c['dphi'], w=c['w'][..., 0],
        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
    )
    if coef_a is not None:
        arrays["coef_a"] = coef_a.astype(np.float32).astype(np.float64)
    for i, p in enumerate(solver.u_net.parameters()):
        arrays["thu_%02d" % i] = p.detach().numpy()
        arrays["gu_%02d" % i] = ou['grads'][i]
    for i, p in enumerate(solver.v_net.parameters()):
        arrays["thv_%02d" % i] = p.detach().numpy()
        arrays["gv_%02d" % i] = ov['grads'][i]
    path = os.path.join(OUT, "extra" if name in GENERAL else "", name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-28s loss_u=%.10e loss_v=%.10e I=%.6e  (%d KB)" % (
        name, ou['loss'], ov['loss'], c['I'], os.path.getsize(path) // 1024))


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        run_case(name, *spec)
    for name, spec in SPHERE_CASES.items():
        if only and name not in only:
            continue
        run_sphere_case(name, *spec)
    for name, spec in GENERAL.items():
        if only and name not in only:
            continue
        run_case(name, *spec)
