"""helpers shared by the CPU and GPU tests: load a golden fixture made by tests/golden/make_golden.py"""
import glob
import json
import os

import numpy as np

from oracle import closed_form as cf

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(degenerate=False):
    """regular cases; degenerate=True: the single-time-point interior groups (reference rank-2 shortcut)"""
    alln = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in alln if ("single_time" in n) == degenerate]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    p = meta["params"]
    thu_list = [z["thu_%02d" % i] for i in range(14)]
    thv_list = [z["thv_%02d" % i] for i in range(6)]
    thu, thv = cf.theta_from_state(thu_list, thv_list)
    cfg = dict(nu=p["u_layers"], nv=p["v_layers"], solver=p["solver"], alpha=p["alpha"], V=meta["V"],
               domain=tuple(meta["domain"]))
    coef = dict(h=z["h"].astype(np.float64), f=z["f"].astype(np.float64), g=z["g"].astype(np.float64),
                grad_h=z["grad_h"].astype(np.float64), sb=z["sb"].astype(np.float64),
                c0=meta["c0"], c1=meta["c1"])
    if "s0" in z.files:
        coef["s0"] = z["s0"].astype(np.float64)
    if "coef_a" in z.files:          # constant a != I the reference was run with (tests/golden/make_golden.py A_CONST)
        coef["a"] = z["coef_a"].astype(np.float64)
    gu = [z["gu_%02d" % i] for i in range(14)]
    gv = [z["gv_%02d" % i] for i in range(6)]
    return dict(z=z, meta=meta, params=p, thu=thu, thv=thv, thu_list=thu_list, thv_list=thv_list,
                cfg=cfg, coef=coef, gu=gu, gv=gv)


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# ---------------------------------------------------------------------------------------------
# full-size goldens (tests/golden/big, made by tests/golden/make_big_golden.py): the fixture holds seeds +
# the reference's outputs; inputs and weights are re-created with this package's bit-identical sampler /
# initialiser and verified against the stored SHA-1 digests.
# ---------------------------------------------------------------------------------------------
BIG_DIR = os.path.join(GOLDEN_DIR, "big")


class _Z(dict):
    @property
    def files(self):
        return list(self.keys())


def big_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(BIG_DIR, "*.npz")))


def load_big(name):
    import hashlib

    import torch

    import xnode_wan_b200 as xw

    def digest(a):
        return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()

    zz = np.load(os.path.join(BIG_DIR, name + ".npz"))
    meta = json.loads(bytes(zz["meta"]).decode())
    p = dict(meta["params"])
    p["domain"] = "Hypercube"
    seed = meta["seed"]
    rng_t, rng_n = torch.get_rng_state(), np.random.get_state()
    try:
        torch.manual_seed(seed)
        np.random.seed(seed)
        prob = xw.problems.by_name(meta["funcs"], p["dim"])
        s = xw.NODE_WAN_solver(p, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g, "cpu",
                               "./", func_u_sol=prob.func_u_sol, p=2, log_json=False)
        if meta["rand_bias"]:
            g = torch.Generator().manual_seed(seed + 100)
            with torch.no_grad():
                for q in list(s.u_net.parameters()) + list(s.v_net.parameters()):
                    q.add_(0.1 * torch.randn(q.shape, generator=g, dtype=q.dtype))
        dom = s.new_domain()
        pts = xw.Comb_loader(p["N_r"], p["N_b"], dom, "cpu")
        X, XV, BX = pts[0]
    finally:
        torch.set_rng_state(rng_t)
        np.random.set_state(rng_n)
    thu_list = [q.detach().numpy().copy() for q in s.u_net.parameters()]
    thv_list = [q.detach().numpy().copy() for q in s.v_net.parameters()]
    got = dict(X=digest(X.numpy()), XV=digest(XV.numpy()), BX=digest(BX.numpy()),
               thu=digest(np.concatenate([t.reshape(-1) for t in thu_list])),
               thv=digest(np.concatenate([t.reshape(-1) for t in thv_list])))
    assert got == meta["sha1"], "re-created inputs differ from the ones the reference ran on: %r" % (
        [k for k in got if got[k] != meta["sha1"][k]],)
    x0 = X[:, 0, :].clone().detach().double().requires_grad_(True)
    prob.func_h(x0).sum().backward()
    z = _Z(X=X.numpy(), XV=XV.numpy(), BX=BX.numpy(), h=prob.func_h(X[:, 0, :]).numpy(), f=prob.func_f(X).numpy(),
           g=prob.func_g(BX).numpy(), sb=prob.func_h(BX[:, 0, :]).numpy(), grad_h=x0.grad[:, 1:].numpy())
    for k in ("loss_u", "loss_v", "I", "S", "init", "bdry", "u_head", "v_head", "du_head"):
        z[k] = zz[k]
    thu, thv = cf.theta_from_state(thu_list, thv_list)
    cfg = dict(nu=p["u_layers"], nv=p["v_layers"], solver=p["solver"], alpha=p["alpha"], V=meta["V"],
               domain=tuple(meta["domain"]))
    coef = dict(h=z["h"].astype(np.float64), f=z["f"].astype(np.float64), g=z["g"].astype(np.float64),
                grad_h=z["grad_h"].astype(np.float64), sb=z["sb"].astype(np.float64), c0=meta["c0"], c1=meta["c1"])
    gu = [zz["gu_%02d" % i] for i in range(14)]
    gv = [zz["gv_%02d" % i] for i in range(6)]
    return dict(z=z, meta=meta, params={k: v for k, v in p.items() if k != "domain"}, thu=thu, thv=thv,
                thu_list=thu_list, thv_list=thv_list, cfg=cfg, coef=coef, gu=gu, gv=gv)
