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
    gu = [z["gu_%02d" % i] for i in range(14)]
    gv = [z["gv_%02d" % i] for i in range(6)]
    return dict(z=z, meta=meta, params=p, thu=thu, thv=thv, thu_list=thu_list, thv_list=thv_list,
                cfg=cfg, coef=coef, gu=gu, gv=gv)


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
