"""bisect graph-replay vs eager differences: prints every sub-step loss of a few outer iterations for combinations of
(CUDA-graph replay, fused optimiser, coefficient cache)"""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xnode_wan_b200 as xw  # noqa: E402


def run(graph, fused, cache, iters=4, u_rate=None, v_nocache=False):
    torch.manual_seed(11)
    np.random.seed(11)
    p = xw.problems.cube_params(dim=5, N_r=4000, N_b=4000)
    if u_rate is not None:
        p['u_rate'] = u_rate
    prob = xw.problems.ex4_1()
    s = xw.NODE_WAN_solver(p, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g, "cuda:0", "./",
                           func_u_sol=prob.func_u_sol, p=2, log_json=False, use_cuda_graph=True, fused_optimizer=fused)
    s.use_cuda_graph = graph
    if v_nocache:
        orig2 = s._step
        s._step = lambda ph, d_, b_, vp=(None, 0), token=None: orig2(ph, d_, b_, vp, token=(None if ph == "v" else token))
    if not cache:
        orig = s._step
        s._step = lambda ph, d_, b_, vp=(None, 0), token=None: orig(ph, d_, b_, vp, token=None)
    out = []
    for it in range(iters):
        dom = s.new_domain()
        pts = xw.Comb_loader(4000, 4000, dom, "cuda:0")
        row = []
        for ph in ["u"] * s.n1 + ["v"] * s.n2:
            val = s.sub_step(ph, dom, pts)
            comp = val.components
            row.append((val.item(), comp["I"].item(), comp["S"].item(), comp["init"].item()))
        s._warm += 1
        out.append(row)
    return out


base = run(False, True, True, iters=10)
for rep in range(3):
    for graph in (False, True):
        r = run(graph, True, True, iters=10)
        print("rep", rep, "graph=%d vs eager" % graph, " | ".join(" ".join("%.0e" % (abs(a[0] - b[0]) / max(abs(b[0]), 1.0)) for a, b in zip(x, y)) for x, y in zip(r, base)), flush=True)
