"""shipped configuration (d=5, N_r=N_b=4000): where a sub-iteration's time goes.
(1) pure sub-steps on a resident sample, eager and CUDA-graph replay; (2) cProfile of NODE_WAN_solver.train()"""
import cProfile
import io
import json
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xnode_wan_b200 as xw  # noqa: E402


def make(graph, **kw):
    prob = xw.problems.ex4_1()
    params = xw.problems.cube_params(dim=5, iterations=40)
    torch.manual_seed(0)
    np.random.seed(0)
    return xw.NODE_WAN_solver(params, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g, "cuda:0",
                              "./", func_u_sol=prob.func_u_sol, p=2, log_json=False, use_cuda_graph=graph, **kw), prob


out = {}
for graph in (False, True):
    s, prob = make(graph)
    dom = s.new_domain()
    pts = xw.Comb_loader(4000, 4000, dom, "cuda:0")
    for _ in range(5):
        s.train_iteration(dom, pts)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 100
    for _ in range(n):
        s.train_iteration(dom, pts)
    torch.cuda.synchronize()
    out["graph" if graph else "eager"] = {"ms_per_sub_iter": 1e3 * (time.perf_counter() - t0) / (3 * n)}
s, prob = make(True)
trace = []


def stop(sv, points, domain):
    r = xw.rel_err(points, sv.u_net, sv.func_u_sol, sv.p, domain.V(), sv.params['N_r']).item()
    trace.append(r)
    return False
s.stop = stop
s.keep_l2_history = False        # as bench.py's time_to_target: the per-iteration L2 on a fresh sample is logging only
pr = cProfile.Profile()
torch.cuda.synchronize()
t0 = time.perf_counter()
pr.enable()
s.train(report=False)
pr.disable()
torch.cuda.synchronize()
out["train_40_outer_iterations_s"] = time.perf_counter() - t0
out["train_ms_per_sub_iter"] = 1e3 * out["train_40_outer_iterations_s"] / 80
buf = io.StringIO()
pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(45)
print(buf.getvalue()[:6000], file=sys.stderr)
print(json.dumps(out))
