"""three eager training iterations at the shipped configuration (d=5, N_r=N_b=4000): run under
`ncu --metrics gpu__time_duration.sum` for the launch list of one iteration"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xnode_wan_b200 as xw  # noqa: E402

prob = xw.problems.ex4_1()
params = xw.problems.cube_params(dim=5, iterations=40)
torch.manual_seed(0)
np.random.seed(0)
s = xw.NODE_WAN_solver(params, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g, "cuda:0", "./",
                       func_u_sol=prob.func_u_sol, p=2, log_json=False, use_cuda_graph=False)
dom = s.new_domain()
for _ in range(int(os.environ.get("ITERS", "3"))):
    pts = xw.Comb_loader(4000, 4000, dom, "cuda:0")
    s.train_iteration(dom, pts)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("iter"); torch.cuda.nvtx.range_pop()
print("ok")
