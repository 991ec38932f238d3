"""end-to-end training of the shipped example (configs/cube_pde.yaml + Ex4_1) through the public API:
time / sub-iterations until the reference's own stop criterion (rel-L2 < 0.01,
/root/reference/configs/Ex4_1_funcs.py:36-37) fires.
usage: python tools/train_ex4_1.py [max_outer_iters] [seed] [dim]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xnode_wan_b200 as xw  # noqa: E402


def main():
    max_it = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dim = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    graph = (sys.argv[4] != "eager") if len(sys.argv) > 4 else True
    ondev = (sys.argv[5] == "device") if len(sys.argv) > 5 else False
    dev = "cuda:0"
    prob = xw.problems.ex4_1()
    params = xw.problems.cube_params(dim=dim, iterations=max_it)
    torch.manual_seed(seed)
    import numpy as np
    np.random.seed(seed)
    solver = xw.NODE_WAN_solver(params, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g,
                                dev, "./", func_u_sol=prob.func_u_sol, p=2, log_json=False, use_cuda_graph=graph,
                                sample_on_device=ondev, collapsed_layout=ondev)
    # rel-L2 trace on a fixed evaluation sample + the reference's stop criterion on the training sample
    trace = []
    t0 = time.time()

    def stop(s, points, domain):
        r = xw.rel_err(points, s.u_net, s.func_u_sol, s.p, domain.V(), s.params['N_r']).item()
        trace.append((len(trace) + 1, time.time() - t0, r))
        return r < 0.01
    solver.stop = stop
    hist = solver.train(report=False)
    torch.cuda.synchronize()
    wall = time.time() - t0
    miles = {}
    for thr in (0.10, 0.05, 0.03, 0.02, 0.015, 0.01):
        hit = next((t for t in trace if t[2] < thr), None)
        miles[str(thr)] = {"sub_iter": hit[0], "seconds": round(hit[1], 3)} if hit else None
    out = {"config": "cube_pde.yaml + Ex4_1, d=%d, N_r=N_b=4000, N_t=20, n1=2, n2=1" % dim, "seed": seed, "cuda_graph": graph, "sampling": "device, collapsed layout" if ondev else "cpu (reference RNG stream), [N,L,C] layout",
           "stopped_at_subiter": hist.get("stopped_at_subiter"), "sub_iters_run": len(trace), "wall_s": round(wall, 2),
           "final_rel_l2": trace[-1][2] if trace else None, "min_rel_l2": min(t[2] for t in trace) if trace else None,
           "milestones": miles, "ms_per_sub_iter": round(1e3 * wall / max(1, len(trace)), 3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
