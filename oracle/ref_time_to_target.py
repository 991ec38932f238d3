"""TEST INFRASTRUCTURE ONLY (build container: needs /root/reference) -- runs the UNMODIFIED reference training loop
(/root/reference/src/training.py:109-187) on CPU for the shipped config (configs/cube_pde.yaml + configs/Ex4_1_funcs.py,
d=5, N_r=N_b=4000) until its own stop criterion fires (rel-L2 < 0.01, configs/Ex4_1_funcs.py:36-37), and records the
number of u sub-iterations and the rel-L2 trace.  One seed per process:

    python oracle/ref_time_to_target.py SEED [max_outer] > profiles/r02_ref_time_to_target_seed<SEED>.json

bench.py prints these committed counts next to its own multi-seed runs (the Python reference cannot travel to the
GPU box).  Note: on CPU the reference re-uses the same leaf tensors across sub-iterations, so every 2nd u-pass and the
v-pass see stale X.grad contamination; that is the reference's behaviour on CPU and is kept.
"""
import json
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_runner as rr  # noqa: E402


def main():
    seed = int(sys.argv[1])
    max_outer = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    threads = int(os.environ.get("REF_THREADS", "2"))
    torch.set_num_threads(threads)
    ref = rr.load_reference(5)
    funcs = rr.load_funcs("Ex4_1_funcs", 5)
    params = rr.base_params()
    params['iterations'] = max_outer
    torch.manual_seed(seed)
    import numpy as np
    np.random.seed(seed)
    trace = []
    t0 = time.time()

    def stop(self, points, domain):
        r = float(ref["aux"].rel_err(points, self.u_net, self.func_u_sol, self.p, domain.V(), self.params['N_r']))
        trace.append(r)
        return r < 0.01
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)                      # the reference writes its json logs / checkpoints into the CWD
    stopped = False
    try:
        solver = ref["training"].NODE_WAN_solver(params, funcs.func_a, funcs.func_b, funcs.func_c, funcs.func_h, funcs.func_f,
                                                 funcs.func_g, 'cpu', tmp + "/", stop=stop, func_u_sol=funcs.func_u_sol, p=2)
        try:
            solver.train(report=False)
        except SystemExit:
            stopped = True
    finally:
        os.chdir(cwd)
    miles = {}
    for thr in (0.10, 0.05, 0.03, 0.02, 0.015, 0.01):
        hit = next((i + 1 for i, r in enumerate(trace) if r < thr), None)
        miles[str(thr)] = hit
    print(json.dumps({"impl": "reference (unmodified, CPU, fp64)", "config": "cube_pde.yaml + Ex4_1, d=5, N_r=N_b=4000, N_t=20, n1=2, n2=1",
                      "seed": seed, "stopped": stopped, "sub_iters": len(trace), "final_rel_l2": trace[-1] if trace else None,
                      "milestones_sub_iter": miles, "wall_s": round(time.time() - t0, 1), "threads": threads,
                      "trace_every_10": [round(r, 5) for r in trace[::10]]}))


if __name__ == "__main__":
    main()
