"""where the wall time of a short training run at the shipped configuration goes (d=5, N_r=N_b=4000, CUDA-graph replay):
wall-clock accumulators around the pieces of NODE_WAN_solver.train(), with a device synchronisation at the end of each
piece so that GPU time is charged to the piece that launched it (the synchronisations themselves cost a few us each)."""
import collections
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xnode_wan_b200 as xw  # noqa: E402

ACC = collections.defaultdict(float)
CNT = collections.Counter()
SYNC = int(os.environ.get("PROBE_SYNC", "1"))


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        if SYNC:
            torch.cuda.synchronize()
        ACC[name] += time.perf_counter() - t0
        CNT[name] += 1
        return r
    return w


def run(iters, warm):
    prob = xw.problems.ex4_1()
    params = xw.problems.cube_params(dim=5, iterations=iters)
    torch.manual_seed(0)
    np.random.seed(0)
    s = xw.NODE_WAN_solver(params, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g, "cuda:0",
                           "./", func_u_sol=prob.func_u_sol, p=2, log_json=False, use_cuda_graph=True)
    s.keep_l2_history = False
    s.stop = lambda sv, points, domain: bool(xw.rel_err(points, sv.u_net, sv.func_u_sol, sv.p, domain.V(), sv.params['N_r']).item() < 0)
    if warm:
        s.iterations = 6
        s.train()
        s.iterations = iters
    s.sub_step = timed("sub_step", s.sub_step)
    s.stop = timed("stop", s.stop)
    s.new_domain = timed("new_domain", s.new_domain)
    real = xw.training.Comb_loader

    class Loader(real):
        pass
    Loader.__init__ = timed("Comb_loader", real.__init__)
    xw.training.Comb_loader = Loader
    ACC.clear(); CNT.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.train()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    xw.training.Comb_loader = real
    out = {"iterations": iters, "wall_s": dt, "ms_per_outer": 1e3 * dt / iters,
           "pieces_ms_per_outer": {k: 1e3 * v / iters for k, v in ACC.items()}, "calls": dict(CNT), "sync": SYNC}
    out["other_ms_per_outer"] = out["ms_per_outer"] - sum(out["pieces_ms_per_outer"].values())
    return out


if __name__ == "__main__":
    print(json.dumps({"warm": run(100, True)}))


def cold(n_solvers=6, iters=40):
    """per-outer-iteration wall time of fresh solvers (eager first iteration, graph captures, then replay)"""
    import gc
    out = []
    for seed in range(n_solvers):
        prob = xw.problems.ex4_1()
        params = xw.problems.cube_params(dim=5, iterations=iters)
        torch.manual_seed(seed)
        np.random.seed(seed)
        s = xw.NODE_WAN_solver(params, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g, "cuda:0",
                               "./", func_u_sol=prob.func_u_sol, p=2, log_json=False, use_cuda_graph=True)
        s.keep_l2_history = False
        marks = []
        real = s.new_domain

        def nd(**kw):
            marks.append(time.perf_counter())
            return real(**kw)
        s.new_domain = nd
        s.stop = lambda sv, points, domain: bool(xw.rel_err(points, sv.u_net, sv.func_u_sol, sv.p, domain.V(), sv.params['N_r']).item() < 0)
        gc.collect()
        if os.environ.get("PROBE_EMPTY"):
            torch.cuda.empty_cache()
        torch.cuda.synchronize()
        ev = []

        def wrap(obj, name, tag):
            fn = getattr(obj, name)

            def w(*a, **k):
                ta = time.perf_counter()
                r = fn(*a, **k)
                ev.append((tag, round(1e3 * (time.perf_counter() - ta), 2)))
                return r
            setattr(obj, name, w)
        wrap(s, "_capture", "_capture")
        wrap(s, "_graph_for", "_graph_for")
        wrap(s, "_step", "_step")
        G_ = torch.cuda.CUDAGraph
        b0, e0 = G_.capture_begin, G_.capture_end

        def cb(self_, *a, **k):
            ta = time.perf_counter(); r = b0(self_, *a, **k); ev.append(("begin", round(1e3 * (time.perf_counter() - ta), 2))); return r

        def ce(self_, *a, **k):
            ta = time.perf_counter(); r = e0(self_, *a, **k); ev.append(("end", round(1e3 * (time.perf_counter() - ta), 2))); return r
        G_.capture_begin, G_.capture_end = cb, ce
        gcev = []

        def gccb(phase, info):
            if phase == "start":
                gcev.append([info["generation"], time.perf_counter()])
            else:
                gcev[-1][1] = round(1e3 * (time.perf_counter() - gcev[-1][1]), 2)
                gcev[-1].append(info["collected"])
        gc.callbacks.append(gccb)
        ms0 = torch.cuda.memory_stats()
        t0 = time.perf_counter()
        s.train()
        ms1 = torch.cuda.memory_stats()
        gc.callbacks.remove(gccb)
        ev.append(("device_alloc", ms1["num_device_alloc"] - ms0["num_device_alloc"]))
        ev.append(("device_free", ms1["num_device_free"] - ms0["num_device_free"]))
        ev.append(("gc_over_1ms", [g for g in gcev if g[1] > 1.0]))
        G_.capture_begin, G_.capture_end = b0, e0
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        d = [round(1e3 * (b - a), 2) for a, b in zip([t0] + marks, marks + [t1])]
        out.append({"seed": seed, "wall_s": round(t1 - t0, 4), "between_domain_draws_ms": d[:8], "max_later_ms": max(d[8:]), "median_later_ms": sorted(d[8:])[len(d[8:]) // 2],
                    "events_over_10ms": [e for e in ev if isinstance(e[1], float) and e[1] > 10.0]})
        if os.environ.get("PROBE_KEEP"):
            KEEP.append(s)
        del s
    return out


KEEP = []


if __name__ == "__main__" and os.environ.get("PROBE_COLD"):
    print(json.dumps({"cold": cold()}))
