#!/usr/bin/env python
"""bench.py -- weak-loss+grad path-points/s of the XNODE-WAN hot path on B200.

One "step" = the hot part of one reference training iteration on ONE sample
(/root/reference/src/training.py:125-162): n1=2 u-steps (loss_u + theta_u gradients + Adam) then
n2=1 v-step (loss_v + theta_v gradients + Adam), coefficient evaluation included.  A path-point is
one (path, time-index) sample; a u-step covers (N_r + N_b) * N_t of them, a v-step N_r * N_t.

  python bench.py --gpus N --steps K --warmup W             # this framework (sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle port)
  python bench.py --config xl ...                            # BASELINE configs[4]: d=100, 2^22 paths

Workload (BASELINE.json configs[3]): cube PDE (Ex4_1), d=20, N_t=20, N_r = N_b = 2^20 paths PER
RANK (weak scaling, `value`), alpha=1e8, shipped network sizes, synthetic seeded samples, xavier
weights.  In the same run the STRONG split of configs[3] is timed too (2^20 paths in total, 2^20/N per
rank, `strong`), at N=1 the shipped config is trained to the reference's stop criterion on several seeds
(`time_to_target`), and at N>1 the sharded loss/gradients are checked against the unsharded ones on
rank 0 (`shard_check`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

L_T = 20
CONFIGS = {"m": dict(dim=20, log2n=20), "xl": dict(dim=100, log2n=22)}


def alg_flops(d, H=20, hh=10, nu=8, Hv=50, nv=9, L=L_T):
    """algorithmic FLOPs per path-point (un-hoisted dense counts, MAC = 2 FLOP)"""
    F = (H + d + 1) * hh + (nu - 1) * hh * hh + hh * H
    U = 2 * (L - 1) / L * F + H + (H + 2 * H * H) / L
    Vm = (d + 1) * Hv + nv * Hv * Hv + Hv
    # what the generation-2 XNODE kernels actually execute per field evaluation (reduced state z = Wy y):
    Fr = (nu - 1) * hh * hh + hh * hh + hh          # shared layers + M tau + v.tau
    Ur = 2 * (L - 1) / L * Fr
    return {"u_interior": 2 * (4 * U + 2 * Vm), "u_boundary": 2 * 3 * U, "v_interior": 2 * (2 * U + 4 * Vm),
            "U": U, "Vm": Vm, "U_reduced_hw": Ur,
            # per C-ABI call (what one launch group computes), per point of its own batch
            "xw_interior_forward": 2 * (2 * U + 2 * Vm),          # XNODE forward + du sweep, v net value + tangent
            "xw_interior_forward:cached_v": 2 * (2 * U),          # test-function values come from the cache
            "xw_boundary_u": 2 * 3 * U,
            "xw_interior_backward_u": 2 * 2 * U, "xw_interior_backward_v": 2 * 2 * Vm}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def cpu_step_time(d, n, n_b, steps, warmup, threads):
    """the reference's CPU path (oracle/torch_port.py: same tensor ops as the reference, fp64,
    autograd, dense a[d,d,N,L]) on a bounded sample: one step = 2 u-steps + 1 v-step"""
    import xnode_wan_b200 as xw
    from oracle import torch_port as tp
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    prob = xw.problems.ex4_1()
    pu, pv = tp.make_params(d, seed=0)
    dom = xw.Hypercube((-1.0, 1.0), d, 0, 1, L_T)
    pts = xw.Comb_loader(n, n_b, dom, "cpu")
    X, XV, BX = pts[0]
    cfg = dict(nu=8, nv=9, solver="midpoint", alpha=1e8, bot=-1.0, top=1.0, V=dom.V())
    ou = torch.optim.Adam(pu, lr=0.015)
    ov = torch.optim.Adam(pv, lr=0.04)

    def one():
        for _ in range(2):
            tp.step("u", pu, pv, X, XV, BX, prob, cfg)
            ou.step()
        tp.step("v", pu, pv, X, XV, BX, prob, cfg)
        ov.step()
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    pp = (2 * (n + n_b) + n) * L_T
    return dt, pp


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.cpu_paths
    dt, pp = cpu_step_time(args.dim, n, n, args.steps, args.warmup, threads)
    val = pp / dt
    sample = "d=%d, N_r=N_b=%d paths (of 2^%d), N_t=%d, fp64, %d threads" % (args.dim, n, args.log2n, L_T, threads)
    line = {"impl": "reference", "metric": "weak-loss+grad path-points/sec", "value": val, "unit": "path-points/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args),
            "cpu_baseline": {"value": val, "unit": "path-points/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "path-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference CPU path = oracle/torch_port.py (the reference itself cannot travel to the GPU box; "
                    "the port is pinned to it by golden vectors); the reference cannot run the full 2^%d-path "
                    "config (dense a[d,d,N,L]), so a bounded sample is timed and quoted per path-point" % args.log2n}
    print(json.dumps(line), flush=True)


def config_of(args):
    return {"workload": "cube PDE Ex4_1 d=%d, N_r=N_b=2^%d paths per rank, N_t=%d, alpha=1e8, H=20 hh=10 nu=8 Hv=50 nv=9, "
                        "midpoint; step = 2 u-steps + 1 v-step on one sample" % (args.dim, args.log2n, L_T),
            "name": args.config, "dim": args.dim, "N_r_per_rank": 1 << args.log2n, "N_b_per_rank": 1 << args.log2n, "N_t": L_T,
            "layout": "collapsed (times[L] + x[N,d]); kernels also take the reference [N,L,C] layout (e2e_nlc)",
            "l2": "inputs_exceed_L2 (3 x %.0f MB of coordinates + 2 x %.0f MB of per-point seeds per step vs 126 MB L2)"
                  % ((1 << args.log2n) * args.dim * 4 / 1e6, (1 << args.log2n) * L_T * 4 / 1e6),
            "arithmetic": "fp32 results: XNODE kernels FP32 FFMA2 (reduced state z = Wy y, shared layer in registers); "
                          "test-function net on tcgen05 kind::tf32 with 3xTF32 error compensation (1e-6 vs fp64; "
                          "XW_VNET_IMPL=tile selects the pure-FP32 kernels)",
            "parallelism": "dp%d (paths sharded, 2 small all-reduces per sub-step)" % args.gpus,
            "solver_options": {"use_cuda_graph": bool(getattr(args, "graph", 0)), "fused_optimizer": bool(getattr(args, "fused", 1))}}


XNODE_NAMES = {1: ("k_xnode_fwd", "k_xnode_bwd"), 2: ("k_xnode2_fwd", "k_xnode2_bwd (+ k_xnode2_lift, k_xnode2_finish)"),
               3: ("k_xnode2_fwd", "k_xnode3_bwd (+ k_xnode2_fwd<history> in front of the boundary pass, k_xnode2_lift, k_xnode2_finish)")}
VNET_FWD = {1: "k_vnet_points", 2: "k_vnet_tile_fwd", 3: "k_vnet_tc_fwd (+ k_vnet_tc_row0)"}
VNET_BWD = {1: "k_vnet_bwd", 2: "k_vnet_tile_bwd", 3: "k_vnet_tc_bwd3",
            4: "k_vnet_tc_bwd3 on the virtual net of input width Hv (+ k_vv_prep, k_vv_dwx, k_vv_finish)"}


def make_solver(xw, d, n_glob, dev, **kw):
    prob = xw.problems.ex4_1()
    params = xw.problems.cube_params(dim=d, N_r=n_glob, N_b=n_glob, N_t=L_T, shape_param=[-1.0, 1.0])
    torch.manual_seed(0)
    return xw.NODE_WAN_solver(params, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g,
                              dev, "./", func_u_sol=prob.func_u_sol, p=2, log_json=False, **kw), prob


def time_to_target(xw, dev, seeds, max_outer=600):
    """shipped config (configs/cube_pde.yaml + Ex4_1: d=5, N_r=N_b=4000, n1=2, n2=1) trained through the public API
    until the reference's own stop criterion fires (rel-L2 < 0.01 on the training sample,
    /root/reference/configs/Ex4_1_funcs.py:36-37, evaluated after every u sub-iteration as src/training.py:142)"""
    import numpy as np
    out = {"config": "cube_pde.yaml + Ex4_1, d=5, N_r=N_b=4000, N_t=20, n1=2, n2=1; CPU sampling with the reference's RNG stream; "
                     "CUDA-graph replay of the sub-steps", "criterion": "rel-L2 < 0.01 (reference stop())",
           "seeds": [], "sub_iters": [], "seconds": [], "final_rel_l2": [], "ms_per_sub_iter": [],
           "steady_ms_per_outer_iter": [],
           "note": "seconds = wall time of train() of a FRESH solver: one eager iteration and four CUDA-graph captures (0.1-0.4 s, "
                   "varies from solver to solver) come first; steady_ms_per_outer_iter = median time of a later outer iteration "
                   "(2 u sub-iterations with stop() after each + 1 v sub-iteration + the next sample drawn on the host)"}
    for seed in seeds:
        prob = xw.problems.ex4_1()
        params = xw.problems.cube_params(dim=5, iterations=max_outer)
        torch.manual_seed(seed)
        np.random.seed(seed)
        trace = []
        solver = xw.NODE_WAN_solver(params, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g,
                                    dev, "./", func_u_sol=prob.func_u_sol, p=2, log_json=False, use_cuda_graph=True)

        def stop(s, points, domain):
            r = xw.rel_err(points, s.u_net, s.func_u_sol, s.p, domain.V(), s.params['N_r']).item()
            trace.append(r)
            return r < 0.01
        solver.stop = stop
        solver.keep_l2_history = False           # (the per-iteration L2 on a fresh sample is logging only; stop() is what counts)
        import gc
        gc.collect()                             # (the previous seed's solver -- CUDA graphs, their memory pools -- is released here,
        torch.cuda.synchronize()                 #  not by a collection that happens to run inside the timed region)
        t0 = time.perf_counter()
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):       # train() prints 'Stopping Criterion Reached' as the reference does
            hist = solver.train(report=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["seeds"].append(seed)
        out["sub_iters"].append(hist.get("stopped_at_subiter"))
        out["seconds"].append(round(dt, 3))
        out["final_rel_l2"].append(trace[-1] if trace else None)
        out["ms_per_sub_iter"].append(round(1e3 * dt / max(1, len(trace)), 3))
        tt = hist.get("time", [])
        gaps = sorted(b - a for a, b in zip(tt[5:], tt[6:]))          # outer iterations after the eager one and the graph captures
        out["steady_ms_per_outer_iter"].append(round(1e3 * gaps[len(gaps) // 2], 3) if gaps else None)
    ref = []
    try:      # the UNMODIFIED reference on the same seeds (CPU, build container; oracle/ref_time_to_target.py)
        ref = json.load(open(os.path.join(ROOT, "profiles", "r02_ref_time_to_target.json")))["runs"]
    except Exception:
        pass
    out["reference_cpu"] = {"source": "profiles/r02_ref_time_to_target.json (unmodified reference, CPU fp64, build container)",
                            "seeds": [r["seed"] for r in ref], "sub_iters": [r["sub_iters"] if r["stopped"] else None for r in ref],
                            "final_rel_l2": [r["final_rel_l2"] for r in ref], "seconds": [r["wall_s"] for r in ref]}
    return out


def shard_check(xw, dev, world, rank, d, log2n=14):
    """sharded == unsharded on the real path: every rank evaluates loss_u / loss_v and their gradients on its shard of
    one sample (all-reduced sums and gradients); rank 0 then evaluates the gathered sample alone."""
    import numpy as np
    n_loc = 1 << log2n
    solver, _ = make_solver(xw, d, n_loc * world, dev)
    torch.manual_seed(777 + rank)
    dom = solver.new_domain(sample_device=dev, collapsed=True)
    pts = xw.Comb_loader(n_loc, n_loc, dom, dev)
    g0 = torch.distributed.new_group(ranks=[0])      # a one-rank group: rank 0's unsharded evaluation reduces over itself only

    def evaluate(s, points, domain, group=None):
        res = {}
        for phase in ("u", "v"):
            s.optimizer_u.zero_grad(set_to_none=True)
            s.optimizer_v.zero_grad(set_to_none=True)
            X, XV, BX = points[0]
            pv, pu = s.v_net(XV), s.u_net(X)
            h, f, g, a, b, c = xw.func_eval(X, BX, s.setup, pu, s.func_a, s.func_b, s.func_c, s.func_h, s.func_f, s.func_g)
            Lo = xw.loss(s.config["alpha"], a, b, c, h, f, g, s.setup, domain, dev)
            if s.world > 1:
                Lo.N_glob, Lo.Nb_glob = X.shape[0] * s.world, BX.shape[0] * s.world
            if group is not None:
                Lo.group = group
            val = Lo.u(pu, pv, s.u_net, X, XV, BX) if phase == "u" else Lo.v(pu, pv, X, XV)
            val.backward()
            net = s.u_net if phase == "u" else s.v_net
            res[phase] = (val.detach().clone(), torch.cat([q.grad.reshape(-1) for q in net.parameters()]).clone())
        return res
    sharded = evaluate(solver, pts, dom)
    # gather the shards on every rank (rank 0 uses them)
    parts = []
    for t in (pts.interioru.x, pts.interiorv.x, pts.boundary.x):
        buf = [torch.empty_like(t) for _ in range(world)]
        torch.distributed.all_gather(buf, t.contiguous())
        parts.append(torch.cat(buf, 0))
    out = None
    if rank == 0:
        solver.world = 1                    # evaluate the gathered sample alone: no all-reduce, local == global counts
        full = xw.Comb_loader.from_tensors(*[xw.CollapsedPaths(dom.times.to(dev), p) for p in parts], dev)
        single = evaluate(solver, full, dom, g0)
        solver.world = world
        out = {"paths_per_rank": n_loc, "dim": d}
        for phase in ("u", "v"):
            lv, gv = sharded[phase]
            ls, gs = single[phase]
            out["loss_%s_rel" % phase] = abs(lv.item() - ls.item()) / max(abs(ls.item()), 1e-300)
            out["grad_%s_rel_l2" % phase] = (torch.linalg.norm(gv - gs) / torch.linalg.norm(gs)).item()
        out["ok"] = bool(out["loss_u_rel"] < 1e-6 and out["loss_v_rel"] < 1e-6 and out["grad_u_rel_l2"] < 1e-4 and out["grad_v_rel_l2"] < 1e-4)
    torch.distributed.barrier()
    return out


def run_ours(args):
    import xnode_wan_b200 as xw
    hp = xw.hotpath
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    lib = xw._lib.get()
    d = args.dim

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item() / steps

    def workload(n_loc):
        solver, _ = make_solver(xw, d, n_loc * world, dev, use_cuda_graph=bool(args.graph), fused_optimizer=bool(args.fused))
        torch.manual_seed(1000 + rank)
        domain = solver.new_domain(sample_device=dev, collapsed=True)
        points = xw.Comb_loader(n_loc, n_loc, domain, dev)
        points[0]
        return solver, domain, points

    # ------------------------------------------------------------------ weak workload (the headline `value`)
    n_loc = 1 << args.log2n
    solver, domain, points = workload(n_loc)
    # pinned host copy of the same sample for the end-to-end arm
    host = [t.to("cpu").pin_memory() for t in (points.interioru, points.interiorv, points.boundary)]
    pp_step = (2 * (n_loc + n_loc) + n_loc) * L_T * world

    def step_resident():
        solver.train_iteration(domain, points)

    last = {}

    nxt = {}

    def step_e2e():
        # input pipeline of depth 1 (Comb_loader.prefetch): the H2D copy of the NEXT step's sample runs on a copy stream
        # under this step's kernels; every step still moves its 252 MB inside the timed region and reads its losses back
        pts = nxt.get("pts") or xw.Comb_loader.from_tensors(host[0], host[1], host[2], dev).prefetch()
        nxt["pts"] = xw.Comb_loader.from_tensors(host[0], host[1], host[2], dev).prefetch()
        lu, lv = solver.train_iteration(domain, pts)
        last["lu"], last["lv"] = lu.item(), lv.item()

    # FP32-FMA peak of this GPU, measured in this run (roofline denominator)
    import ctypes as C
    fl = C.c_double(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        lib.call("xw_fma_probe", 0, 2048, C.byref(fl), st)
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    lib.call("xw_fma_probe", 0, 4096, C.byref(fl), st)
    eb.record()
    torch.cuda.synchronize()
    fma_peak = fl.value / (ea.elapsed_time(eb) * 1e-3) / 1e12

    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    hp.PROFILE = {}
    hp.CALLS.clear()
    hp.LAUNCHES[0] = 0
    ms = timed(step_resident, args.steps)
    prof = hp.PROFILE
    hp.PROFILE = None
    clk = clocks.stop() if rank == 0 else None
    xnode_code = lib.cdll.xw_last_xnode_impl()
    xnode_impl = (xnode_code >> 4) & 15 or (xnode_code & 15)       # name the step after its BACKWARD kernels
    vimpl = lib.cdll.xw_last_vnet_impl()
    # per-entry-point device time (CUDA events on the launching stream, inside the timed region)
    prof_steps = args.steps
    launches = hp.LAUNCHES[0] // args.steps
    if not prof:
        # CUDA-graph replay (--graph 1): the C-ABI calls ran at capture time only, there are no per-call events inside the
        # timed region.  The per-kernel breakdown then comes from two eager steps of the same solver AFTER the timed
        # region (`value` / `ms_per_step` stay the replayed ones); the launch count is the eager one as well.
        solver.use_cuda_graph = False
        hp.PROFILE = {}
        hp.LAUNCHES[0] = 0
        prof_steps = 2
        for _ in range(prof_steps):
            step_resident()
        torch.cuda.synchronize()
        prof = hp.PROFILE
        hp.PROFILE = None
        launches = hp.LAUNCHES[0] // prof_steps
        solver.use_cuda_graph = True
    per_call = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in prof.items()}
    per_step = {k: sum(a.elapsed_time(b) for a, b in v) / prof_steps for k, v in prof.items()}

    for _ in range(max(3, args.warmup)):        # (the end-to-end arm allocates per-step device buffers: let the allocator settle)
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    h2d = sum(t.nbytes() for t in host)
    d2h = 16
    nxt.clear()
    del host

    # e2e through the reference's own [N, L, C] layout (time in channel 0, x repeated along L): 3 x N*L*C*4 bytes per step
    e2e_nlc = None
    nlc_bytes = 3 * n_loc * L_T * (d + 1) * 4
    if nlc_bytes <= args.nlc_max_gb * 1e9:
        host_nlc = [t.dense().to("cpu").pin_memory() for t in (points.interioru, points.interiorv, points.boundary)]

        def step_nlc():
            pts = xw.Comb_loader.from_tensors(host_nlc[0], host_nlc[1], host_nlc[2], dev)
            lu, lv = solver.train_iteration(domain, pts)
            last["lu_nlc"] = lu.item() + 0 * lv.item()
        step_nlc()
        ms_nlc = timed(step_nlc, max(1, args.steps // 2))
        e2e_nlc = {"value": pp_step / (ms_nlc * 1e-3), "unit": "path-points/s", "ms_per_step": ms_nlc,
                   "h2d_bytes_per_step": sum(t.nbytes for t in host_nlc), "d2h_bytes_per_step": 16, "loss_u": last.get("lu_nlc"),
                   "layout": "reference [N,L,C] fp32 from pinned host memory (what /root/reference/src/dataset.py:321 moves)"}
        del host_nlc
    # ------------------------------------------------------------------ opt-in kernel variants on the same resident sample
    # XW_TC_SPLIT=1: a layer's tcgen05 MMAs of the forward kernel issued by three warps (faster; values reproducible to fp32
    # rounding instead of bit for bit: xw_capi.cu tc_split_issue).  The headline `value` is the DEFAULT configuration.
    variants = {}
    if world == 1:
        os.environ["XW_TC_SPLIT"] = "1"
        try:
            for _ in range(2):
                step_resident()
            ms_v = timed(step_resident, args.steps)
            variants["XW_TC_SPLIT=1"] = {"ms_per_step": ms_v, "value": pp_step / (ms_v * 1e-3), "unit": "path-points/s",
                                         "note": "forward tcgen05 kernel with three issuing warps per tile stream (not bit-reproducible from run to run)"}
        finally:
            os.environ.pop("XW_TC_SPLIT", None)
    # use_cuda_graph=True: the same step replayed from captured graphs (one launch per sub-step; the boundary pass and the
    # interior pass are parallel branches, hotpath.WeakLoss).  A second solver on the same resident sample; one GPU only
    # (the headline stays the eager path: its per-entry CUDA events are what `roofline` is computed from).
    if world == 1 and not args.graph and args.config == "m":
        solver_g = None
        try:
            solver_g, _ = make_solver(xw, d, n_loc * world, dev, use_cuda_graph=True, fused_optimizer=bool(args.fused))
            for _ in range(4):                   # eager iteration, captures, first replays
                solver_g.train_iteration(domain, points)
            ms_g = timed(lambda: solver_g.train_iteration(domain, points), args.steps)
            replayed = solver_g._graphs is not None and len(solver_g._graphs["graphs"]) > 0
            variants["use_cuda_graph=True"] = {"ms_per_step": ms_g, "value": pp_step / (ms_g * 1e-3), "unit": "path-points/s",
                                               "graphs_replayed": bool(replayed),
                                               "note": "NODE_WAN_solver(use_cuda_graph=True): sub-steps replayed from CUDA graphs"}
        except Exception as e:                   # (a variant must never take the headline line down with it)
            variants["use_cuda_graph=True"] = {"error": str(e)[:300]}
        finally:
            del solver_g
            torch.cuda.empty_cache()
    del solver, points
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ strong split of the same config (2^log2n paths in total)
    strong = None
    if world > 1 and (n_loc % world) == 0:
        n_s = n_loc // world
        solver_s, domain_s, points_s = workload(n_s)

        def step_strong():
            solver_s.train_iteration(domain_s, points_s)
        for _ in range(args.warmup):
            step_strong()
        ms_s = timed(step_strong, args.steps)
        strong = {"paths_total": n_loc, "paths_per_rank": n_s, "ms_per_step": ms_s, "value": (2 * (n_loc + n_loc) + n_loc) * L_T / (ms_s * 1e-3),
                  "unit": "path-points/s", "note": "BASELINE configs[3] as written: N_r = N_b = 2^%d paths in total, sharded over %d ranks" % (args.log2n, world)}
        del solver_s, points_s
        torch.cuda.empty_cache()
    elif world == 1:
        strong = {"paths_total": n_loc, "paths_per_rank": n_loc, "ms_per_step": ms, "value": pp_step / (ms * 1e-3), "unit": "path-points/s",
                  "note": "N=1: identical to the weak line"}

    check = shard_check(xw, dev, world, rank, d) if world > 1 else None

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    fl_ = alg_flops(d)
    pts_call = n_loc * L_T
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    ncalls = {k: len(v) / prof_steps for k, v in prof.items()}
    tj = {}
    try:       # DRAM bytes per point of each kernel from the committed `ncu --set full` captures (NOT measured in this run)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        pass
    nv_, kin = 9, (d + 2 + 7) // 8 * 8

    def mma_flops(n_cols):              # one tcgen05.mma kind::tf32, M = 128, K = 8
        return 2 * 128 * n_cols * 8
    # executed tensor FLOPs per point (3xTF32 = 3 MMAs per product, shapes padded to 128 x 56 x 56):
    # forward recompute + R-op + P-op (16 k-steps of 8 points) + input-layer gradient, 128 points per tile
    tc_bwd_exec = (3 * (kin // 8 + 7 * nv_) * mma_flops(56) + 3 * 7 * nv_ * mma_flops(56) + 3 * 16 * nv_ * mma_flops(56) +
                   3 * 16 * mma_flops(kin)) / 128.0
    tf32_peak = peaks.get("bf16_tflops", 1590.0) / 2.0
    xf, xb = XNODE_NAMES.get(xnode_impl, ("xnode_fwd?", "xnode_bwd?"))
    vf, vb = VNET_FWD.get(vimpl & 15, "vnet_fwd?"), VNET_BWD.get((vimpl >> 4) & 15, "vnet_bwd?")
    groups = {
        # kernel (named after what actually launched: xw_last_xnode_impl / xw_last_vnet_impl) -> (C-ABI entries, bound)
        xb: (["xw_boundary_u", "xw_interior_backward_u"], "fp32_fma"),
        vb: (["xw_interior_backward_v"], "tensor" if (vimpl >> 4) & 15 in (3, 4) else "fp32_fma"),
        "%s + k_weak_combine (interior forward, test-function values cached)" % xf: (["xw_interior_forward:cached_v"], "fp32_fma"),
        "%s + %s (interior forward, test-function net evaluated)" % (xf, vf): (["xw_interior_forward"], "mixed"),
    }
    kern = {}
    for name, (entries, bound) in groups.items():
        t_ms = sum(per_step.get(e, 0.0) for e in entries)
        if t_ms <= 0:
            continue
        alg = sum(fl_[e] * pts_call * ncalls.get(e, 0) for e in entries)
        launches_k = sum(ncalls.get(e, 0) for e in entries)
        r = {"bound": bound, "ms_per_step": t_ms, "calls_per_step": launches_k, "entries": entries,
             "algorithmic_tflops": alg / (t_ms * 1e-3) / 1e12}
        if bound == "fp32_fma":
            r.update(achieved=r["algorithmic_tflops"], peak=fma_peak, unit="TFLOP/s", frac=r["algorithmic_tflops"] / fma_peak,
                     peak_source="xw_fma_probe (FFMA chains) measured in this run; nominal 74.4",
                     note="achieved = ALGORITHMIC FLOPs (SURVEY 8d un-hoisted dense counts) / time; the generation-2 XNODE kernels "
                          "execute fewer (reduced state: %.0f instead of %.0f MAC per point and pass)" % (fl_["U_reduced_hw"], fl_["U"]))
        elif bound == "tensor":
            ex = tc_bwd_exec * pts_call * launches_k / (t_ms * 1e-3) / 1e12
            r.update(achieved=r["algorithmic_tflops"], peak=tf32_peak, unit="TFLOP/s", frac=r["algorithmic_tflops"] / tf32_peak,
                     executed_tflops=ex, frac_executed=ex / tf32_peak,
                     useful_3x_tflops=3 * r["algorithmic_tflops"], frac_useful_3x=3 * r["algorithmic_tflops"] / tf32_peak,
                     frac_of_fp32_fma_peak=r["algorithmic_tflops"] / fma_peak,
                     executed_flops_per_point=tc_bwd_exec,
                     peak_source="MEASURED_PEAKS.json bf16_tflops (burst) / 2: kind::tf32 runs at half the bf16 rate",
                     note="achieved / frac = ALGORITHMIC FLOPs (4 Vm per point); fp32-accurate results need 3 TF32 MMAs per product "
                          "(frac_useful_3x) and the tiles are padded 50->56 with a forward recompute (frac_executed); "
                          "frac_of_fp32_fma_peak = the same algorithmic rate against the FP32 pipe it replaces")
        kern[name] = r
    dom_name = max((k for k in kern if kern[k]["bound"] != "mixed"), key=lambda k: kern[k]["ms_per_step"])
    domr = dict(kern[dom_name])
    dom_entries = groups[dom_name][0]
    traffic = None
    try:
        per_pt = [tj["dram_bytes_per_point"][e] * ncalls.get(e, 0) for e in dom_entries]
        traffic = sum(per_pt) / max(1e-9, sum(ncalls.get(e, 0) for e in dom_entries)) * pts_call
    except Exception:
        pass
    bytes_call = {"xw_interior_forward": n_loc * (2 * d * 4 + 2 * L_T * 4 + (d + 1) * 4) + 2 * pts_call * 4,
                  "xw_interior_forward:cached_v": n_loc * (d * 4 + (d + 1) * 4) + 2 * pts_call * 4 + 4 * pts_call * 4,
                  "xw_interior_backward_v": n_loc * d * 4 + pts_call * 4,
                  "xw_interior_backward_u": n_loc * d * 4 + pts_call * 4 + n_loc * 4,
                  "xw_boundary_u": n_loc * d * 4 + pts_call * 4 + n_loc * 4}
    alg_bytes = sum(bytes_call[e] * ncalls.get(e, 0) for e in dom_entries) / max(1e-9, sum(ncalls.get(e, 0) for e in dom_entries))
    step_flops = (2 * (fl_["u_interior"] + fl_["u_boundary"]) + fl_["v_interior"]) * n_loc * L_T
    # FLOPs of the work that actually ran (the test-function net is evaluated once per step, not three times)
    step_flops_run = sum(fl_[e] * pts_call * ncalls.get(e, 0) for e in per_step if e in fl_)
    roofline = {"bound": "tensor" if domr["bound"] == "tensor" else "fp32_fma", "kernel": dom_name, "achieved": domr["achieved"], "peak": domr["peak"],
                "unit": "TFLOP/s", "frac": domr["frac"], "traffic": traffic,
                "traffic_source": (tj.get("source", "") + " (static: from a committed ncu capture, not this run)") if traffic is not None else None,
                "algorithmic_bytes": alg_bytes, "peak_source": domr["peak_source"], "note": domr.get("note"),
                "points_per_call": pts_call, "share_of_step": domr["ms_per_step"] / ms,
                "kernels": kern,
                "step_algorithmic_tflops_run": step_flops_run / (ms * 1e-3) / 1e12,
                "step_algorithmic_tflops_uncached": step_flops / (ms * 1e-3) / 1e12,
                "step_note": "whole-step algorithmic FLOP/s over two different pipes (context only): `_run` charges the test-function "
                             "forward once per step (it IS evaluated once: the 2nd u-step and the v-step read the cache), `_uncached` "
                             "charges it in all three sub-steps as the reference does; FP32 FFMA peak measured here: %.1f TFLOP/s" % fma_peak,
                "hbm_context": {"algorithmic_GBs": alg_bytes / (domr["ms_per_step"] / max(1e-9, domr["calls_per_step"]) * 1e-3) / 1e9,
                                "peak_GBs": peaks.get("hbm_gbs"), "peak_source": "MEASURED_PEAKS.json"}}
    # u-step and v-step separately (sums of the CUDA-event times of the C-ABI entries a sub-step calls; the
    # optimiser launch and the glue between the entries are in `ms_per_step`, not here)
    def _phase(entries, pts):
        t = sum(per_call.get(e, 0.0) for e in entries)
        return {"entries": entries, "ms": t, "path_points": pts, "value": pts / (t * 1e-3) if t > 0 else None, "unit": "path-points/s"}
    pts_u, pts_v = (n_loc + n_loc) * L_T, n_loc * L_T
    fwd_cached = "xw_interior_forward:cached_v" if "xw_interior_forward:cached_v" in per_call else "xw_interior_forward"
    phases = {"u_step_test_function_evaluated": _phase(["xw_interior_forward", "xw_boundary_u", "xw_interior_backward_u"], pts_u),
              "u_step_test_function_cached": _phase([fwd_cached, "xw_boundary_u", "xw_interior_backward_u"], pts_u),
              "v_step": _phase([fwd_cached, "xw_interior_backward_v"], pts_v),
              "note": "one step = the reference iteration n1 = 2 u-steps + n2 = 1 v-step on one sample; a u-step processes "
                      "(N_r + N_b) N_t path-points, a v-step N_r N_t; the first u-step evaluates the test-function net, the "
                      "second u-step and the v-step read its cached values (unchanged v parameters)"}
    line = {"metric": "weak-loss+grad path-points/sec", "value": pp_step / (ms * 1e-3), "unit": "path-points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args),
            "e2e": {"value": pp_step / (ms_e2e * 1e-3), "unit": "path-points/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "loss_u": last.get("lu"), "loss_v": last.get("lv")},
            "e2e_nlc": e2e_nlc,
            "strong": strong,
            "variants": variants,
            "shard_check": check,
            "gpu_launches": launches,
            "roofline": roofline,
            "kernels_launched": {"xnode_generation": xnode_impl, "vnet_forward": vf, "vnet_backward": vb},
            "phases": phases,
            "kernels_ms_per_step": per_step, "kernels_ms_per_call": per_call,
            "kernels_alg_tflops": {k: fl_[k] * pts_call / (per_call[k] * 1e-3) / 1e12 for k in per_call if k in fl_},
            "clocks": clk}
    if world == 1 and not args.no_ttt:
        line["time_to_target"] = time_to_target(xw, dev, list(range(args.ttt_seeds)))
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        n = args.cpu_paths
        dt, pp = cpu_step_time(d, n, n, 1, 1, threads)
        line["cpu_baseline"] = {"value": pp / dt, "unit": "path-points/s", "cores": threads, "kind": "port",
                                "sample": "d=%d, N_r=N_b=%d paths, N_t=%d, fp64 PyTorch CPU port of the reference path "
                                          "(oracle/torch_port.py), 1 warm-up + 1 timed step" % (d, n, L_T)}
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="m", choices=sorted(CONFIGS), help="m: BASELINE configs[3] (d=20, 2^20 paths); xl: configs[4] (d=100, 2^22)")
    ap.add_argument("--log2n", type=int, default=None, help="log2 of interior (= boundary) paths per rank (overrides --config)")
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--cpu-paths", type=int, default=None, help="paths of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ttt", action="store_true", help="skip the time-to-target runs (N=1 only)")
    ap.add_argument("--ttt-seeds", type=int, default=5)
    ap.add_argument("--graph", type=int, default=0, help="1: CUDA-graph replay of the sub-steps (NODE_WAN_solver(use_cuda_graph=True))")
    ap.add_argument("--fused", type=int, default=1, help="1: flat parameter buffers + single-launch Adam (fused_optimizer=True)")
    ap.add_argument("--nlc-max-gb", type=float, default=8.0, help="skip the [N,L,C]-layout e2e arm above this many GB of pinned host memory")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    args.dim = args.dim if args.dim is not None else cfg["dim"]
    args.log2n = args.log2n if args.log2n is not None else cfg["log2n"]
    if args.cpu_paths is None:
        args.cpu_paths = 4096 if args.dim <= 20 else 1024      # dense a[d,d,N,L] of the reference: 0.8 GB at d=100, N=1024
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
