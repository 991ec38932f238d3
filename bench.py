#!/usr/bin/env python
"""bench.py -- weak-loss+grad path-points/s of the XNODE-WAN hot path on B200.

One "step" = the hot part of one reference training iteration on ONE sample
(/root/reference/src/training.py:125-162): n1=2 u-steps (loss_u + theta_u gradients + Adam) then
n2=1 v-step (loss_v + theta_v gradients + Adam), coefficient evaluation included.  A path-point is
one (path, time-index) sample; a u-step covers (N_r + N_b) * N_t of them, a v-step N_r * N_t.

  python bench.py --gpus N --steps K --warmup W             # this framework (sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[3]): cube PDE (Ex4_1), d=20, N_t=20, N_r = N_b = 2^20 paths PER
RANK (weak scaling), alpha=1e8, shipped network sizes, synthetic seeded samples, xavier weights.
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


def alg_flops(d, H=20, hh=10, nu=8, Hv=50, nv=9, L=L_T):
    """algorithmic FLOPs per path-point (un-hoisted dense counts, MAC = 2 FLOP)"""
    F = (H + d + 1) * hh + (nu - 1) * hh * hh + hh * H
    U = 2 * (L - 1) / L * F + H + (H + 2 * H * H) / L
    Vm = (d + 1) * Hv + nv * Hv * Hv + Hv
    return {"u_interior": 2 * (4 * U + 2 * Vm), "u_boundary": 2 * 3 * U, "v_interior": 2 * (2 * U + 4 * Vm),
            "U": U, "Vm": Vm,
            # per C-ABI call (what one launch group computes), per point of its own batch
            "xw_interior_forward": 2 * (2 * U + 2 * Vm), "xw_boundary_u": 2 * 3 * U,
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
                    "the port is pinned to it by golden vectors); the reference cannot run the full 2^20-path "
                    "config (dense a[d,d,N,L] = 33.5 GB), so a bounded sample is timed and quoted per path-point"}
    print(json.dumps(line), flush=True)


def config_of(args):
    return {"workload": "cube PDE Ex4_1 d=%d, N_r=N_b=2^%d paths per rank, N_t=%d, alpha=1e8, H=20 hh=10 nu=8 Hv=50 nv=9, "
                        "midpoint; step = 2 u-steps + 1 v-step on one sample" % (args.dim, args.log2n, L_T),
            "dim": args.dim, "N_r_per_rank": 1 << args.log2n, "N_b_per_rank": 1 << args.log2n, "N_t": L_T,
            "layout": "collapsed (times[L] + x[N,d]); kernels also take the reference [N,L,C] layout",
            "l2": "inputs_exceed_L2 (3 x %.0f MB of coordinates + 2 x %.0f MB of per-point seeds per step vs 126 MB L2)"
                  % ((1 << args.log2n) * args.dim * 4 / 1e6, (1 << args.log2n) * L_T * 4 / 1e6),
            "arithmetic": "fp32 results: XNODE kernels FP32 FFMA (+ mma.sync 3xTF32 for the weight-gradient outer products); "
                          "test-function net on tcgen05 kind::tf32 with 3xTF32 error compensation (1e-6 vs fp64; "
                          "XW_VNET_IMPL=tile selects the pure-FP32 kernels)",
            "parallelism": "dp%d (paths sharded, 2 small all-reduces per sub-step)" % args.gpus}


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
    n_loc = 1 << args.log2n
    d = args.dim
    prob = xw.problems.ex4_1()
    params = xw.problems.cube_params(dim=d, N_r=n_loc * world, N_b=n_loc * world, N_t=L_T, shape_param=[-1.0, 1.0])
    torch.manual_seed(0)
    solver = xw.NODE_WAN_solver(params, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g,
                                dev, "./", func_u_sol=prob.func_u_sol, p=2, log_json=False)
    torch.manual_seed(1000 + rank)
    domain = solver.new_domain(sample_device=dev, collapsed=True)
    points = xw.Comb_loader(n_loc, n_loc, domain, dev)
    points[0]
    # pinned host copy of the same sample for the end-to-end arm
    host = [t.to("cpu").pin_memory() for t in (points.interioru, points.interiorv, points.boundary)]
    pp_step = (2 * (n_loc + n_loc) + n_loc) * L_T * world

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

    def step_resident():
        solver.train_iteration(domain, points)

    last = {}

    def step_e2e():
        pts = xw.Comb_loader.from_tensors(host[0], host[1], host[2], dev)
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
    prof, calls = hp.PROFILE, dict(hp.CALLS)
    hp.PROFILE = None
    clk = clocks.stop() if rank == 0 else None
    # per-entry-point device time (CUDA events on the launching stream, inside the timed region)
    per_call = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in prof.items()}
    per_step = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in prof.items()}
    launches = hp.LAUNCHES[0] // args.steps

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    h2d = sum(t.nbytes() for t in host)
    d2h = 16

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
    ncalls = {k: len(v) / args.steps for k, v in prof.items()}
    tj = {}
    try:       # DRAM bytes per point of each kernel from the committed `ncu --set full` captures
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
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
    groups = {
        # kernel -> (C-ABI entries it serves, bound)
        "k_xnode_bwd": (["xw_boundary_u", "xw_interior_backward_u"], "fp32_fma"),
        "k_vnet_tc_bwd3": (["xw_interior_backward_v"], "tensor"),
        "interior_forward (k_xnode_fwd + k_vnet_tc_fwd | k_weak_combine + k_vnet_points<row0>)": (["xw_interior_forward"], "mixed"),
    }
    kern = {}
    for name, (entries, bound) in groups.items():
        t_ms = sum(per_step.get(e, 0.0) for e in entries)
        if t_ms <= 0:
            continue
        alg = sum(fl_[e] * pts_call * ncalls.get(e, 0) for e in entries)
        launches_k = sum(ncalls.get(e, 0) for e in entries)
        r = {"bound": bound, "ms_per_step": t_ms, "launches_per_step": launches_k,
             "algorithmic_tflops": alg / (t_ms * 1e-3) / 1e12}
        if bound == "fp32_fma":
            r.update(achieved=r["algorithmic_tflops"], peak=fma_peak, unit="TFLOP/s", frac=r["algorithmic_tflops"] / fma_peak,
                     peak_source="xw_fma_probe (FFMA chains) measured in this run; nominal 74.4")
        elif bound == "tensor":
            ex = tc_bwd_exec * pts_call * launches_k / (t_ms * 1e-3) / 1e12
            r.update(achieved=ex, peak=tf32_peak, unit="TFLOP/s", frac=ex / tf32_peak,
                     executed_flops_per_point=tc_bwd_exec,
                     peak_source="MEASURED_PEAKS.json bf16_tflops (burst) / 2: kind::tf32 runs at half the bf16 rate; "
                                 "achieved counts EXECUTED tensor FLOPs (3 MMAs per product, 128 x 56 x 56 padded tiles)")
        kern[name] = r
    dom_name = max((k for k in kern if kern[k]["bound"] != "mixed"), key=lambda k: kern[k]["ms_per_step"])
    domr = dict(kern[dom_name])
    dom_entries = groups[dom_name][0]
    traffic = None
    try:
        per_pt = [tj["dram_bytes_per_point"][tj["entry_to_kernel"][e]] * ncalls.get(e, 0) for e in dom_entries]
        traffic = sum(per_pt) / max(1e-9, sum(ncalls.get(e, 0) for e in dom_entries)) * pts_call
    except Exception:
        pass
    bytes_call = {"xw_interior_forward": n_loc * (2 * d * 4 + 2 * L_T * 4 + (d + 1) * 4) + 2 * pts_call * 4,
                  "xw_interior_backward_v": n_loc * d * 4 + pts_call * 4,
                  "xw_interior_backward_u": n_loc * d * 4 + pts_call * 4 + n_loc * 4,
                  "xw_boundary_u": n_loc * d * 4 + pts_call * 4 + n_loc * 4}
    alg_bytes = sum(bytes_call[e] * ncalls.get(e, 0) for e in dom_entries) / max(1e-9, sum(ncalls.get(e, 0) for e in dom_entries))
    step_flops = (2 * (fl_["u_interior"] + fl_["u_boundary"]) + fl_["v_interior"]) * n_loc * L_T
    roofline = {"bound": domr["bound"], "kernel": dom_name, "achieved": domr["achieved"], "peak": domr["peak"],
                "unit": "TFLOP/s", "frac": domr["frac"], "traffic": traffic, "traffic_source": tj.get("source"),
                "algorithmic_bytes": alg_bytes, "peak_source": domr["peak_source"],
                "points_per_launch": pts_call, "share_of_step": domr["ms_per_step"] / ms,
                "kernels": kern,
                "step_algorithmic_tflops": step_flops / (ms * 1e-3) / 1e12,
                "step_note": "whole-step algorithmic FLOP/s; the test-function net runs on the tensor cores (3xTF32), so this "
                             "is context, not a fraction of one pipe's peak (FP32 FFMA peak measured here: %.1f TFLOP/s)" % fma_peak,
                "hbm_context": {"algorithmic_GBs": alg_bytes / (domr["ms_per_step"] / max(1e-9, domr["launches_per_step"]) * 1e-3) / 1e9,
                                "peak_GBs": peaks.get("hbm_gbs"), "peak_source": "MEASURED_PEAKS.json"}}
    line = {"metric": "weak-loss+grad path-points/sec", "value": pp_step / (ms * 1e-3), "unit": "path-points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args),
            "e2e": {"value": pp_step / (ms_e2e * 1e-3), "unit": "path-points/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "loss_u": last.get("lu"), "loss_v": last.get("lv")},
            "gpu_launches": launches,
            "roofline": roofline,
            "kernels_ms_per_step": per_step, "kernels_ms_per_call": per_call,
            "kernels_alg_tflops": {k: fl_[k] * pts_call / (per_call[k] * 1e-3) / 1e12 for k in per_call},
            "clocks": clk}
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
    ap.add_argument("--log2n", type=int, default=20, help="log2 of interior (= boundary) paths per rank")
    ap.add_argument("--dim", type=int, default=20)
    ap.add_argument("--cpu-paths", type=int, default=4096, help="paths of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
