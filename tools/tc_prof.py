"""Where the three warp roles of k_vnet_tc_bwd3 wait (development tool, not part of the product or the tests).

`python tools/tc_prof.py build` (build container) compiles tools/tcb/tcb.cu -- the kernel alone, from the product's own
xw_vnet_tc.cuh, seconds instead of the 3 minutes of the product library -- twice: plain (timing) and with -DXW_TC_PROF,
where every mbarrier / named-barrier wait of the kernel is bracketed by clock64() and the cycles are summed per role.
`python tools/tc_prof.py [log2n] [dim]` (GPU box) runs the kernel on synthetic collapsed-layout points, times the plain
build with CUDA events and prints the per-role wait shares of the instrumented one."""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "xnode-wan-pde-solver_b200")
TCB = os.path.join(ROOT, "tools", "tcb")

_F = ["prev tile consumed (mFC, mPC)", "bar: input operand stored", "mF input-layer MMAs", "bar: layer operand stored",
      "mF hidden-layer MMAs", "bar: output dot", "MMA issue", "total", "tmem_ld28", "tmem_st A_f + wait::st", "-", "-"]
_R = ["mFD (F's delta_nv)", "bar: A_r stored", "mPh (P done with the delta image)", "mR R-op MMAs", "-", "-", "MMA issue", "total",
      "split + tmem_st A_r + wait::st", "delta image STS + fence", "tmem_ld56", "apply masks"]
_P = ["mFD (F's tile)", "mPh own previous P-op", "mDPh delta image", "bar: half images stored", "mPh tile end", "flush",
      "MMA issue", "total", "r image STS + fence", "-", "-", "-"]
SLOTS = {"F issuer": _F, "R issuer": _R, "P half 0 issuer": _P, "P half 1 issuer": _P,
         "F warp 1": _F, "R warp 1": _R, "P half 0 warp 1": _P, "P half 1 warp 3": _P}


def build():
    from importlib import import_module
    b = import_module("xnode-wan-pde-solver_b200.build")
    flags = [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    split = ["-DXW_TC_BWD_SPLIT=%s" % os.environ.get("XW_TC_BWD_SPLIT", "0"), "-DXW_TC_SPLIT_MASK=%s" % os.environ.get("XW_TC_SPLIT_MASK", "3")]
    suffix = os.environ.get("TCB_SUFFIX", "")
    for name, extra in (("libtcb%s.so" % suffix, split), ("libtcb_prof%s.so" % suffix, ["-DXW_TC_PROF"] + split)):
        cmd = ["nvcc"] + flags + extra + ["-o", os.path.join(TCB, name), os.path.join(TCB, "tcb.cu")]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise SystemExit(r.stdout[-3000:])
        print(os.path.join(TCB, name))


def run(log2n=20, d=20, L=20, reps=3, packed=0, flush=4):
    import numpy as np
    import torch
    dev = torch.device("cuda:0")
    n = 1 << log2n
    Hv, nv = 50, 9
    g = torch.Generator(device=dev).manual_seed(1)
    Pv = Hv * (d + 1) + Hv + Hv * Hv + Hv + Hv + 1
    thv = (torch.rand(Pv, device=dev, generator=g) - 0.5) * 0.35
    x = torch.rand(n, d, device=dev, generator=g) * 2 - 1
    times = torch.sort(torch.rand(L, device=dev, generator=g))[0].contiguous()
    cot = torch.randn(n * L, device=dev, generator=g)
    kv = torch.tensor([1e-3, 1e-3, 1.0], dtype=torch.float64, device=dev)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    st = torch.cuda.current_stream().cuda_stream
    out = {"log2n": log2n, "dim": d, "packed": packed, "flush_tiles": flush}
    for name in ("libtcb.so", "libtcb_prof.so"):
        lib = C.CDLL(os.path.join(TCB, name.replace(".so", os.environ.get("TCB_SUFFIX", "") + ".so")))
        lib.tcb_workspace_bytes.restype = C.c_size_t
        lib.tcb_run.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong,
                                               C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        ws = torch.zeros(lib.tcb_workspace_bytes(d, Hv, nv, sms), dtype=torch.uint8, device=dev)

        def call():
            rc = lib.tcb_run(d, Hv, nv, n, L, thv.data_ptr(), times.data_ptr(), 0, 1, x.data_ptr(), d, 0, cot.data_ptr(),
                             kv.data_ptr(), ws.data_ptr(), packed, flush, None, st)
            assert rc == 0
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        prof = name.endswith("prof.so")
        buf = (C.c_ulonglong * 96)()
        if prof:
            lib.tcb_prof_read(C.cast(buf, C.c_void_p))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if not prof:
            out["ms_per_call"] = ms
            continue
        out["ms_per_call_instrumented"] = ms
        lib.tcb_prof_read(C.cast(buf, C.c_void_p))
        v = np.array(list(buf), dtype=np.float64).reshape(8, 12)
        ntiles = (n * L + 127) // 128
        out["tiles_per_cta"] = ntiles / sms
        out["roles"] = {}
        for r, rname in enumerate(SLOTS):
            tot = v[r, 7]
            out["roles"][rname] = {"cycles_per_tile": round(tot / reps / ntiles, 1),
                                   "wait_share": {SLOTS[rname][s]: round(v[r, s] / tot, 4) for s in range(12) if s != 7 and SLOTS[rname][s] != "-"}}
    print(json.dumps(out, indent=1))


def acc(log2n=18, d=20, L=20, flushes=(1, 4)):
    """gradient of the kernel (per-CTA partials summed in fp64) against a PyTorch fp64 autograd evaluation of the same net
    on the same points, per parameter tensor (rel-L2), for several flush intervals of the weight-gradient accumulators"""
    import torch
    dev = torch.device("cuda:0")
    n = 1 << log2n
    Hv, nv, Cc = 50, 9, d + 1
    g = torch.Generator(device=dev).manual_seed(1)
    sizes = [("Wi", Hv * Cc), ("bi", Hv), ("Wh", Hv * Hv), ("bh", Hv), ("Wz", Hv), ("bz", 1)]
    Pv = sum(k for _, k in sizes)
    sc = float(os.environ.get("TCB_SCALE", "1"))
    thv = (torch.rand(Pv, device=dev, generator=g) - 0.5) * 0.35 * sc
    x = (torch.rand(n, d, device=dev, generator=g) * 2 - 1) * sc
    times = torch.sort(torch.rand(L, device=dev, generator=g))[0].contiguous()
    cot = torch.randn(n * L, device=dev, generator=g)
    k0, k1 = 1e-3, 2e-3
    kv = torch.tensor([k0, k1, 0.0], dtype=torch.float64, device=dev)
    # fp64 reference, in chunks of paths
    th = thv.double().requires_grad_(True)
    parts, o = {}, 0
    for name, k in sizes:
        parts[name] = th[o:o + k]; o += k
    Wi, Wh = parts["Wi"].view(Hv, Cc), parts["Wh"].view(Hv, Hv)
    ref = torch.zeros(Pv, dtype=torch.float64, device=dev)
    ch = 1 << 14
    for s0 in range(0, n, ch):
        xs = x[s0:s0 + ch].double()
        pts = torch.cat((times.double().view(1, L, 1).expand(xs.shape[0], L, 1), xs.unsqueeze(1).expand(-1, L, -1)), 2).reshape(-1, Cc)
        a = pts @ Wi.T + parts["bi"]
        for _ in range(nv):
            a = torch.relu(a) @ Wh.T + parts["bh"]
        v = torch.tanh(a) @ parts["Wz"] + parts["bz"]
        G = (k0 * cot[s0 * L:(s0 + xs.shape[0]) * L].double() + k1 * v).detach()
        ref += torch.autograd.grad((G * v).sum(), th)[0]
    lib = C.CDLL(os.path.join(TCB, "libtcb%s.so" % os.environ.get("TCB_SUFFIX", "")))
    lib.tcb_workspace_bytes.restype = C.c_size_t
    lib.tcb_run.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong,
                                           C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    ws = torch.zeros(lib.tcb_workspace_bytes(d, Hv, nv, sms), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    res = {"log2n": log2n, "dim": d, "tiles_per_cta": (n * L + 127) // 128 / sms, "rel_l2_vs_fp64": {}}
    for fl in flushes:
        gv = torch.zeros(Pv, device=dev)
        rc = lib.tcb_run(d, Hv, nv, n, L, thv.data_ptr(), times.data_ptr(), 0, 1, x.data_ptr(), d, 0, cot.data_ptr(),
                         kv.data_ptr(), ws.data_ptr(), int(os.environ.get("XW_TC_TMEM_PACKED", "0")), fl, gv.data_ptr(), st)
        assert rc == 0
        torch.cuda.synchronize()
        e, o = {}, 0
        for name, k in sizes:
            r_ = ref[o:o + k]
            e[name] = float((gv[o:o + k].double() - r_).norm() / r_.norm()); o += k
        res["rel_l2_vs_fp64"]["flush_every_%s" % ("never" if fl >= 1 << 30 else fl)] = e
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    elif len(sys.argv) > 1 and sys.argv[1] == "acc":
        acc(*(int(a) for a in sys.argv[2:4]))
    else:
        run(*(int(a) for a in sys.argv[1:3]), packed=int(os.environ.get("XW_TC_TMEM_PACKED", "0")),
            flush=int(os.environ.get("XW_TC_FLUSH", "4")))
