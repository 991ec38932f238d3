"""Host side of the fused hot path: one `torch.autograd.Function` per phase that replaces
/root/reference/src/training.py:129-137 (u-phase) and :153-161 (v-phase).

forward  : xw_interior_forward (+ xw_boundary_u in the u-phase) -> Monte-Carlo partial sums
           -> [one small all-reduce when sharded] -> loss scalar (fp64, on device, no host sync)
backward : xw_interior_backward_u / xw_interior_backward_v -> flat fp32 gradient
           -> [one small all-reduce] -> per-parameter .grad in the parameters' dtype

PyTorch is plumbing here (device memory, streams, torch.distributed); all arithmetic of the path
runs in the sm_100a kernels behind the C ABI (include/xnode_wan_b200.h).  No CPU fallback: CPU
tensors or a missing library raise.
"""
import contextlib
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import _lib

# optional per-call CUDA-event timing (bench.py's roofline): PROFILE = {} to switch on;
# maps C-ABI entry name -> list of (start_event, end_event) recorded on the launching stream
PROFILE = None
CALLS = {}                    # C-ABI entry name -> number of calls (always counted)
KERNELS_PER_CALL = {"xw_interior_forward": 2, "xw_boundary_u": 3, "xw_interior_backward_u": 3,
                    "xw_interior_backward_v": 2, "xw_xnode_eval": 1, "xw_vnet_eval": 1, "xw_loss_scalars": 1}


LAUNCHES = [0]                # kernels launched by this package (counted per C-ABI call)


def _call(lib, name, dev, *args):
    CALLS[name] = CALLS.get(name, 0) + 1
    k = KERNELS_PER_CALL[name]
    key = name
    if name == "xw_interior_forward":          # xnode_fwd + (row-0 kernel + tiled pass | combine from the cache)
        cached = args[-4] == 2
        k = 2 if cached else 3
        if cached:
            key = name + ":cached_v"           # the test-function net is NOT evaluated in this call
    LAUNCHES[0] += k
    if PROFILE is not None and dev.type == "cuda":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(dev))
        lib.call(name, *args)
        e1.record(torch.cuda.current_stream(dev))
        PROFILE.setdefault(key, []).append((e0, e1))
    else:
        lib.call(name, *args)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(dev):
    if dev.type != "cuda":
        return None
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _check_dev(t, what, lib):
    if not t.is_cuda and not lib.host_memory:
        raise RuntimeError("%s must be a CUDA tensor: the XNODE-WAN hot path has no CPU fallback" % what)


@dataclass
class NetSpec:
    """network sizes (configs/cube_pde.yaml keys) -> xw_dims"""
    d: int
    H: int
    hh: int
    nu: int
    Hv: int
    nv: int
    solver: str = "midpoint"

    def c(self):
        if self.solver not in _lib.SOLVERS:
            raise RuntimeError("unsupported solver %r (supported: euler, midpoint, rk4)" % (self.solver,))
        return _lib.Dims(self.d, self.H, self.hh, self.nu, self.Hv, self.nv, _lib.SOLVERS[self.solver])


@dataclass
class DomainSpec:
    kind: str                      # cube | cone | hourglass
    p: tuple = (0.0, 0.0, 0.0)
    V: float = 1.0

    def c(self):
        p = tuple(float(x) for x in self.p) + (0.0, 0.0, 0.0)
        return _lib.Domain(_lib.DOMAINS[self.kind], p[0], p[1], p[2])


@dataclass
class CoefSpec:
    """structured PDE coefficients (include/xnode_wan_b200.h, xw_coef): c(X,u) = c0 + c1*u; a None=identity, a constant
    [d,d] matrix or per-path values [N,d,d] (a_per_path); b None=0, a constant [d] vector or per-path [N,d];
    A_val / A_der: optional per-point [N*L] values of A(u) = c(X,u) u and dA/du for a general c (replace c0, c1)"""
    c0: float = 0.0
    c1: float = 0.0
    a: Optional[torch.Tensor] = None
    b: Optional[torch.Tensor] = None
    a_per_path: bool = False
    b_per_path: bool = False
    A_val: Optional[torch.Tensor] = None
    A_der: Optional[torch.Tensor] = None

    def c(self):
        a_sn = self.a.shape[-1] * self.a.shape[-2] if (self.a is not None and self.a_per_path) else 0
        b_sn = self.b.shape[-1] if (self.b is not None and self.b_per_path) else 0
        return _lib.Coef(float(self.c0), float(self.c1), self.a.data_ptr() if self.a is not None else None,
                         self.b.data_ptr() if self.b is not None else None, a_sn, b_sn,
                         self.A_val.data_ptr() if self.A_val is not None else None,
                         self.A_der.data_ptr() if self.A_der is not None else None)


def as_f32(t):
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


@dataclass
class Batch:
    """one (datau, datav, bdata) triple of Comb_loader plus the coefficient values on it.
    X/XV/BX keep the reference layout [N, L, C] (time in channel 0) or the collapsed layout
    (x[N,d] + times[L]); everything is fp32 on the device."""
    N: int
    L: int
    d: int
    times: torch.Tensor            # [L]   path 0's time grid (src/model.py:92)
    x: torch.Tensor                # base tensor holding X
    x_off: int
    x_sn: int
    xv: torch.Tensor               # base tensor holding XV
    tv_off: int
    tv_sn: int
    tv_sl: int
    xv_off: int
    xv_sn: int
    xv_sl: int
    tv: Optional[torch.Tensor] = None   # separate tensor for XV times (collapsed layout)
    h: Optional[torch.Tensor] = None        # [N]   func_h(X[:,0,:])
    s0: Optional[torch.Tensor] = None       # [N]   initial scalar when it is not h (batch not starting at T0)
    grad_h: Optional[torch.Tensor] = None   # [N, d]
    f: Optional[torch.Tensor] = None        # [N, L]
    # boundary
    Nb: int = 0
    Lb: int = 0
    times_b: Optional[torch.Tensor] = None
    xb: Optional[torch.Tensor] = None
    xb_off: int = 0
    xb_sn: int = 0
    sb: Optional[torch.Tensor] = None       # [Nb]
    g: Optional[torch.Tensor] = None        # [Nb, Lb]
    # global (all-rank) counts used in the normalisers
    N_glob: int = 0
    Nb_glob: int = 0
    keep: list = field(default_factory=list)

    def points(self):
        tbase = self.tv if self.tv is not None else self.xv
        return _lib.Points(tbase.data_ptr() + 4 * self.tv_off, self.tv_sn, self.tv_sl,
                           self.xv.data_ptr() + 4 * self.xv_off, self.xv_sn, self.xv_sl)


_BATCH_VALUE_FIELDS = ("times", "tv", "h", "s0", "grad_h", "f", "times_b", "sb", "g")


def copy_batch_values_(dst, src):
    """copy the sample-dependent VALUES of `src` into the tensors of `dst` (same shapes): used to keep one persistent
    home for them across CUDA graphs (training.py); path tensors (x, xv, xb) are views of the caller's inputs"""
    for name in _BATCH_VALUE_FIELDS:
        a, b = getattr(dst, name), getattr(src, name)
        if a is None or b is None:
            if (a is None) != (b is None):
                raise RuntimeError("batch structure changed (%s)" % name)
            continue
        if a.data_ptr() != b.data_ptr():
            a.copy_(b)


def batch_from_reference_layout(X, XV, BX=None):
    """X, XV [N,L,C], BX [Nb,Lb,C] (any float dtype) -> Batch (fp32 views, no repack when already fp32)"""
    Xf, XVf = as_f32(X), as_f32(XV)
    N, L, Cc = Xf.shape
    b = Batch(N=N, L=L, d=Cc - 1, times=Xf[0, :, 0].contiguous(), x=Xf, x_off=1, x_sn=L * Cc,
              xv=XVf, tv_off=0, tv_sn=L * Cc, tv_sl=Cc, xv_off=1, xv_sn=L * Cc, xv_sl=Cc, N_glob=N)
    if BX is not None:
        BXf = as_f32(BX)
        Nb, Lb, _ = BXf.shape
        b.Nb, b.Lb, b.times_b, b.xb, b.xb_off, b.xb_sn, b.Nb_glob = Nb, Lb, BXf[0, :, 0].contiguous(), BXf, 1, Lb * Cc, Nb
    return b


def batch_from_collapsed(times, x, xv, xb=None, times_b=None):
    """collapsed layout: x, xv [N,d], xb [Nb,d], shared times[L] (the cube's repeated [N,L,C]
    tensors carry no more information, src/dataset.py:252-254)"""
    times, x, xv = as_f32(times), as_f32(x), as_f32(xv)
    N, d = x.shape
    L = times.numel()
    b = Batch(N=N, L=L, d=d, times=times, x=x, x_off=0, x_sn=d, xv=xv, tv=times, tv_off=0, tv_sn=0, tv_sl=1,
              xv_off=0, xv_sn=d, xv_sl=0, N_glob=N)
    if xb is not None:
        xb = as_f32(xb)
        tb = as_f32(times_b) if times_b is not None else times
        b.Nb, b.Lb, b.times_b, b.xb, b.xb_off, b.xb_sn, b.Nb_glob = xb.shape[0], tb.numel(), tb, xb, 0, d, xb.shape[0]
    return b


class _Workspace:
    """per-device scratch reused across calls (grown on demand; allocated by torch's caching allocator)"""

    def __init__(self):
        self.buf = {}
        self.graph_buf = {}

    def reserve_for_graphs(self, dev, nbytes):
        """a persistent scratch, allocated EAGERLY (outside any capture), that the graphs captured afterwards use: graphs
        replay one after the other on one stream, so they can share it -- a private copy per graph and per call is
        3 x 1.8 GB x 4 graphs at 2^20 paths and does not fit at d = 100, 2^22 paths (17 GB each).  Whoever captures keeps
        the returned tensor alive for as long as its graphs live; a larger request replaces the buffer for LATER captures."""
        b = self.graph_buf.get(dev)
        if b is None or b.numel() < nbytes:
            b = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
            self.graph_buf[dev] = b
        return b

    def get(self, dev, nbytes):
        if dev.type == "cuda" and torch.cuda.is_current_stream_capturing():
            b = self.graph_buf.get(dev)
            if b is not None and b.numel() >= nbytes:
                return b
            # otherwise the graph owns its scratch (the shared eager buffer below may be re-grown later)
            return torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        b = self.buf.get(dev)
        if b is None or b.numel() < nbytes:
            b = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
            self.buf[dev] = b
        return b


_WS = _Workspace()
_WS_SIDE = _Workspace()           # workspace of the boundary pass when it runs beside the interior forward (forward_sums)
_SIDE_STREAMS = {}
# largest sample whose boundary pass runs beside the interior forward (None: any).  Measured on B200 (profiles/README
# r02bf): at the shipped N = 4000 a launch is less than one wave and the two passes simply run side by side (-12 % per
# outer iteration); at 2^17 paths per rank (configs[3] split over 8 GPUs) the tail of one kernel overlaps the head of
# the next, 11.28 -> 10.94 ms per step (10.58 under graph replay, where the join is deferred); at 2^20 77.1 -> 76.7 ms.
# Eagerly the limit stays at one wave: bench.py times every C-ABI entry with CUDA events inside the timed region, and
# overlapping entries would blur the per-kernel roofline figures for 0.6 % of the step; under capture (no per-call events)
# any size runs the two branches.
CONCURRENT_BOUNDARY_MAX_PATHS = 16384
CONCURRENT_BOUNDARY_MAX_WORKSPACE = 4 << 30


def _side_stream(dev):
    if dev not in _SIDE_STREAMS:
        _SIDE_STREAMS[dev] = torch.cuda.Stream(dev)
    return _SIDE_STREAMS[dev]


DEFER_BOUNDARY_JOIN = os.environ.get("XW_DEFER_BOUNDARY_JOIN", "1") == "1"


def _is_distributed(group):
    return group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                 and torch.distributed.get_world_size() > 1)


def _concurrent_boundary(n, nb, ws_bytes):
    flag = os.environ.get("XW_CONCURRENT_BOUNDARY")
    if flag is not None:
        return flag == "1"
    # (every captured graph owns its workspaces: a second one of 17 GB per graph -- d = 100, 2^22 paths -- does not fit)
    return max(n, nb) <= CONCURRENT_BOUNDARY_MAX_PATHS or \
        (torch.cuda.is_current_stream_capturing() and ws_bytes <= CONCURRENT_BOUNDARY_MAX_WORKSPACE)


def flatten_params(params):
    """flat fp32 parameter vector in named_parameters() order.  Parameters that live in one flat buffer
    (optim.FlatParameters) come back without re-packing."""
    from .optim import flat_of
    fp = flat_of(list(params))
    if fp is not None:
        return fp.theta32()
    return torch.cat([p.detach().reshape(-1) for p in params]).float()


def unflatten_like(flat, meta):
    """meta: list of (shape, dtype)"""
    out, o = [], 0
    for shape, dtype in meta:
        n = 1
        for s_ in shape:
            n *= s_
        out.append(flat[o:o + n].reshape(shape).to(dtype))
        o += n
    return out


def _allreduce(t, group):
    if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                             and torch.distributed.get_world_size() > 1):
        torch.distributed.all_reduce(t, group=group)


def vcache_buffer(lib, spec, batch, dev):
    """device buffer for the test-function cache of xw_interior_forward (see include/xnode_wan_b200.h)"""
    dims = spec.c()
    n = lib.cdll.xw_vcache_floats(C.byref(dims), batch.N, batch.L)
    return torch.empty(int(n), dtype=torch.float32, device=dev)


def forward_sums(lib, spec, dom, coef, theta_u, theta_v, batch, with_boundary, alpha, boundary_grad=None,
                 vcache=None, vmode=0, y_hist=None, defer=None):
    """launches the forward kernels; returns (sums[8] fp64 device tensor, cot_u, cot_v).
    vcache/vmode: 0 none, 1 evaluate the v net and fill `vcache`, 2 reuse `vcache` (same sample, same theta_v).
    defer: a dict -- when the boundary pass runs on the second stream, do NOT join it here but hand the stream back in
    defer["side"]; the caller joins (WeakLoss, under CUDA-graph capture only)."""
    dev = theta_u.device
    dims = spec.c()
    st = _stream(dev)
    N, L = batch.N, batch.L
    wsb = lib.workspace_bytes(dims, max(N, batch.Nb if with_boundary else 1), max(L, batch.Lb if with_boundary else 1))
    ws = _WS.get(dev, wsb)
    sums = torch.zeros(_lib.NSUMS, dtype=torch.float64, device=dev)
    cot_u = torch.empty(N * L, dtype=torch.float32, device=dev)
    cot_v = torch.empty(N * L, dtype=torch.float32, device=dev)
    cdom, ccoef, pts = dom.c(), coef.c(), batch.points()
    # The boundary pass (forward + reverse sweep of its own paths, own slots of `sums`, own gradient buffer) does not depend
    # on the interior forward, so it runs beside it on a second stream with its own workspace.  Small samples (the shipped
    # N = 4000: 125 warps on 148 SMs) leave most of the GPU idle and every kernel is one dependent chain per path: the two
    # passes run side by side; with full waves the tail of one kernel overlaps the head of the next (a few per cent).
    # Under CUDA-graph capture the fork / join become parallel branches of the graph.
    side, ws_b, st_b = None, ws, st
    if with_boundary and dev.type == "cuda" and _concurrent_boundary(N, batch.Nb, wsb):
        side = _side_stream(dev)
        ws_b = _WS_SIDE.get(dev, wsb)
        side.wait_stream(torch.cuda.current_stream(dev))                # fork: inputs, `sums` and the gradient buffer are ready
        st_b = C.c_void_p(side.cuda_stream)
    _call(lib, "xw_interior_forward", dev, C.byref(dims), C.byref(cdom), C.byref(ccoef), _ptr(theta_u), _ptr(theta_v),
             C.c_void_p(batch.x.data_ptr() + 4 * batch.x_off), batch.x_sn, _ptr(batch.times), L, C.byref(pts),
             _ptr(batch.h), _ptr(batch.grad_h), _ptr(batch.f), N, _ptr(sums), _ptr(cot_u), _ptr(cot_v), None,
             _ptr(ws), ws.numel(), st, _ptr(batch.s0), _ptr(vcache) if vmode else None, int(vmode),
          _ptr(y_hist), vcache.numel() if (vmode and vcache is not None) else 0, y_hist.numel() if y_hist is not None else 0)
    if with_boundary:
        gscale = float(alpha) / (batch.Nb_glob * batch.Lb)
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):   # (_call's timing events follow)
            _call(lib, "xw_boundary_u", dev, C.byref(dims), _ptr(theta_u), C.c_void_p(batch.xb.data_ptr() + 4 * batch.xb_off),
                     batch.xb_sn, _ptr(batch.times_b), batch.Lb, _ptr(batch.sb), _ptr(batch.g), batch.Nb, gscale,
                     _ptr(sums), _ptr(boundary_grad), 0, _ptr(ws_b), ws_b.numel(), st_b)
        if side is not None and defer is not None:
            defer["side"] = side                                      # the caller joins ...
            defer["keep"] = ws_b      # ... and keeps the side branch's workspace alive until then: the allocator would hand a
            #                           freed block to the next allocation on the main stream while the side stream still uses it
        elif side is not None:
            torch.cuda.current_stream(dev).wait_stream(side)          # join: everything after this sees both passes
    return sums, cot_u, cot_v


def loss_from_sums(sums, batch, V, alpha):
    """device-side fp64 scalars (no host sync): I, S, init, bdry, int (src/loss.py:64-96)"""
    N, L = batch.N_glob, batch.L
    I = (V / N) * sums[_lib.SUM_S1] - (V / (N * L)) * (sums[_lib.SUM_S2] - sums[_lib.SUM_S3])
    S = V * sums[_lib.SUM_VV] / (N * L)
    init = sums[_lib.SUM_INIT] / N
    bdry = sums[_lib.SUM_BDRY] / (batch.Nb_glob * batch.Lb) if batch.Nb_glob else torch.zeros_like(I)
    integ = torch.log(I * I) - torch.log(S)
    return I, S, init, bdry, integ


class WeakLoss(torch.autograd.Function):
    """loss_u (phase 'u') or loss_v (phase 'v') as a differentiable function of the phase's own
    parameters.  Reproduces the reference's EFFECTIVE gradients, including the
    side effects of the two helper backward calls at src/loss.py:55,60 (`side_effect=True`)."""

    @staticmethod
    def forward(ctx, phase, lib, spec, dom, coef, alpha, batch, group, side_effect, vcache, vmode, sink, nu_params, *params):
        pu, pv = params[:nu_params], params[nu_params:]
        theta_u, theta_v = flatten_params(pu), flatten_params(pv)
        _check_dev(theta_u, "parameters", lib)
        dev = theta_u.device
        gb = torch.zeros(theta_u.numel(), dtype=torch.float32, device=dev) if phase == "u" else None
        yh = None
        if phase == "u":       # XNODE state history for the interior backward (saves its forward sweep)
            dims_ = spec.c()
            yh = torch.empty(int(lib.cdll.xw_yhist_floats(C.byref(dims_), batch.N, batch.L)), dtype=torch.float32, device=dev)
        # Under CUDA-graph capture on one rank the boundary pass (second stream, small samples: forward_sums) is joined
        # only AFTER the interior backward has been launched: the interior's cotangent coefficients k[] need the interior
        # sums alone, so the graph gets two branches -- interior forward -> backward | boundary pass -> loss value -- that
        # meet in front of the optimiser.  (Eagerly the loss tensor must be complete on the caller's stream when forward
        # returns, so the join stays in forward_sums; with several ranks the all-reduce of the sums needs both passes.)
        defer = {} if (phase == "u" and dev.type == "cuda" and DEFER_BOUNDARY_JOIN and not _is_distributed(group)
                       and any(ctx.needs_input_grad)                 # (no backward, no join: a capture would end unjoined)
                       and torch.cuda.is_current_stream_capturing()) else None
        sums, cot_u, cot_v = forward_sums(lib, spec, dom, coef, theta_u, theta_v, batch, phase == "u", alpha, gb,
                                          vcache, vmode, yh, defer)
        side = defer.get("side") if defer else None
        ctx.pending_side = side
        # (same for `sums`: the side branch adds the boundary sum to it and reads it for the loss value after forward returned)
        ctx.pending_keep = (defer.get("keep"), sums) if side is not None else None
        _allreduce(sums, group)
        ctx.phase, ctx.lib, ctx.spec, ctx.dom, ctx.batch, ctx.group = phase, lib, spec, dom, batch, group
        ctx.nu_params, ctx.side, ctx.sink = nu_params, 1.0 if side_effect else 0.0, sink
        ctx.meta = [(tuple(p.shape), p.dtype) for p in params]
        # loss, I, S, init, bdry and the backward's coefficients k[3] from the sums: one launch (xw_loss_scalars) instead
        # of ~22 one-element torch kernels; `loss_from_sums` below is the same arithmetic in torch (tests compare them)
        sargs = (float(dom.V), float(batch.N_glob), float(batch.L))
        nb_lb = (float(batch.Nb_glob if phase == "u" else 0), float(max(batch.Lb, 1)))
        if side is None:
            sc = torch.empty(8, dtype=torch.float64, device=dev)
            _call(lib, "xw_loss_scalars", dev, _ptr(sums), 0 if phase == "u" else 1, *sargs, *nb_lb, float(alpha), ctx.side,
                  _ptr(sc), _stream(dev))
            k = sc[5:8]
        else:
            main = torch.cuda.current_stream(dev)
            sck = torch.empty(8, dtype=torch.float64, device=dev)          # k[] (and I, S, init) from the interior sums:
            _call(lib, "xw_loss_scalars", dev, _ptr(sums), 0, *sargs, 0.0, 1.0, float(alpha), ctx.side, _ptr(sck), _stream(dev))
            k = sck[5:8]                                                    # nb = 0: the boundary slot is not read
            side.wait_stream(main)                                          # interior sums are final for the side branch
            with torch.cuda.stream(side):
                sc = torch.empty(8, dtype=torch.float64, device=dev)
                _call(lib, "xw_loss_scalars", dev, _ptr(sums), 0, *sargs, *nb_lb, float(alpha), ctx.side, _ptr(sc), _stream(dev))
                out = sc[0].clone()
        if phase == "u":
            ctx.save_for_backward(theta_u, cot_u, k, gb, yh)
        else:
            ctx.save_for_backward(theta_v, cot_v, k)
        ctx.components = dict(I=sc[1], S=sc[2], init=sc[3], bdry=sc[4])
        return sc[0].clone() if side is None else out

    @staticmethod
    def backward(ctx, go):
        lib, spec, batch = ctx.lib, ctx.spec, ctx.batch
        dims = spec.c()
        nup = ctx.nu_params
        none = (None,) * 13
        sink = ctx.sink          # flat fp32 gradient buffer of a FusedAdam: the gradient goes there, not into p.grad
        if ctx.phase == "u":
            theta_u, cot_u, k, gb, yh = ctx.saved_tensors
            dev = theta_u.device
            st = _stream(dev)
            ws = _WS.get(dev, lib.workspace_bytes(dims, batch.N, batch.L))
            ks = k.clone()
            ks[0:2] *= go.to(ks.dtype)
            side = getattr(ctx, "pending_side", None)
            if side is None:
                grad = torch.mul(gb, go.to(gb.dtype), out=sink) if sink is not None else (gb * go.to(gb.dtype)).contiguous()
            else:                    # the boundary pass may still be running: its gradient is added after the join below
                grad = sink if sink is not None else torch.empty_like(gb)
            _call(lib, "xw_interior_backward_u", dev, C.byref(dims), _ptr(theta_u),
                     C.c_void_p(batch.x.data_ptr() + 4 * batch.x_off), batch.x_sn, _ptr(batch.times), batch.L,
                     _ptr(batch.h), _ptr(cot_u), batch.N, _ptr(ks), _ptr(grad), 1 if side is None else 0, _ptr(ws), ws.numel(), st,
                     _ptr(batch.s0), _ptr(yh))
            if side is not None:
                torch.cuda.current_stream(dev).wait_stream(side)      # join: boundary gradient and loss value are final
                grad.addcmul_(gb, go.to(gb.dtype).expand_as(gb))
                ctx.pending_side = ctx.pending_keep = None
            _allreduce(grad, ctx.group)
            if sink is not None:
                return none + (None,) * len(ctx.meta)
            gl = unflatten_like(grad, ctx.meta[:nup])
            return none + tuple(gl) + (None,) * (len(ctx.meta) - nup)
        theta_v, cot_v, k = ctx.saved_tensors
        dev = theta_v.device
        st = _stream(dev)
        ws = _WS.get(dev, lib.workspace_bytes(dims, batch.N, batch.L))
        ks = k.clone()
        ks[0:2] *= go.to(ks.dtype)
        grad = sink if sink is not None else torch.empty(theta_v.numel(), dtype=torch.float32, device=dev)
        cdom, pts = ctx.dom.c(), batch.points()
        _call(lib, "xw_interior_backward_v", dev, C.byref(dims), C.byref(cdom), _ptr(theta_v), C.byref(pts), _ptr(cot_v),
                 batch.N, batch.L, _ptr(ks), _ptr(grad), 0, _ptr(ws), ws.numel(), st)
        _allreduce(grad, ctx.group)
        if sink is not None:
            return none + (None,) * len(ctx.meta)
        gl = unflatten_like(grad, ctx.meta[nup:])
        return none + (None,) * nup + tuple(gl)


def weak_loss(phase, spec, dom, coef, alpha, batch, u_params, v_params, group=None, side_effect=True, lib=None,
              vcache=None, vmode=0, grad_sink=None):
    """public functional entry: returns the scalar loss tensor (fp64) with autograd wired to the
    phase's own parameters.  `loss.components` is attached for logging (I, S, init, bdry)."""
    assert phase in ("u", "v")
    lib = lib or _lib.get()
    u_params, v_params = list(u_params), list(v_params)
    out = WeakLoss.apply(phase, lib, spec, dom, coef, alpha, batch, group, side_effect, vcache, vmode, grad_sink,
                         len(u_params), *(u_params + v_params))
    out.components = getattr(out.grad_fn, "components", None)
    return out


def xnode_eval(spec, u_params, x_base, x_off, x_sn, times, s0, n, lib=None):
    """u[n, L] forward only (fp32)"""
    lib = lib or _lib.get()
    theta_u = flatten_params(u_params)
    _check_dev(theta_u, "parameters", lib)
    dims = spec.c()
    L = times.numel()
    out = torch.empty(n, L, dtype=torch.float32, device=theta_u.device)
    _call(lib, "xw_xnode_eval", theta_u.device, C.byref(dims), _ptr(theta_u), C.c_void_p(x_base.data_ptr() + 4 * x_off), x_sn,
             _ptr(times), L, _ptr(s0), n, _ptr(out), _stream(theta_u.device))
    return out


def vnet_eval(spec, v_params, XV, lib=None):
    """v[N, L] forward only (fp32) for XV [N, L, C]"""
    lib = lib or _lib.get()
    theta_v = flatten_params(v_params)
    _check_dev(theta_v, "parameters", lib)
    XVf = as_f32(XV)
    N, L, Cc = XVf.shape
    dims = spec.c()
    pts = _lib.Points(XVf.data_ptr(), L * Cc, Cc, XVf.data_ptr() + 4, L * Cc, Cc)
    out = torch.empty(N, L, dtype=torch.float32, device=theta_v.device)
    _call(lib, "xw_vnet_eval", theta_v.device, C.byref(dims), _ptr(theta_v), C.byref(pts), N, L, _ptr(out), _stream(theta_v.device))
    return out
