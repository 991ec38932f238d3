"""`func_eval` and `NODE_WAN_solver` with the reference's signatures
(/root/reference/src/training.py:13-43, :54-187).  The min-max loop, Adam x2, init, logging and the
stop criterion stay plain Python/PyTorch (they are the CALLER of the hot path); the
per-batch work (net forwards, coefficient evaluation, weak-form loss, backward) goes through the
fused sm_100a kernels."""
import itertools
import json
import os
import time

import torch

from . import dataset as _dataset
from .aux import L_norm
from .dataset import Comb_loader
from .loss import CoefA, CoefB, CoefC, loss
from .model import NeuralODE, discriminator, init_weights

_COEF_CACHE = {}
_PROBE_PATHS = 64


def _constant_value(vals):
    """vals: tensor over probe points -> python float if all equal, else None"""
    v0 = vals.reshape(-1)[0]
    return float(v0) if bool(torch.all(vals == v0)) else None


def classify_coefficients(X, setup, func_a, func_b, func_c):
    """the STRUCTURE of the user callables a_ij(X), b_i(X), c(X,u) (SURVEY.md 7 'User callables'), probed on a spread
    of sample paths once per (callables, dim) and cached:
      a: ("identity",) | ("const", A[d,d]) | ("per_path",)      b: ("zero",) | ("const", B[d]) | ("per_path",)
      c: ("affine", c0, c1) | ("callable",)
    Constant / affine coefficients (every shipped problem) cost nothing per sample; the other forms are evaluated per
    sample by `func_eval` (a, b on time-row 0 of every path) or per kernel pass by the loss (c)."""
    d = setup['dim']
    # keyed on the callables THEMSELVES (strong references): an id() can be recycled by another problem's closures
    # after garbage collection and would silently hand it this problem's structure
    key = (func_a, func_b, func_c, d)
    if key in _COEF_CACHE:
        return _COEF_CACHE[key]
    n_all = X.shape[0]
    if n_all > _PROBE_PATHS:         # a spread of paths, not just the first few
        sel = torch.linspace(0, n_all - 1, _PROBE_PATHS).long()
        Xs = X.take(sel).dense() if hasattr(X, "dense") else X[sel.to(X.device)].detach()
    else:
        Xs = X.dense() if hasattr(X, "dense") else X.detach()
    A = torch.empty(d, d, dtype=torch.float64)
    a_kind = None
    for i, j in itertools.product(range(d), repeat=2):
        v = _constant_value(func_a(Xs, i, j))
        if v is None:
            a_kind = ("per_path",)
            break
        A[i, j] = v
    if a_kind is None:
        a_kind = ("identity",) if bool(torch.equal(A, torch.eye(d, dtype=torch.float64))) else ("const", A.float())
    bv = [_constant_value(func_b(Xs, i)) for i in range(d)]
    if any(v is None for v in bv):
        b_kind = ("per_path",)
    else:
        B = torch.tensor(bv, dtype=torch.float64)
        b_kind = ("zero",) if bool(torch.all(B == 0)) else ("const", B.float())
    shape = (Xs.shape[0], Xs.shape[1], 1)
    u0 = torch.zeros(shape, dtype=torch.float64, device=Xs.device)
    c_at0, c_at1, c_at2 = (func_c(Xs, u0 + k) for k in (0.0, 1.0, 2.0))
    c0, c1 = _constant_value(c_at0), _constant_value(c_at1 - c_at0)
    affine = c0 is not None and c1 is not None and bool(torch.allclose(c_at2, c_at0 + 2.0 * (c_at1 - c_at0), rtol=1e-12, atol=1e-12))
    c_kind = ("affine", c0, c1) if affine else ("callable",)
    out = (a_kind, b_kind, c_kind)
    _COEF_CACHE[key] = out
    return out


def _row0(X):
    """time-row 0 of every path as a dense [N, 1, C] tensor (what a_ij / b_i are evaluated on)"""
    return X[:, 0, :].detach().unsqueeze(1)


def func_eval(X: torch.Tensor, BX: torch.Tensor, setup: dict, y_output_u, func_a, func_b, func_c, func_h, func_f,
              func_g):
    """h, f, g evaluated on the sample by the user's callables (as the reference does); a, b, c as structure
    (`CoefA / CoefB / CoefC`) instead of the reference's dense [d,d,N,L] / [d,N,L] tensors (src/training.py:32-41):
    constants where the callables are constant, per-path values on time-row 0 where a or b vary over the sample, the
    callable itself for a c that depends on X or is not affine in u.  `y_output_u` is accepted for signature parity."""
    h = func_h(X[:, 0, :])
    f = func_f(X)
    g = func_g(BX)
    a_kind, b_kind, c_kind = classify_coefficients(X, setup, func_a, func_b, func_c)
    d, n = setup['dim'], X.shape[0]
    if a_kind[0] == "per_path":
        X0 = _row0(X)
        a = CoefA(per_path=torch.stack([func_a(X0, i, j).reshape(n) for i, j in itertools.product(range(d), repeat=2)],
                                       dim=1).reshape(n, d, d).float().contiguous().to(X.device))
    else:
        a = CoefA(a_kind[1] if a_kind[0] == "const" else None)
    if b_kind[0] == "per_path":
        X0 = _row0(X)
        b = CoefB(per_path=torch.stack([func_b(X0, i).reshape(n) for i in range(d)], dim=1).float().contiguous().to(X.device))
    else:
        b = CoefB(b_kind[1] if b_kind[0] == "const" else None)
    c = CoefC(c_kind[1], c_kind[2]) if c_kind[0] == "affine" else CoefC(func=func_c)
    return h.to(X.device), f.to(X.device), g.to(X.device), a, b, c


class NODE_WAN_solver:
    """weak adversarial training loop; constructor arguments as the reference (src/training.py:65-66)"""

    def __init__(self, params: dict, func_a, func_b, func_c, func_h, func_f, func_g, device, path, stop=None,
                 func_u_sol=None, p: float = 1, log_json: bool = True, use_cuda_graph: bool = False,
                 sample_on_device: bool = False, collapsed_layout: bool = False, fused_optimizer=None):
        self.params = params
        self.func_a, self.func_b, self.func_c = func_a, func_b, func_c
        self.func_h, self.func_f, self.func_g = func_h, func_f, func_g
        self.device, self.path, self.stop, self.func_u_sol, self.p = device, path, stop, func_u_sol, p
        self.log_json = log_json
        self.use_cuda_graph = use_cuda_graph and torch.device(device).type == "cuda"
        self._graphs = None
        # extensions over the reference: draw the samples on the GPU (same distributions, different RNG
        # stream) and keep them in the collapsed layout (times[L] + x[N,d]); both default to the
        # reference behaviour (CPU sampling, repeated [N,L,C] tensors)
        self.sample_on_device, self.collapsed_layout = sample_on_device, collapsed_layout
        it = iter(params.items())                       # positional split, as the reference
        self.config = dict(itertools.islice(it, 13))
        self.setup = dict(itertools.islice(it, 7))
        self.iterations = dict(itertools.islice(it, 1))['iterations']
        dom = params['domain']
        self.domain = getattr(_dataset, dom) if isinstance(dom, str) else dom
        self.n1, self.n2 = self.config['n1'], self.config['n2']
        domain = self.new_domain()
        dev = torch.device(device)
        ids = [dev.index if dev.index is not None else torch.cuda.current_device()] if dev.type == "cuda" else None
        self.u_net = torch.nn.DataParallel(
            NeuralODE(self.config['u_hidden_dim'], 1, func_h, func_g, self.setup, self.config['u_hidden_hidden_dim'],
                      self.config['u_layers'], domain, self.config['solver'], self.config['min_steps'],
                      self.config['adjoint']), device_ids=ids).to(device)
        self.v_net = torch.nn.DataParallel(discriminator(self.config, self.setup), device_ids=ids).to(device)
        self.u_net.apply(init_weights)
        self.v_net.apply(init_weights)
        cap = self.use_cuda_graph        # step counters on the device so that Adam can live inside a CUDA graph
        # fused_optimizer (default: on with CUDA-graph replay): one flat fp64 buffer per net + a single-launch Adam that
        # consumes the kernels' flat fp32 gradient (optim.py); off = two torch.optim.Adam as the reference
        self.fused_optimizer = bool(self.use_cuda_graph if fused_optimizer is None else fused_optimizer)
        if self.fused_optimizer and self.config['u_layers'] > 1:
            from . import hotpath as _hp
            from .optim import FlatParameters, FusedAdam
            lib = _hp._lib.get()
            self.optimizer_u = FusedAdam(FlatParameters(self.u_net.module.kernel_parameters()), self.config['u_rate'], lib)
            self.optimizer_v = FusedAdam(FlatParameters(self.v_net.module.flat_parameters()), self.config['v_rate'], lib)
        else:
            self.fused_optimizer = False
            self.optimizer_u = torch.optim.Adam(self.u_net.parameters(), lr=self.config['u_rate'], capturable=cap)
            self.optimizer_v = torch.optim.Adam(self.v_net.parameters(), lr=self.config['v_rate'], capturable=cap)
        # one process per GPU: N_r / N_b are GLOBAL counts, every rank samples its own shard; identical
        # seeds give identical initial weights, all-reduced sums/gradients keep the replicas identical
        self.world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        self.rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        if self.world > 1:
            # replicas must START identical whatever the callers' seeds were (the reference sets none): rank 0's
            # initial weights win; all-reduced gradients then keep the Adam replicas identical
            with torch.no_grad():
                for q in list(self.u_net.parameters()) + list(self.v_net.parameters()):
                    torch.distributed.broadcast(q.data, src=0)
                if self.fused_optimizer:         # (.data writes do not bump version counters: refresh the fp32 copies)
                    for opt in (self.optimizer_u, self.optimizer_v):
                        opt.flat.f32.copy_(opt.flat.flat)
            # ... and the shards must DIFFER: with equal seeds every rank would draw the same paths.  Decorrelate the
            # sampling streams (torch CPU + CUDA generators and numpy) per rank, derived from rank 0's current state.
            base = torch.tensor([int(torch.initial_seed()) % (2 ** 31)], dtype=torch.int64, device=self.device)
            torch.distributed.broadcast(base, src=0)
            import numpy as _np
            torch.manual_seed(int(base.item()) + 7919 * (self.rank + 1))
            _np.random.seed((int(base.item()) + 7919 * (self.rank + 1)) % (2 ** 32))
        self._warm = 0                   # completed eager iterations (the first one warms up before graph capture)
        self.reuse_v = True              # cache the test-function values across the sub-steps of one iteration
        self._vc_buf, self._vc_key, self._theta_v_gen = None, None, 0
        self.keep_l2_history = True      # False: skip the per-iteration L2 evaluation on a fresh sample when nothing logs it
        self._coef = None                # (sample token, (h, f, g, a, b, c), hotpath.Batch) of the sample in flight
        self._load_gen = 0               # bumped whenever new data lands in the graphs' static buffers
        self.best_l = float('inf')
        self.av_l = 0
        self.history = {"loss_u": [], "loss_v": [], "L2": [], "time": []}

    @property
    def av_l(self):
        """loss of the last u sub-iteration (reference src/training.py:136); synchronises on first access"""
        if self._av_pending is not None:
            self._av_value = sum(v.item() for v in self._av_pending)
            self._av_pending = None
        return self._av_value

    @av_l.setter
    def av_l(self, value):
        self._av_pending, self._av_value = None, value

    def new_domain(self, **kw):
        s = self.setup
        if getattr(self, "sample_on_device", False):
            kw.setdefault("sample_device", self.device)
        if getattr(self, "collapsed_layout", False) and self.domain is _dataset.Hypercube:
            kw.setdefault("collapsed", True)       # (only the cube repeats ONE time grid for every path)
        dom = self.domain(s['shape_param'], s['dim'], s['T0'], s['T'], s['N_t'], **kw)
        if torch.device(self.device).type == "cuda" and hasattr(dom, "pin_host") and kw.get("sample_device") is None:
            dom.pin_host = True         # samples drawn on the host land in page-locked memory: their H2D copy is asynchronous
        if getattr(self, "world", 1) > 1:          # the time grid must be the same on every rank
            t = dom.times.to(self.device)
            torch.distributed.broadcast(t, src=0)
            dom.times = t.to(dom.times.device)
        return dom

    def local_counts(self):
        """(N_r, N_b) of this rank's shard"""
        w = getattr(self, "world", 1)
        if self.setup['N_r'] % w or self.setup['N_b'] % w:
            raise RuntimeError("N_r and N_b must be divisible by the number of ranks")
        return self.setup['N_r'] // w, self.setup['N_b'] // w

    def _vcache_plan(self, points):
        """(buffer, mode) for the next sub-step on `points`: the test-function values depend only on
        the sample and theta_v, both unchanged between the sub-steps of one outer iteration until a
        v-step runs (src/training.py:125-162), so only the first sub-step evaluates the v net."""
        if not self.reuse_v or isinstance(points.interioru, list):
            return None, 0
        key = (points.uid, self._theta_v_gen)
        self._ensure_vcache(points)
        return self._vc_buf, (2 if self._vc_key == key else 1)

    def _ensure_vcache(self, points):
        """(re)allocate the test-function cache for the batch about to run: the buffer is sized by (N, L) of a batch and
        a later, larger batch must not write past it (the C ABI checks the capacity as well)"""
        X = points.interioru
        N, L = X.shape[0], X.shape[1]
        um, vm = self.u_net.module, self.v_net.module
        from . import hotpath as _hp
        dims = um.spec(vm).c()
        import ctypes as _C
        need = int(_hp._lib.get().cdll.xw_vcache_floats(_C.byref(dims), N, L))
        if self._vc_buf is None or self._vc_buf.numel() < need or self._vc_buf.device != torch.device(self.device):
            self._vc_buf = torch.empty(need, dtype=torch.float32, device=self.device)
            self._vc_key = None            # cached values are gone
            self._graphs = None            # captured graphs hold the old buffer
        self._vc_shape = (N, L)

    def _vcache_commit(self, phase, points):
        if self.reuse_v and not isinstance(points.interioru, list):
            if phase == "v":
                self._theta_v_gen += 1          # theta_v just changed: cached values are stale
            else:
                self._vc_key = (points.uid, self._theta_v_gen)

    def _step(self, phase, domain, batch, vplan=(None, 0), token=None):
        """one sub-step on one batch.  `token` identifies the SAMPLE: the coefficient values on it (h, f, g, grad h, the
        boundary's initial scalars -- about 60 small torch launches through the user callables) do not depend on the
        parameters, so the first sub-step on a sample evaluates them and the following ones (same token) reuse them."""
        datau, datav, bdata = batch
        prediction_v = self.v_net(datav)
        prediction_u = self.u_net(datau)
        cached = self._coef if (token is not None and self._coef is not None and self._coef[0] == token) else None
        persist = self._graphs.get("persist") if (self._graphs is not None and token is not None and token[0] == "graph") else None
        if cached is not None:
            h, f, g, a, b, c = cached[1]
        else:
            h, f, g, a, b, c = func_eval(datau.detach(), bdata.detach(), self.setup, prediction_u, self.func_a,
                                         self.func_b, self.func_c, self.func_h, self.func_f, self.func_g)
            if persist is not None:
                # captured "fresh" graph: the sample-dependent values go to their PERSISTENT home (allocated eagerly in
                # _capture), which every other captured graph reads -- whichever fresh graph ran last filled it
                (ph, pf, pg, _, _, _), pb = persist
                tmp = loss(self.config['alpha'], a, b, c, h, f, g, self.setup, domain, self.device)
                nb = tmp._batch(self.u_net.module, datau, datav, bdata)
                ph.copy_(h); pf.copy_(f); pg.copy_(g)
                from . import hotpath as _hp
                _hp.copy_batch_values_(pb, nb)
                h, f, g = ph, pf, pg
                cached = (token, (h, f, g, a, b, c), pb)
                self._coef = cached
        Loss = loss(self.config['alpha'], a, b, c, h, f, g, self.setup, domain, self.device)
        if cached is not None and (phase == "v" or cached[2].Nb > 0):
            Loss.batch_cache = cached[2]
        if self.world > 1:
            Loss.N_glob, Loss.Nb_glob = datau.shape[0] * self.world, bdata.shape[0] * self.world
        vbuf, vmode = vplan
        Loss.vcache = (vbuf, vmode)
        if self.fused_optimizer:
            Loss.grad_sink = (self.optimizer_u if phase == "u" else self.optimizer_v).grad32
        Loss._u_module = self.u_net.module
        if phase == "u":
            val = Loss.u(prediction_u, prediction_v, self.u_net, datau, datav, bdata)
            val.backward()
            self.optimizer_u.step()
        else:
            val = Loss.v(prediction_u, prediction_v, datau, datav)
            val.backward()
            self.optimizer_v.step()
        if token is not None and Loss.last_batch is not None and (cached is None or Loss.last_batch.Nb > cached[2].Nb):
            self._coef = (token, (h, f, g, a, b, c), Loss.last_batch)
        # hand back a detached scalar: keeping the autograd graph alive would pin the parameters'
        # AccumulateGrad nodes to the stream of this call (and break later CUDA-graph capture)
        out = val.detach()
        out.components = val.components
        return out

    # ------------------------------------------------------------------ CUDA-graph replay of a sub-step
    def _capture(self, domain, batch):
        """record one u-step and one v-step (coefficient evaluation + fused loss + backward + Adam)
        on static copies of the sample; later samples are copied into the static buffers and the
        graphs replayed: one launch per sub-step instead of ~150 (matters at the shipped N=4000,
        where a sub-step is host-bound)."""
        static = tuple(t.clone() for t in batch)
        for t, src in zip(static, batch):
            t._xw_start = getattr(src, "_xw_start", None)
        # persistent home of the sample-dependent values (coefficient values + packed batch) on the static copies,
        # allocated OUTSIDE any graph: "fresh" graphs write into it, the others read it (see _step)
        h, f, g, a, b, c = func_eval(static[0].detach(), static[2].detach(), self.setup, None, self.func_a, self.func_b,
                                     self.func_c, self.func_h, self.func_f, self.func_g)
        if a.per_path is not None or b.per_path is not None or c.func is not None:
            # coefficients that vary over the sample / a callable c are re-evaluated by PyTorch per sample / per pass
            # (host-side shape logic, autograd through the user's callable): those steps run eagerly
            self.use_cuda_graph = False
            return
        tmp = loss(self.config['alpha'], a, b, c, h, f, g, self.setup, domain, self.device)
        if self.world > 1:
            tmp.N_glob, tmp.Nb_glob = static[0].shape[0] * self.world, static[2].shape[0] * self.world
        pb = tmp._batch(self.u_net.module, static[0], static[1], static[2])
        # one persistent scratch for all graphs of this sample size (and a second one for the boundary branch), allocated
        # here, outside the captures (hotpath._Workspace.reserve_for_graphs); the references keep them alive with the graphs
        from . import hotpath as _hp
        from . import _lib as _xl
        dev = torch.device(self.device)
        wsb = _xl.get().workspace_bytes(self.u_net.module.spec(self.v_net.module).c(), max(pb.N, pb.Nb, 1), max(pb.L, pb.Lb, 1))
        scratch = [_hp._WS.reserve_for_graphs(dev, wsb)]
        if wsb <= _hp.CONCURRENT_BOUNDARY_MAX_WORKSPACE:
            scratch.append(_hp._WS_SIDE.reserve_for_graphs(dev, wsb))
        self._graphs = dict(static=static, graphs={}, outs={}, domain=domain, persist=((h, f, g, a, b, c), pb), scratch=scratch)

    def _graph_fits(self, batch):
        """the captured graphs are bound to static copies of one layout ([N,L,C] tensors or CollapsedPaths) and size"""
        for dst, src in zip(self._graphs["static"], batch):
            if hasattr(dst, "times") != hasattr(src, "times") or tuple(dst.shape) != tuple(src.shape):
                return False
        return True

    def _graph_for(self, phase, vmode, fresh):
        # `fresh`: first sub-step after new data was copied into the static buffers -> this graph contains the
        # coefficient evaluation; the graphs of the following sub-steps read its results
        key = (phase, vmode, fresh)
        if key not in self._graphs["graphs"]:
            g = torch.cuda.CUDAGraph()
            opt = self.optimizer_u if phase == "u" else self.optimizer_v
            # capture on a side stream with the low-level API: `with torch.cuda.graph(g)` runs gc.collect() and
            # torch.cuda.empty_cache() on entry (~0.15 s per capture, 4 captures per solver: a third of a short training run)
            dev = torch.device(self.device)
            side = self._graphs.setdefault("stream", torch.cuda.Stream(dev))
            cur = torch.cuda.current_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                g.capture_begin()
                try:
                    opt.zero_grad(set_to_none=True)
                    if fresh:
                        self._coef = None
                    self._graphs["outs"][key] = self._step(phase, self._graphs["domain"], self._graphs["static"],
                                                           (self._vc_buf, vmode), token=("graph", id(self._graphs)))
                finally:
                    g.capture_end()
            cur.wait_stream(side)
            self._graphs["graphs"][key] = g
        return key

    def _graph_step(self, phase, batch, vmode):
        st = self._graphs["static"]
        fresh = batch is not self._graphs.get("loaded")
        if fresh:
            for dst, src in zip(st, batch):
                if hasattr(dst, "times"):
                    dst.times.copy_(src.times, non_blocking=True)
                    dst.x.copy_(src.x, non_blocking=True)
                else:
                    dst.copy_(src, non_blocking=True)
            self._graphs["loaded"] = batch
        key = self._graph_for(phase, vmode, fresh)
        self._graphs["graphs"][key].replay()
        self._last_losses = [self._graphs["outs"][key]]
        return self._graphs["outs"][key]

    def sub_step(self, phase, domain, points):
        """one u- or v- sub-iteration over every batch of `points` (zero_grad once, a step per batch,
        as the reference: src/training.py:127-138 / :152-162); returns the last loss tensor"""
        single = not isinstance(points.interioru, list)
        vplan = self._vcache_plan(points)
        if self.use_cuda_graph and single and self._graphs is not None and not self._graph_fits(points[0]):
            self._graphs = None                      # another layout / size than the captured one: capture again later
            self._coef = None
        if self.use_cuda_graph and single and self._graphs is not None:
            val = self._graph_step(phase, points[0], vplan[1])
            self._vcache_commit(phase, points)
            return val
        (self.optimizer_u if phase == "u" else self.optimizer_v).zero_grad()
        val = None
        self._last_losses = []
        for k, batch in enumerate(points):
            val = self._step(phase, domain, batch, vplan, token=(points.uid, k))
            self._last_losses.append(val)
        self._vcache_commit(phase, points)
        if self.use_cuda_graph and single and self._graphs is None and self._warm >= 1:
            torch.cuda.synchronize()
            self._capture(domain, points[0])
        return val

    def train_iteration(self, domain, points):
        """the hot part of one outer iteration: n1 u-steps then n2 v-steps on the same sample
        (reference src/training.py:125-162 without logging / stop / checkpoint).  Returns the last
        (loss_u, loss_v) tensors; nothing here synchronises the host."""
        loss_u = loss_v = None
        for _ in range(self.n1):
            loss_u = self.sub_step("u", domain, points)
        for _ in range(self.n2):
            loss_v = self.sub_step("v", domain, points)
        self._warm += 1
        return loss_u, loss_v

    def _agree(self, flag):
        """the stop decision must be the SAME on every rank (each evaluates it on its own shard): a rank that returned
        while the others entered the next all-reduce would hang the job.  Stop when ANY rank's criterion fires."""
        if self.world <= 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=self.device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return bool(t.item())

    def train(self, report: bool = False, report_it: int = 10, show_plt: bool = False, max_seconds=None):
        past_losses = []
        t_start = time.time()
        times = [t_start]
        loss_u = loss_v = None
        ahead = None
        for k in range(self.iterations):
            n_r, n_b = self.local_counts()
            if ahead is None:
                domain = self.new_domain()
                points = Comb_loader(n_r, n_b, domain, self.device)
            else:
                domain, points = ahead
            for i in range(self.n1):
                loss_u = self.sub_step("u", domain, points)
                # `av_l` (sum of the batch losses; all-reduced sums: identical on every rank) is read from the device the
                # first time someone looks at it: stop() enqueues its evaluation behind the sub-step BEFORE the host
                # waits for the loss value, so its host-side work runs while the GPU is still busy with the sub-step
                self._av_pending = list(self._last_losses)
                stopped = self.stop is not None and self._agree(self.stop(self, points.interioru, domain))
                past_losses.append(self.av_l)
                if self.log_json and self.rank == 0:
                    with open('losses_NODE_' + str(self.setup['dim']) + '.json', 'w') as fh:
                        json.dump(past_losses, fh)
                if stopped:
                    if self.rank == 0:
                        torch.save(self.u_net.state_dict(), os.path.join(self.path, 'best_model_weights_NODE.pth'))
                        print('Stopping Criterion Reached')
                    self.history["stopped_at_subiter"] = len(past_losses)
                    return self.history
                if self.av_l < self.best_l:
                    if self.log_json and self.rank == 0:
                        torch.save(self.u_net.state_dict(), 'best_model_weights_NODE.pth')
                    self.best_l = self.av_l
            for j in range(self.n2):
                loss_v = self.sub_step("v", domain, points)
            self._warm += 1
            # Nothing above waited for the v-steps: the host draws the samples it needs next WHILE the GPU runs them --
            # the logging sample of this iteration, then the domain and sample of the next one, in the order the
            # reference draws them (src/training.py:114-115, 165), so the RNG stream is the reference's.
            L2 = fresh = None
            if self.func_u_sol is not None and (self.log_json or report or self.keep_l2_history):
                # (reference src/training.py:165-170: a fresh sample and one more forward, for logging only)
                fresh = Comb_loader(n_r, n_b, domain, self.device)
            ahead = None
            if k + 1 < self.iterations:
                domain_next = self.new_domain()
                ahead = (domain_next, Comb_loader(n_r, n_b, domain_next, self.device).prefetch())
            if fresh is not None:
                L2 = L_norm(fresh.interioru, self.u_net, self.p, self.func_u_sol, domain.V(), n_r)
                if self.world > 1:         # shards of equal size: the global L^p error is the p-mean of the shard errors
                    Lp = L2.detach().double() ** self.p
                    torch.distributed.all_reduce(Lp)
                    L2 = (Lp / self.world) ** (1.0 / self.p)
                L2 = L2.item()
                if self.log_json and self.rank == 0:
                    with open('L2_NODE_' + str(self.setup['dim']) + '.json', 'w') as fh:
                        json.dump([L2], fh)
            times.append(time.time())
            if self.log_json and self.rank == 0:
                with open('Time_NODE_' + str(self.setup['dim']) + '.json', 'w') as fh:
                    json.dump(times, fh)
            self.history["loss_u"].append(self.av_l)
            self.history["loss_v"].append(loss_v.item() if loss_v is not None else None)
            self.history["L2"].append(L2)
            self.history["time"].append(times[-1] - t_start)
            if report and k % report_it == 0 and self.rank == 0:
                print('iteration: ' + str(k), 'Loss u: ' + str(self.av_l), 'Loss v: ' + str(self.history["loss_v"][-1]))
                if L2 is not None:
                    print('L^2 norm error: ' + str(L2))
            if max_seconds is not None and self._agree(times[-1] - t_start > max_seconds):
                break
        return self.history
