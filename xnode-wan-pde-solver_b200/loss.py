"""Drop-in `loss` class (reference /root/reference/src/loss.py:12-96).

Same constructor and method signatures; `u()` / `v()` return a float64 scalar tensor whose
`.backward()` fills `.grad` of u_net's (resp. v_net's) parameters with exactly what the reference's
optimisers see.  The work is done by the fused sm_100a kernels; the
arguments `a, b, c` are the structured coefficients produced by training.func_eval (CoefA / CoefB /
CoefC) instead of the reference's dense [d,d,N,L] tensors.
"""
import torch

from . import hotpath
from .model import LazyPrediction, unwrap
from .paths import CollapsedPaths


class CoefA:
    """a_ij: identity (matrix=None), a constant [d,d] matrix, or -- for a_ij(X) that varies over the sample -- its values
    on time-row 0 of every path, per_path[N,d,d] (the only ones that enter the loss: src/loss.py:66-68 multiplies a with
    du, which lives on row 0)"""
    def __init__(self, matrix=None, per_path=None):
        self.matrix, self.per_path = matrix, per_path


class CoefB:
    """b_i: zero (vector=None), a constant [d] vector, or per_path[N,d] (values on time-row 0, as for a)"""
    def __init__(self, vector=None, per_path=None):
        self.vector, self.per_path = vector, per_path


class CoefC:
    """c(X,u) = c0 + c1*u, or any callable func(X, u) (src/training.py:30): then A(u) = c(X,u) u and dA/du are
    evaluated per point with the callable (PyTorch autograd, elementwise) at the current u before every kernel pass"""
    def __init__(self, c0=0.0, c1=0.0, func=None):
        self.c0, self.c1, self.func = float(c0), float(c1), func


def domain_spec(domain):
    """reference domain object (src/dataset.py) -> DomainSpec for the kernels"""
    name = type(domain).__name__
    V = float(domain.V())
    if name == "Hypercube":
        return hotpath.DomainSpec("cube", (domain.bot, domain.top, 0.0), V)
    if name == "NSphere_TCone":
        return hotpath.DomainSpec("cone", (domain.r, 0.0, 0.0), V)
    if name == "NSphere_THourglass":
        return hotpath.DomainSpec("hourglass", (domain.r, domain.T0, domain.T), V)
    raise NotImplementedError("domain %s: only Hypercube, NSphere_TCone and NSphere_THourglass have an in-kernel "
                              "func_w / grad func_w" % name)


class loss:
    def __init__(self, alpha: float, a, b, c, h: torch.Tensor, f: torch.Tensor, g: torch.Tensor, setup: dict,
                 domain, device):
        self.T, self.T0 = setup['T'], setup['T0']
        self.alpha = alpha
        self.a, self.b, self.c = a, b, c
        self.h, self.f, self.g = h, f, g
        self.setup = setup
        self.domain = domain
        self.func_w = domain.func_w
        self.V = domain.V()
        self.device = device
        self.group = None            # torch.distributed process group (None = default group if initialised)
        self.N_glob = None           # global path counts when the batch is one rank's shard
        self.Nb_glob = None
        self.side_effect = True      # reproduce the reference's helper-backward side effects
        self.vcache = None           # (buffer or None, mode): test-function cache, managed by NODE_WAN_solver
        self.grad_sink = None        # flat fp32 gradient buffer of an optim.FusedAdam (NODE_WAN_solver), else p.grad is filled
        self.batch_cache = None      # a hotpath.Batch built for this very sample by an earlier sub-step (NODE_WAN_solver)
        self.last_batch = None       # the Batch the last .u / .v call ran on
        for nm, obj, cls in (("a", a, CoefA), ("b", b, CoefB), ("c", c, CoefC)):
            if not isinstance(obj, cls):
                raise TypeError("coefficient %s must be a %s produced by func_eval (dense tensors of the reference "
                                "layout are not accepted: d=100 would need 3.4 TB)" % (nm, cls.__name__))

    # -------------------------------------------------------------------------------- internals
    def _nets(self, y_output_u, y_output_v):
        if not isinstance(y_output_u, LazyPrediction) or not isinstance(y_output_v, LazyPrediction):
            raise TypeError("loss.u / loss.v need the objects returned by u_net(X) and v_net(XV) of this package "
                            "(lazy predictions); got plain tensors")
        return y_output_u.net, y_output_v.net

    def _batch(self, u_net, X, XV, border):
        dev = X.device
        if isinstance(X, CollapsedPaths):
            b = hotpath.batch_from_collapsed(X.times, X.x, XV.x, border.x if border is not None else None,
                                             border.times if border is not None else None)
        else:
            b = hotpath.batch_from_reference_layout(X, XV, border)
        kind = u_net.start_kind(X)
        if kind == "pad":
            raise NotImplementedError("interior batch must start at T0 or on the boundary")
        if b.L == 1 and kind == "h":
            raise NotImplementedError("single-time-point interior groups (rank-2 shortcut, src/model.py:89-91)")
        b.h = hotpath.as_f32(self.h)
        x0 = X[:, 0, :].detach().clone().requires_grad_(True)        # [N, C]
        with torch.enable_grad():
            hv = u_net.h(x0) if kind == "h" else u_net.g(x0.unsqueeze(1)).reshape(-1)
            gh, = torch.autograd.grad(hv.sum(), x0, allow_unused=True)
        b.grad_h = hotpath.as_f32(gh[:, 1:]) if gh is not None else torch.zeros(b.N, b.d, device=dev)
        if kind == "g":
            b.s0 = hotpath.as_f32(hv)
        b.f = hotpath.as_f32(self.f)
        if border is not None:
            kb = u_net.start_kind(border)
            if kb == "pad":
                raise NotImplementedError("boundary batch must start at T0 or on the boundary")
            b.sb = hotpath.as_f32(u_net.initial_scalar(border.detach(), kb))
            b.g = hotpath.as_f32(self.g)
        if self.N_glob:
            b.N_glob = self.N_glob
        if self.Nb_glob and border is not None:
            b.Nb_glob = self.Nb_glob
        return b

    def _coef(self, dev, A_val=None, A_der=None):
        a_pp, b_pp = self.a.per_path is not None, self.b.per_path is not None
        a = self.a.per_path if a_pp else self.a.matrix
        bb = self.b.per_path if b_pp else self.b.vector
        return hotpath.CoefSpec(self.c.c0, self.c.c1,
                                hotpath.as_f32(a).to(dev) if a is not None else None,
                                hotpath.as_f32(bb).to(dev) if bb is not None else None, a_pp, b_pp, A_val, A_der)

    def _general_A(self, u_mod, X):
        """A(u) = c(X,u) u and dA/du per point for a callable c: u by the forward-only XNODE kernel (the numbers the
        interior forward pass computes), the callable and its u-derivative by PyTorch (elementwise on [N,L])"""
        with torch.no_grad():
            u = u_mod.evaluate(X)                                     # [N, L, 1] float64
        Xd = X.dense() if hasattr(X, "dense") else X
        u_t = u.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            cval = self.c.func(Xd.detach(), u_t)
            A = cval.reshape(u_t.shape[0], -1).to(u_t.dtype) * u_t.reshape(u_t.shape[0], -1)
            dA, = torch.autograd.grad(A.sum(), u_t)
        return hotpath.as_f32(A), hotpath.as_f32(dA.reshape(A.shape))

    def _single_time_group(self, phase, u_mod, v_mod, X, XV, border):
        """Interior group with ONE time point at T0 (first group of the sphere domains).  The reference
        returns rank-2 predictions there (src/model.py:89-91) and its loss then broadcasts [n]x[n,1]
        into [n, n] (src/loss.py:65-72,79,84): reproduced here as the equivalent sums, in PyTorch on
        the device -- there is no ODE and no path structure to accelerate (n points, a few kFLOP)."""
        d = self.setup['dim']
        lift, fin = u_mod.initial_layers, u_mod.final_linear
        Xl = X.detach().clone().requires_grad_(True)
        XVl = XV.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            U = fin(lift(u_mod.h(Xl[:, 0, :]).unsqueeze(1).double()))[:, 0]           # [n]
            Vv = v_mod.net(XVl.double())[:, 0, 0]                                       # [n]
            w = self.func_w(XVl)[:, 0]
            Phi = Vv * w
            du = torch.autograd.grad(U.sum(), Xl, retain_graph=True)[0][:, 0, 1:].to(X.dtype)
            dphi = torch.autograd.grad(Phi.sum(), XVl, retain_graph=True)[0][:, 0, :].to(XV.dtype)
            n = U.shape[0]
            V = float(self.V)
            h, f = self.h.to(U.dtype), self.f[:, 0].to(U.dtype)
            if self.c.func is not None:
                raise NotImplementedError("a callable c(X, u) on a single-time-point group (the reference's rank-2 "
                                          "shortcut broadcasts it to [n, n])")
            if self.a.per_path is not None:
                q = torch.einsum("nij,nj->ni", self.a.per_path.to(du), du)
            else:
                A = self.a.matrix.to(du) if self.a.matrix is not None else None
                q = du if A is None else du @ A.T                                       # sum_l a_kl du_l
            s31 = (dphi[:, 1:] * q).sum(1).double()
            bvec = self.b.per_path if self.b.per_path is not None else self.b.vector
            s32 = ((du * bvec.to(du)).sum(1).double().sum() * Phi.detach().sum()) if bvec is not None else 0.0
            cU = self.c.c0 + self.c.c1 * U
            s1 = (V / n) * (U * Vv - h * Vv).sum()
            s2 = (V / n) * (U.detach().sum() * dphi[:, 0].double().sum())
            s3 = (V / n) * (n * s31.sum() + s32 + n * (cU * U * Phi).sum() + f.sum() * Phi.sum())
            I = s1 - (s2 - s3)
            S = V * (Vv ** 2).sum() / n
            integ = torch.log(I ** 2) - torch.log(S)
            comps = dict(I=I.detach(), S=S.detach())
            if phase == "u":
                init = ((U.unsqueeze(0) - h.unsqueeze(1)) ** 2).mean()
                g = self.g.to(U.dtype)
                kb = u_mod.start_kind(border)
                sb = u_mod.initial_scalar(border.detach(), kb).double()
                ub = fin(lift(sb.unsqueeze(1)))[:, 0]
                if border.shape[1] == 1 and kb == "h":
                    bdry = ((ub.unsqueeze(0) - g[:, 0].unsqueeze(1)) ** 2).mean()       # rank-2 boundary: [nb, nb] too
                else:
                    bdry = ((ub - g[:, 0]) ** 2).mean()
                val = integ + self.alpha * (init + bdry)
                comps.update(init=init.detach(), bdry=bdry.detach())
                side = U.sum()
            else:
                val = -integ
                side = Phi.sum()
            if self.side_effect:              # the helper backward calls of src/loss.py:55,60 also fill .grad
                val = val + (side - side.detach())
        val.components = comps
        return val

    def _eval(self, phase, y_output_u, y_output_v, X, XV, border):
        if not isinstance(y_output_u, LazyPrediction) and X.shape[1] == 1 and isinstance(y_output_v, LazyPrediction):
            v_mod = y_output_v.net
            u_mod = getattr(y_output_u, "_xw_net", None) or self._u_module
            return self._single_time_group(phase, u_mod, v_mod, X, XV, border)
        u_mod, v_mod = self._nets(y_output_u, y_output_v)
        # the Batch holds only sample-dependent values (fp32 views, h, grad_h, f, g, initial scalars): the sub-steps of one
        # outer iteration run on ONE sample (src/training.py:125-162), so the solver may hand back the one built earlier;
        # a v-phase call (no boundary) can use a Batch that also carries the boundary part
        batch = self.batch_cache if self.batch_cache is not None else self._batch(u_mod, X, XV, border)
        self.last_batch = batch
        spec = u_mod.spec(v_mod)
        dom = domain_spec(self.domain)
        vbuf, vmode = self.vcache if self.vcache is not None else (None, 0)
        if vmode and vbuf is None:
            raise RuntimeError("vcache mode without a buffer")
        A_val, A_der = self._general_A(u_mod, X) if self.c.func is not None else (None, None)
        return hotpath.weak_loss(phase, spec, dom, self._coef(X.device, A_val, A_der), float(self.alpha), batch,
                                 u_mod.kernel_parameters(), v_mod.flat_parameters(), group=self.group,
                                 side_effect=self.side_effect, vcache=vbuf, vmode=vmode, grad_sink=self.grad_sink)

    # ------------------------------------------------------------------------------ reference API
    _u_module = None

    def u(self, y_output_u, y_output_v, u_net, X, XV, border):
        """loss_u = int + alpha*(init + bdry)   (reference src/loss.py:92-93)"""
        self._u_module = unwrap(u_net)
        return self._eval("u", y_output_u, y_output_v, X, XV, border)

    def v(self, y_output_u, y_output_v, X, XV):
        """loss_v = -int   (reference src/loss.py:95-96)"""
        return self._eval("v", y_output_u, y_output_v, X, XV, None)

    def _components(self, y_output_u, y_output_v, X, XV, border=None):
        with torch.no_grad():
            out = self._eval("u" if border is not None else "v", y_output_u, y_output_v, X, XV, border)
        return out.components if out.components is not None else {}

    def I(self, y_output_u, y_output_v, X, XV):
        """value of <A[u], phi> (reference src/loss.py:46-76); value only"""
        return self._value("I", y_output_u, y_output_v, X, XV)

    def int(self, y_output_u, y_output_v, X, XV):
        I = self._value("I", y_output_u, y_output_v, X, XV)
        S = self._value("S", y_output_u, y_output_v, X, XV)
        return torch.log(I ** 2) - torch.log(S)

    def init(self, y_output_u):
        u = y_output_u.materialize() if isinstance(y_output_u, LazyPrediction) else y_output_u
        return torch.mean((u[:, 0] - self.h.unsqueeze(1).to(u.dtype)) ** 2)

    def bdry(self, u_net, border_data):
        net = unwrap(u_net)
        with torch.no_grad():
            ub = net.evaluate(border_data)
        return torch.mean((ub - self.g.unsqueeze(2).to(ub.dtype)) ** 2)

    def _value(self, key, y_output_u, y_output_v, X, XV):
        u_mod, v_mod = self._nets(y_output_u, y_output_v)
        batch = self._batch(u_mod, X, XV, None)
        spec, dom = u_mod.spec(v_mod), domain_spec(self.domain)
        lib = hotpath._lib.get()
        with torch.no_grad():
            thu = hotpath.flatten_params(u_mod.kernel_parameters())
            thv = hotpath.flatten_params(v_mod.flat_parameters())
            A_val, A_der = self._general_A(u_mod, X) if self.c.func is not None else (None, None)
            sums, _, _ = hotpath.forward_sums(lib, spec, dom, self._coef(X.device, A_val, A_der), thu, thv, batch, False, 0.0)
            hotpath._allreduce(sums, self.group)
            I, S, init, bdry, integ = hotpath.loss_from_sums(sums, batch, dom.V, 0.0)
        return {"I": I, "S": S}[key]
