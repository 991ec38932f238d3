"""Drop-in module containers for the two networks of XNODE-WAN.

Same constructor signatures, sub-module names, parameter shapes, dtypes (float64) and sharing as
the reference (/root/reference/src/model.py:18-156), so state dicts interchange
(`initial_layers.{0,2,4}`, `ODE_rhs.net.{0,2,2*nu}`, `final_linear`; `input`, `hidden`, `output`,
`net.*`).  The arithmetic does not run in PyTorch: `forward` returns a `LazyPrediction` that the
fused loss (loss.py -> hotpath.py -> sm_100a kernels) consumes; if a caller needs the values
(L_norm, plotting) the prediction materialises through the forward-only kernels.
"""
import torch
from torch import nn

from . import hotpath
from .paths import CollapsedPaths


def init_weights(layer):
    """xavier-uniform weights, zero bias on every nn.Linear (reference src/model.py:12-15)"""
    if type(layer) == nn.Linear:
        nn.init.xavier_uniform_(layer.weight)
        layer.bias.data.fill_(0)


class LazyPrediction:
    """stand-in for `u_net(X)` / `v_net(XV)`: remembers the module and its input; turns into a real
    tensor ([N, L, 1], float64, no autograd graph) the first time something asks for values."""

    def __init__(self, net, inputs, kind):
        self.net, self.inputs, self.kind = net, inputs, kind
        self._value = None

    def materialize(self):
        if self._value is None:
            self._value = self.net.evaluate(self.inputs)
        return self._value

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def conv(a):
            if isinstance(a, LazyPrediction):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(conv(x) for x in a)
            return a
        return func(*conv(args), **{k: conv(v) for k, v in (kwargs or {}).items()})

    def __neg__(self): return -self.materialize()
    def __add__(self, o): return self.materialize() + o
    def __radd__(self, o): return o + self.materialize()
    def __sub__(self, o): return self.materialize() - o
    def __rsub__(self, o): return o - self.materialize()
    def __mul__(self, o): return self.materialize() * o
    def __rmul__(self, o): return o * self.materialize()
    def __truediv__(self, o): return self.materialize() / o
    def __pow__(self, o): return self.materialize() ** o
    def __getitem__(self, i): return self.materialize()[i]
    def __len__(self): return len(self.materialize())


def unwrap(net):
    """strip torch.nn.DataParallel (reference src/training.py:93,97 wraps both nets)"""
    return net.module if isinstance(net, nn.DataParallel) else net


class discriminator(nn.Module):
    """test-function net v_phi: Linear(d+1,Hv), [ReLU, hidden] x v_layers sharing ONE Linear,
    Tanh, Linear(Hv,1); float64 (reference src/model.py:30-47)."""

    def __init__(self, config: dict, setup: dict):
        super().__init__()
        self.num_layers = config['v_layers']
        self.hidden_dim = config['v_hidden_dim']
        self.dim = setup['dim']
        self.input = nn.Linear(self.dim + 1, self.hidden_dim)
        self.hidden = nn.Linear(self.hidden_dim, self.hidden_dim)
        self.output = nn.Linear(self.hidden_dim, 1)
        stack = [self.input]
        for _ in range(self.num_layers):
            stack += [nn.ReLU(), self.hidden]
        stack += [nn.Tanh(), self.output]
        self.net = nn.Sequential(*stack)
        self.net.double()

    def flat_parameters(self):
        return [self.input.weight, self.input.bias, self.hidden.weight, self.hidden.bias,
                self.output.weight, self.output.bias]

    def evaluate(self, XV):
        spec = hotpath.NetSpec(self.dim, 1, 1, 1, self.hidden_dim, self.num_layers)
        if isinstance(XV, CollapsedPaths):
            XV = XV.dense()
        v = hotpath.vnet_eval(spec, self.flat_parameters(), XV)
        return v.double().unsqueeze(2)

    def forward(self, XV: torch.Tensor):
        if torch.is_grad_enabled():
            return LazyPrediction(self, XV, "v")
        return self.evaluate(XV)


class _ODEField(nn.Module):
    """parameter container of the vector field F (reference src/model.py:115-141): Linear(H+d+1,hh),
    [ReLU, Linear(hh,hh)] x (num_layers-1) sharing ONE Linear, Tanh, Linear(hh,H); input order (x,t,y)."""

    def __init__(self, input_dim: int, setup: dict, num_layers: int, hidden_dim: int):
        super().__init__()
        if num_layers < 1:
            raise RuntimeError("u_layers < 1 is not supported by the fused XNODE kernels")
        self.input_dim, self.hidden_dim, self.num_layers = input_dim, hidden_dim, num_layers
        shared = nn.Linear(hidden_dim, hidden_dim)
        stack = [nn.Linear(input_dim + setup['dim'] + 1, hidden_dim)]
        for _ in range(num_layers - 1):
            stack += [nn.ReLU(), shared]
        stack += [nn.Tanh(), nn.Linear(hidden_dim, input_dim)]
        self.net = nn.Sequential(*stack).double()


class NeuralODE(nn.Module):
    """XNODE primal net u_theta (reference src/model.py:54-112)."""

    def __init__(self, hidden_dim: int, output_dim: int, func_h, func_g, setup: dict, hidden_hidden_dim: int,
                 num_layers: int, domain, solver: str = 'midpoint', min_steps: int = 5, adjoint: bool = False):
        super().__init__()
        if output_dim != 1:
            raise RuntimeError("output_dim must be 1")
        if adjoint:
            raise RuntimeError("adjoint=True is not supported (the fused backward kernel replaces it)")
        if solver not in ("euler", "midpoint", "rk4"):
            raise RuntimeError("unsupported solver %r (supported: euler, midpoint, rk4)" % (solver,))
        self.hidden_dim, self.output_dim = hidden_dim, output_dim
        self.h, self.g = func_h, func_g
        self.setup = setup
        self.hidden_hidden_dim, self.num_layers = hidden_hidden_dim, num_layers
        self.domain, self.solver, self.min_steps, self.adjoint = domain, solver, min_steps, adjoint
        self.initial_layers = nn.Sequential(nn.Linear(1, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim),
                                            nn.ReLU(), nn.Linear(hidden_dim, hidden_dim)).double()
        self.ODE_rhs = _ODEField(hidden_dim, setup, hidden_dim=hidden_hidden_dim, num_layers=num_layers)
        self.ODE_rhs.apply(init_weights)
        self.final_linear = nn.Linear(hidden_dim, output_dim).double()

    def flat_parameters(self):
        il, net = self.initial_layers, self.ODE_rhs.net
        last = net[len(net) - 1]
        shared = net[2] if self.num_layers > 1 else None
        ps = [il[0].weight, il[0].bias, il[2].weight, il[2].bias, il[4].weight, il[4].bias, net[0].weight, net[0].bias]
        if shared is not None:
            ps += [shared.weight, shared.bias]
        ps += [last.weight, last.bias, self.final_linear.weight, self.final_linear.bias]
        return ps

    def kernel_parameters(self):
        """the 14 tensors of the flat layout; u_layers == 1 has no shared layer -> zeros placeholder"""
        ps = self.flat_parameters()
        if self.num_layers == 1:
            hh = self.hidden_hidden_dim
            z = ps[0]
            ps = ps[:8] + [z.new_zeros(hh, hh), z.new_zeros(hh)] + ps[8:]
        return ps

    def spec(self, v_net=None):
        vh, vl = (v_net.hidden_dim, v_net.num_layers) if v_net is not None else (1, 0)
        return hotpath.NetSpec(self.setup['dim'], self.hidden_dim, self.hidden_hidden_dim, self.num_layers, vh, vl,
                               self.solver)

    def start_kind(self, inputs):
        """which branch of reference src/model.py:89-96 a batch takes: 'h' (starts at T0),
        'g' (starts on the boundary), or 'pad' (needs bound_pad: unsupported here)"""
        tag = getattr(inputs, "_xw_start", None)
        if tag is not None:
            return tag
        if bool(inputs[0, 0, 0] == self.setup['T0']):
            return "h"
        if bool(torch.max(self.domain.func_w(inputs[:, 0, :].unsqueeze(1))) < 1e-5):
            return "g"
        return "pad"

    def initial_scalar(self, inputs, kind):
        if kind == "h":
            return self.h(inputs[:, 0, :])
        return self.g(inputs[:, 0, :].unsqueeze(1)).reshape(-1)

    def _initial_scalar_f32(self, inputs, kind):
        """h / g on time-row 0 as the kernels take it; it does not depend on the parameters, so repeated evaluations on
        one sample (stop() after every u sub-iteration, reference src/training.py:142) compute it once"""
        ver = (inputs.times._version, inputs.x._version) if isinstance(inputs, CollapsedPaths) else inputs._version
        fn = self.h if kind == "h" else self.g
        kept = getattr(inputs, "_xw_s0", None)
        if kept is not None and kept[0] == ver and kept[1] is fn:
            return kept[2]
        s0 = hotpath.as_f32(self.initial_scalar(inputs.detach(), kind))
        try:
            inputs._xw_s0 = (ver, fn, s0)
        except AttributeError:
            pass
        return s0

    def evaluate(self, inputs):
        kind = self.start_kind(inputs)
        if inputs.shape[1] == 1 and kind == "h":
            # rank-2 shortcut of the reference (src/model.py:89-91): no ODE, plain PyTorch
            h_ = self.h(inputs[:, 0, :]).unsqueeze(1).double()
            return self.final_linear(self.initial_layers(h_))
        if kind == "pad":
            return self._evaluate_from_inside(inputs)
        s0 = self._initial_scalar_f32(inputs, kind)
        if isinstance(inputs, CollapsedPaths):
            xs, ts = hotpath.as_f32(inputs.x), hotpath.as_f32(inputs.times)
            u = hotpath.xnode_eval(self.spec(), self.kernel_parameters(), xs, 0, xs.shape[1], ts, s0, xs.shape[0])
            return u.double().unsqueeze(2)
        Xf = hotpath.as_f32(inputs)
        N, L, Cc = Xf.shape
        u = hotpath.xnode_eval(self.spec(), self.kernel_parameters(), Xf, 1, L * Cc, Xf[0, :, 0].contiguous(), s0, N)
        return u.double().unsqueeze(2)

    def _evaluate_from_inside(self, inputs):
        """u at requested times for paths that start inside the domain after T0 (reference src/model.py:92-94,
        104-106): the domain pads the time grid back to T0 and fills large gaps (`bound_pad` / `fillt`), the XNODE is
        integrated on that grid from the spatial point of time-row 0, and the requested times are looked up.  As in the
        reference the initial scalar is g at time-row 0 (`inputs[0,0,0] != T0`, src/model.py:95-96)."""
        if isinstance(inputs, CollapsedPaths):
            raise NotImplementedError("evaluation from inside the domain takes the dense [N, L, C] layout")
        path_i, pos, grid = self.domain.bound_pad(inputs.detach())
        if path_i is not None:
            return self._evaluate_grouped(inputs, path_i, pos, grid)
        if grid.numel() < 2 or int(pos.max()) >= grid.numel():
            raise RuntimeError("fillt produced a %d-point grid for the requested times (no gap larger than "
                               "(T - T0) / N_t between them): the reference fails on this input too" % grid.numel())
        s0 = hotpath.as_f32(self.initial_scalar(inputs.detach(), "g"))
        Xf = hotpath.as_f32(inputs)
        N, L, Cc = Xf.shape
        ts = hotpath.as_f32(grid.to(Xf.device))
        u = hotpath.xnode_eval(self.spec(), self.kernel_parameters(), Xf, 1, L * Cc, ts, s0, N)
        return u[:, pos.to(u.device).long()].double().unsqueeze(2)

    def _evaluate_grouped(self, inputs, path_i, pos, grids):
        """hourglass (reference src/model.py:104-107, the `path_i is not None` branch): one integration per group of
        paths on the group's own grid, rows of the result in GROUP order (the reference concatenates the groups)."""
        s0 = hotpath.as_f32(self.initial_scalar(inputs.detach(), "g"))
        Xf = hotpath.as_f32(inputs)
        L, Cc = Xf.shape[1], Xf.shape[2]
        outs = []
        for pi, ps, grid in zip(path_i, pos, grids):
            if grid.numel() < 2 or int(ps.max()) >= grid.numel():
                raise RuntimeError("fillt produced a %d-point grid for the requested times: the reference fails on this "
                                   "input too" % grid.numel())
            sel = pi.to(Xf.device).long()
            Xg = Xf[sel].contiguous()
            u = hotpath.xnode_eval(self.spec(), self.kernel_parameters(), Xg, 1, L * Cc, hotpath.as_f32(grid.to(Xf.device)),
                                   s0[sel].contiguous(), Xg.shape[0])
            outs.append(u[:, ps.to(u.device).long()])
        return torch.cat(outs, 0).double().unsqueeze(2)

    def forward(self, inputs: torch.Tensor):
        if torch.is_grad_enabled() and not (inputs.shape[1] == 1):
            return LazyPrediction(self, inputs, "u")
        out = self.evaluate(inputs)
        out._xw_net = self            # lets loss.v find the module for the single-time-point group
        return out
