"""Flat parameter storage and the single-launch Adam step for the two nets (reference: two `torch.optim.Adam`
instances, /root/reference/src/training.py:103-104, stepped once per batch at :138 / :162).

The kernels read each net's parameters as ONE flat fp32 vector and return ONE flat fp32 gradient, while the reference
API exposes 14 + 6 separate fp64 `nn.Parameter`s.  `FlatParameters` re-homes the parameters of a net as views of one
fp64 buffer (names, shapes, dtypes and `state_dict()` unchanged) and keeps the fp32 copy the kernels read;
`FusedAdam` steps that buffer from the flat gradient with one kernel launch (`xw_adam_step`) instead of torch's foreach
Adam (~15 launches) + 14 gradient casts + re-packing -- which matters at the shipped N = 4000, where a sub-iteration is
launch-bound.  Same arithmetic as `torch.optim.Adam` defaults (betas 0.9 / 0.999, eps 1e-8, no weight decay) in fp64.
"""
import ctypes as C

import torch


class FlatParameters:
    def __init__(self, params):
        params = list(params)
        p0 = params[0]
        total = sum(p.numel() for p in params)
        flat = torch.empty(total, dtype=p0.dtype, device=p0.device)
        o = 0
        for p in params:
            n = p.numel()
            flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = flat[o:o + n].view(p.shape)
            o += n
        self.params, self.flat, self.total = params, flat, total
        self.f32 = flat.float()
        self._seen = self._versions()
        for p in params:
            p._xw_flat = self

    def _versions(self):
        return tuple(p._version for p in self.params)

    def intact(self):
        """the parameters still live in the flat buffer (a later .to() / .double() would re-allocate them)"""
        o, base, es = 0, self.flat.data_ptr(), self.flat.element_size()
        for p in self.params:
            if p.data_ptr() != base + o * es or p.dtype != self.flat.dtype:
                return False
            o += p.numel()
        return True

    def theta32(self):
        """flat fp32 parameters for the kernels.  Refreshed from the fp64 buffer when a parameter was modified in place
        since the last refresh (version counters), and always inside a CUDA-graph capture (a replay cannot check)."""
        capturing = self.flat.is_cuda and torch.cuda.is_current_stream_capturing()
        v = self._versions()
        if capturing or v != self._seen:
            self.f32.copy_(self.flat)
            self._seen = v
        return self.f32


def flat_of(params):
    """the FlatParameters object holding exactly `params` (in order), or None"""
    fp = getattr(params[0], "_xw_flat", None) if len(params) else None
    if fp is None or len(fp.params) != len(params) or any(a is not b for a, b in zip(fp.params, params)) or not fp.intact():
        return None
    return fp


class FusedAdam:
    """drop-in for the solver's `torch.optim.Adam(net.parameters(), lr=...)`: `zero_grad()` / `step()`; the gradient
    arrives in `grad32` (written by the backward kernels through `loss.grad_sink`), not in `p.grad`"""

    def __init__(self, flat, lr, lib, betas=(0.9, 0.999), eps=1e-8):
        self.flat, self.lib = flat, lib
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        dev = flat.flat.device
        self.grad32 = torch.zeros(flat.total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(flat.total, dtype=torch.float64, device=dev)
        self.exp_avg_sq = torch.zeros(flat.total, dtype=torch.float64, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.param_groups = [{"params": flat.params, "lr": self.lr, "betas": self.betas, "eps": self.eps}]

    def zero_grad(self, set_to_none=True):
        pass        # grad32 is overwritten (not accumulated into) by the next backward

    def step(self):
        f = self.flat
        dev = f.flat.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream) if dev.type == "cuda" else None
        p = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
        self.lib.call("xw_adam_step", p(f.flat), p(self.grad32), p(self.exp_avg), p(self.exp_avg_sq), p(self.step_count),
                      p(f.f32), f.total, self.lr, self.betas[0], self.betas[1], self.eps, st)

    def state_dict(self):
        return {"step": self.step_count.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": self.lr, "betas": self.betas, "eps": self.eps}

    def load_state_dict(self, sd):
        self.step_count.copy_(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
