"""Domain samplers with the reference's protocol (/root/reference/src/dataset.py:34-45):
`interior(N_r)`, `boundary(N_b)`, `func_w(x)`, `V()` and the `Comb_loader(N_r, N_b, shape, device)`
iterable of (datau, datav, bdata) triples in the [N, L, C] layout with time in channel 0.

`Hypercube` draws the SAME random numbers in the same order as the reference when sampling on the
CPU (so a seeded run yields bit-identical samples: times -> interior x -> second interior x for the
test function -> boundary x -> one unused draw -> permutation, src/dataset.py:248-276,306-310), and
can alternatively sample directly on the GPU (`sample_device=`) for the large configurations where
host sampling + H2D of the repeated [N, L, C] tensors would dominate.
"""
import math

import numpy as np
import torch

from .paths import CollapsedPaths


def fillt(times: torch.Tensor, T: float, T0: float, min_steps: int = 5):
    """Time grid for evaluating u at requested times that do not form a usable integration grid
    (reference src/dataset.py:13-32): wherever two consecutive requested times are more than
    (T - T0) / min_steps apart, equally spaced steps of about that size are inserted.  Returns
    (positions of the requested times in the new grid, the new grid).  Like the reference, a request
    without any such gap yields a one-point grid (the reference then fails on the lookup; so do we,
    with a clear message, in NeuralODE.evaluate)."""
    step = (T - T0) / min_steps
    n_req = times.shape[0]
    gaps = torch.cat((torch.tensor(step / 2, device=times.device).view(1), (times[:-1] - times[1:]).abs()), 0)
    starts = torch.nonzero(gaps > step).squeeze(1).tolist() + [n_req]
    pos = torch.arange(n_req, device=times.device)
    grid = times[0].view(1)
    for a, b in zip(starts[:-1], starts[1:]):
        target, last = times[a].item(), grid[-1].item()
        n_fill = round((target - 2 * step - last) / step) + 1
        filler = torch.linspace(last + step, target - step, n_fill).to(times.device)
        pos[a:] += filler.shape[0]
        grid = torch.cat((grid, filler, times[a:b].to(times.device)), 0)
    return pos, grid


class Hypercube:
    """cube (bot..top)^dim x [T0, T]; time grid = sorted uniforms with pinned end points"""

    def __init__(self, top_bot, dim: int, T0: float, T: float, N_t: int, sample_device=None, times=None,
                 collapsed=False):
        assert top_bot[1] > top_bot[0], "The hypercube needs to have volume"
        self.bot, self.top = top_bot[0], top_bot[1]
        self.dim, self.T0, self.T, self.N_t = dim, T0, T, N_t
        self.sample_device = sample_device
        self.collapsed = collapsed      # yield CollapsedPaths(times, x) instead of dense [N, L, C]
        self.pin_host = False           # host-sampled dense tensors in page-locked memory (asynchronous H2D; set by the solver)
        if times is None:
            times = torch.empty(N_t).uniform_(T0, T).sort(0).values
            times[0], times[-1] = T0, T
        self.times = times

    # -- raw draws -------------------------------------------------------------------------
    def _uniform_x(self, n):
        if self.sample_device is None:
            return torch.empty(n, 1, self.dim).uniform_(self.bot, self.top)
        return torch.empty(n, 1, self.dim, device=self.sample_device).uniform_(self.bot, self.top)

    def _with_time(self, x):
        n = x.shape[0]
        if self.collapsed:
            return CollapsedPaths(self.times.to(x.device), x[:, 0, :].contiguous())
        t = self.times.to(x.device).reshape(1, self.N_t, 1).expand(n, self.N_t, 1)
        if self.pin_host and x.device.type == "cpu" and n * self.N_t * (self.dim + 1) * x.element_size() <= (256 << 20):
            out = torch.empty(n, self.N_t, self.dim + 1, dtype=x.dtype, pin_memory=True)
            return torch.cat((t, x.expand(n, self.N_t, self.dim)), dim=2, out=out)
        return torch.cat((t, x.expand(n, self.N_t, self.dim)), dim=2)

    def interior(self, N_r: int):
        return self._with_time(self._uniform_x(N_r))

    def boundary_x(self, N_b: int):
        """[N_b, dim] spatial coordinates on the faces: consecutive blocks of int(N_b/dim/2) paths get
        coordinate k pinned to top / bot, the remainder goes to the last face; then a permutation"""
        x = self._uniform_x(N_b)[:, 0, :].clone()
        self._uniform_x(N_b)                                   # the reference draws (and drops) one more
        n = int(N_b / self.dim / 2)
        edges = [n * i for i in range(2 * self.dim)] + [N_b]
        for k in range(self.dim):
            x[edges[2 * k]:edges[2 * k + 1], k] = self.top
            x[edges[2 * k + 1]:edges[2 * k + 2], k] = self.bot
        perm = torch.randperm(N_b) if self.sample_device is None else torch.randperm(N_b, device=self.sample_device)
        return x[perm]

    def boundary(self, N_b: int):
        return self._with_time(self.boundary_x(N_b).unsqueeze(1))

    def func_w(self, x: torch.Tensor):
        s = x[:, :, 1:]
        return torch.minimum((self.top - s).abs().amin(dim=2), (self.bot - s).abs().amin(dim=2))

    def bound_pad(self, x):
        """integration grid for a batch that starts inside the domain after T0 (src/dataset.py:284-287):
        (None, positions of the requested times, grid starting at T0)"""
        t = torch.cat((torch.tensor(self.T0, dtype=x.dtype).view(1).to(x.device), x[0, :, 0]), 0)
        pos, grid = fillt(t, self.T, self.T0, self.N_t)
        return None, pos[1:], grid

    def V(self):
        return (self.top - self.bot) ** self.dim * (self.T - self.T0)


_LOADER_UID = [0]
_COPY_STREAMS = {}


def _copy_stream(dev):
    if dev not in _COPY_STREAMS:
        _COPY_STREAMS[dev] = torch.cuda.Stream(dev)
    return _COPY_STREAMS[dev]


class Comb_loader:
    """(datau, datav, bdata) batches of equally long paths (reference src/dataset.py:293-322).
    For a tensor-valued domain (the cube) there is one batch; the test-function points are an
    independent second interior sample (src/dataset.py:308)."""

    def __init__(self, N_r: int, N_b: int, shape, device):
        self.N_r, self.N_b, self.shape, self.device = N_r, N_b, shape, device
        _LOADER_UID[0] += 1
        self.uid = _LOADER_UID[0]          # identifies the sample (the solver's test-function cache keys on it)
        interior = shape.interior(N_r)
        if isinstance(interior, list):
            self.interioru = interior
            self.interiorv = [g.clone().detach() for g in interior]
            self.boundary = shape.boundary(N_b)
        else:
            self.interioru = interior
            self.interiorv = shape.interior(N_r)
            self.boundary = shape.boundary(N_b)
        self._cache = {}

    @classmethod
    def from_tensors(cls, interioru, interiorv, boundary, device):
        """wrap an existing sample (e.g. pinned host tensors / CollapsedPaths); the host->device copy
        happens on first access, as in the reference's __getitem__ (src/dataset.py:321)"""
        self = cls.__new__(cls)
        _LOADER_UID[0] += 1
        self.uid = _LOADER_UID[0]
        self.N_r, self.N_b = interioru.shape[0], boundary.shape[0]
        self.shape, self.device = None, device
        self.interioru, self.interiorv, self.boundary = interioru, interiorv, boundary
        self._cache = {}
        return self

    def prefetch(self, stream=None):
        """start the host->device copy of this sample on a side stream NOW (pinned host tensors: truly asynchronous), so
        that it overlaps the previous sample's kernels; the first access makes the consumer's stream wait for it.
        Returns self.  (The reference copies synchronously at access time, src/dataset.py:321.)"""
        dev = torch.device(self.device)
        if dev.type != "cuda" or isinstance(self.interioru, list) or 0 in self._cache:
            return self
        stream = stream or _copy_stream(dev)
        stream.wait_stream(torch.cuda.current_stream(dev))        # the tensors being replaced may still be in use
        with torch.cuda.stream(stream):
            trip = tuple(self._dev(t, "h", publish=False) for t in (self.interioru, self.interiorv, self.boundary))
            ev = torch.cuda.Event()
            ev.record(stream)
        self._cache[0] = trip
        self._pending = (ev, stream, trip)
        return self

    def __len__(self):
        return len(self.interioru) if isinstance(self.interioru, list) else 1

    def _dev(self, t, start, publish=True):
        out = t.to(self.device, non_blocking=True)
        out._xw_start = start
        if publish:
            self._publish(t, out)
        return out

    @staticmethod
    def _publish(t, out):
        """the error norms evaluated on this sample (aux.L_norm <- stop(), every u sub-iteration) reuse the device copy
        instead of moving the host tensor again; the version counter guards against in-place edits of the sample.
        (A prefetched copy is published once the consumer's stream has been ordered after it, in __getitem__.)"""
        if out is not t:
            t._xw_dev = ((t.times._version, t.x._version) if hasattr(t, "times") else t._version, out)

    def __getitem__(self, idx):
        is_list = isinstance(self.interioru, list)
        if not is_list and idx != 0:
            raise IndexError
        pend = getattr(self, "_pending", None)
        if pend is not None:                    # a prefetch is in flight: order the consumer after it, once
            ev, stream, trip = pend
            cur = torch.cuda.current_stream(torch.device(self.device))
            cur.wait_event(ev)
            for t in trip:                      # the caching allocator must not recycle them while `cur` still reads
                for part in ((t.times, t.x) if hasattr(t, "times") else (t,)):
                    part.record_stream(cur)
            for t, out in zip((self.interioru, self.interiorv, self.boundary), trip):
                self._publish(t, out)
            self._pending = None
        if idx not in self._cache:      # one H2D per sample (the reference re-copies on every access)
            if is_list:
                if idx >= min(len(self.interioru), len(self.boundary)):
                    raise IndexError
                trip = (self.interioru[idx], self.interiorv[idx], self.boundary[idx])
                self._cache[idx] = tuple(self._dev(t, None) for t in trip)
            else:
                self._cache[idx] = (self._dev(self.interioru, "h"), self._dev(self.interiorv, "h"),
                                    self._dev(self.boundary, "h"))
        return self._cache[idx]


def _ball_volume(dim, r):
    return math.pi ** (dim / 2) / math.gamma(dim / 2 + 1) * r ** dim


class _SphereDomain:
    """time-varying ball domains of the reference (src/dataset.py:48-229).  Paths are groups of equal
    length (float64, requires no grad here): a point stays in the domain only while |x| < radius(t).
    Random numbers are drawn in the reference's order (numpy normal for the directions, numpy rand
    for the radii, torch uniform for the time grid), so seeded runs reproduce its groups exactly."""

    def __init__(self, r: float, dim: int, T0: float, T: float, N_t: int, times=None, sample_device=None):
        self.r, self.dim, self.T0, self.T, self.N_t = r, dim, T0, T, N_t
        if times is None:
            times = torch.empty(N_t).uniform_(T0, T).sort(0).values
            times[0], times[-1] = T0, T
        self.times = times
        # sample_device: draw the points with torch's generator of that device and build the groups with vectorised
        # tensor ops there (same distributions and group structure, another RNG stream than the reference's numpy one)
        self.sample_device = sample_device

    def _surf_t(self, N: int):
        """N points on the sphere of radius r, [N, dim] float64 on the sampling device"""
        z = torch.randn(N, self.dim, dtype=torch.float64, device=self.sample_device)
        return self.r * z / z.norm(dim=1, keepdim=True)

    def _ball_t(self, N: int):
        return self._surf_t(N) * torch.rand(N, 1, dtype=torch.float64, device=self.sample_device) ** (1.0 / self.dim)

    def _group(self, tt, x):
        """[n, k, C] float64 group: times tt[k] (shared) next to the points x[n, dim]"""
        n, k = x.shape[0], tt.shape[0]
        return torch.cat((tt.view(1, k, 1).expand(n, k, 1), x.unsqueeze(1).expand(n, k, x.shape[1])), 2)

    def surf(self, N: int):
        """N points on the sphere of radius r, [dim, N] (normalised normal deviates)"""
        z = np.random.normal(size=(self.dim, N))
        return self.r * z / np.sqrt((z ** 2).sum(axis=0))

    def _ball(self, N: int):
        pts = self.surf(N)
        pts *= np.random.rand(N) ** (1 / self.dim)
        return pts

    def _radius_scale(self, t):
        raise NotImplementedError

    def boundary(self, N_b: int):
        """one single-time group per grid time with int(N_b * scale(t)^dim) points on the sphere of that time"""
        groups = []
        for t in self.times.numpy():
            sc = self._radius_scale(t)
            n = int(N_b * sc ** self.dim)
            if self.sample_device is not None:
                if n != 0:
                    tt = torch.full((1,), float(t), dtype=torch.float64, device=self.sample_device)
                    groups.append(self._group(tt, self._surf_t(n) * float(sc)))
                continue
            x = torch.from_numpy(self.surf(n) * sc).transpose(0, 1).unsqueeze(1)
            if n != 0:
                groups.append(torch.cat((t * torch.ones(n, 1, 1), x), 2))
        return groups

    def bound_pad(self, x):
        raise NotImplementedError("bound_pad of %s: evaluate from inside the domain on the cube, the cone or the hourglass"
                                  % type(self).__name__)


class NSphere_TCone(_SphereDomain):
    """ball of radius r (1 - t)   (reference src/dataset.py:162-229)"""

    def _radius_scale(self, t):
        return 1 - t

    def interior(self, N_r: int):
        """a path lives on times[0:k) where k = number of grid times with |x| < r (1 - t); paths of equal k
        form one group [n, k, C]; groups in ascending k"""
        if self.sample_device is not None:
            return self._groups_from(self._ball_t(N_r))
        pts = self._ball(N_r)                                       # [dim, N_r]
        grid = self.times.repeat(N_r, 1).unsqueeze(2)               # [N_r, N_t, 1]
        groups = []
        k = self.N_t
        for t in self.times.numpy()[::-1]:
            inside = np.sqrt(np.sum(pts ** 2, 0)) < self.r * (1 - t)
            x = torch.from_numpy(pts[:, inside]).transpose(0, 1).unsqueeze(1).repeat(1, k, 1)
            pts = np.delete(pts, inside, 1)
            if x.shape[0] != 0:
                groups.append(torch.cat((grid[:x.shape[0], :k], x), 2))
            k -= 1
        return groups[::-1]

    def _groups_from(self, pts):
        """the same groups from given points pts[N, dim] with tensor ops on pts' device: the radius shrinks with t, so
        a path's length is the NUMBER of grid times at which it is inside; original order inside a group"""
        tt = self.times.to(pts.device).double()
        k_n = (pts.norm(dim=1).unsqueeze(1) < self.r * (1 - tt).unsqueeze(0)).sum(1)
        return [self._group(tt[:k], pts[k_n == k]) for k in torch.unique(k_n).tolist() if k > 0]

    def func_w(self, x: torch.Tensor):
        return self.r * (1 - x[:, :, 0]) - x[:, :, 1:].pow(2).sum(2).sqrt()

    def bound_pad(self, x):
        """src/dataset.py:220-223: as for the cube, every path starts at T0"""
        t = torch.cat((torch.tensor(self.T0, dtype=x.dtype).view(1).to(x.device), x[0, :, 0]), 0)
        pos, grid = fillt(t, self.T, self.T0, self.N_t)
        return None, pos[1:], grid

    def V(self):
        tc = (1 - self.T0) ** (self.dim + 1) / (self.dim + 1) - (1 - self.T) ** (self.dim + 1) / (self.dim + 1)
        return _ball_volume(self.dim, self.r) * tc


class NSphere_THourglass(_SphereDomain):
    """ball of radius r ((T-T0) - t) for t <= (T-T0)/2 and r t afterwards (reference src/dataset.py:48-159)"""

    def _radius_scale(self, t):
        return (self.T - self.T0) - t if t < (self.T - self.T0) / 2 else t

    def interior(self, N_r: int):
        """every path is inside at T0, leaves when the ball shrinks below |x| and re-enters when it grows
        back: it contributes a first segment (times before the exit) and, if it was ever outside, a second
        segment (times after the re-entry) that gets one extra leading row at the exact entry time
        t = |x| / r.  Segments of equal length are concatenated into groups, ascending length."""
        if self.sample_device is not None:
            return self._groups_from(self._ball_t(N_r))
        pts = torch.from_numpy(self._ball(N_r)).transpose(0, 1)    # [N_r, dim] float64
        grid = self.times.repeat(N_r, 1)                            # [N_r, N_t]
        half = (self.T - self.T0) / 2
        radius = torch.where(grid <= half, self.r * ((self.T - self.T0) - grid).double(), self.r * grid.double())
        inside = pts.pow(2).sum(1).sqrt().unsqueeze(1) < radius     # [N_r, N_t]
        full = torch.cat((grid.unsqueeze(2).double(), pts.unsqueeze(1).repeat(1, self.N_t, 1)), 2).to(torch.float64)
        first, second = [], []
        for k in range(N_r):
            out = torch.nonzero(~inside[k]).flatten()
            if out.numel() == 0:
                first.append(full[k].unsqueeze(0))
                continue
            a, b = int(out[0]), self.N_t - int(out[-1]) - 1
            kept = full[k, inside[k]]
            first.append(kept[:a].unsqueeze(0))
            seg = kept[a:a + b]
            entry = torch.cat(((seg[0, 1:].pow(2).sum().sqrt() / self.r).view(1, 1), seg[0, 1:].unsqueeze(0)), dim=1)
            second.append(torch.cat((entry, seg), 0).unsqueeze(0))

        def grouped(segs):
            out = {}
            for sgm in sorted(segs, key=lambda z: z.shape[1]):
                out.setdefault(sgm.shape[1], []).append(sgm)
            return [torch.cat(v, 0) for _, v in sorted(out.items())]
        return sorted(grouped(first) + grouped(second), key=lambda z: z.shape[1])

    def _groups_from(self, pts):
        """the same groups from given points pts[N, dim] with tensor ops on pts' device.  The radius is V-shaped in t,
        so the times a path spends outside form one interval: a = length of the inside prefix, b = of the inside suffix"""
        tt = self.times.to(pts.device).double()
        N_t, span = self.N_t, self.T - self.T0
        nrm = pts.pow(2).sum(1).sqrt()                 # (the reference's formula, bit for bit: it becomes the entry time)
        radius = torch.where(self.times.to(pts.device) <= span / 2, self.r * (span - tt), self.r * tt)
        inside = nrm.unsqueeze(1) < radius.unsqueeze(0)                       # [N, N_t]
        a_n = torch.cumprod(inside.long(), 1).sum(1)
        b_n = torch.cumprod(inside.flip(1).long(), 1).sum(1)
        never_out = a_n == N_t
        first, second = [], []
        for a in torch.unique(a_n).tolist():
            if a > 0:
                first.append(self._group(tt[:a], pts[a_n == a]))
        for b in torch.unique(b_n[~never_out]).tolist():
            sel = (~never_out) & (b_n == b)
            x = pts[sel]
            seg = self._group(tt[N_t - b:], x)
            entry = torch.cat(((nrm[sel] / self.r).view(-1, 1, 1), x.unsqueeze(1)), 2)
            second.append(torch.cat((entry, seg), 1))
        return sorted(first + second, key=lambda z: z.shape[1])

    def func_w(self, x: torch.Tensor):
        t = x[:, :, 0]
        dist = x[:, :, 1:].pow(2).sum(2).sqrt()
        span = self.T - self.T0
        return torch.where(t <= span / 2, self.r * (span - t) - dist, self.r * t - dist)

    def bound_pad(self, x):
        """integration grids for a batch that starts inside the hourglass after T0 (reference src/dataset.py:127-152).
        Unlike the cube and the cone the paths do not share one grid: a batch whose first time lies in the growing half
        contains paths that were inside all along (|x| <= r (T-T0)/2: integrated from T0) and paths that entered at
        t = |x| / r (integrated from there), so every path gets its own filled grid.  Paths are then ordered by the
        LENGTH of their grid (stable) and paths of equal length form one group that is integrated on the grid -- and
        read at the positions -- of its FIRST path (the reference keeps only `next(r)` of each group).  A path whose first
        point already sits on the boundary (w <= 1e-5) keeps the prepended entry time in its grid but has its positions
        computed without it (the reference's `t_[k][i[k]:]`).  Returns (path indices per group, positions per group,
        grid per group); `NeuralODE` concatenates the groups in this order, as the reference does."""
        n = x.shape[0]
        times = x[:, :, 0]
        half = (self.T - self.T0) / 2
        if bool(x[0, 0, 0] < half):
            start = torch.full_like(times[:, 0], self.T0)
            on_bdry = None
        else:
            nrm = x[:, 0, 1:].pow(2).sum(-1).sqrt()
            start = torch.where(nrm <= self.r * half, torch.full_like(nrm, self.T0), nrm / self.r)
            on_bdry = (self.func_w(x[:, 0].unsqueeze(1)) <= 1e-5).reshape(-1)
        seqs = torch.cat((start.unsqueeze(1), times), dim=1)
        grids, poss = [], []
        for k in range(n):
            pos_k, grid_k = fillt(seqs[k], self.T, self.T0, self.N_t)
            if on_bdry is not None:
                pos_k = fillt(seqs[k][1:], self.T, self.T0, self.N_t)[0] if bool(on_bdry[k]) else pos_k[1:]
            grids.append(grid_k)
            poss.append(pos_k)
        order = sorted(range(n), key=lambda k: grids[k].shape[0])
        path_i, pos_g, grid_g = [], [], []
        for k in order:
            if grid_g and grids[k].shape[0] == grid_g[-1].shape[0]:
                path_i[-1].append(k)
            else:
                path_i.append([k]); pos_g.append(poss[k]); grid_g.append(grids[k])
        return [torch.tensor(q) for q in path_i], pos_g, grid_g

    def V(self):
        tc = 2 * ((1 - self.T0) ** (self.dim + 1) / (self.dim + 1) -
                  (1 - (self.T - self.T0) / 2) ** (self.dim + 1) / (self.dim + 1))
        return _ball_volume(self.dim, self.r) * tc
