"""Golden vectors for the reference's "evaluate u from inside the domain" branch (src/model.py:92-94,104-106 with
src/dataset.py:13-32 fillt and :284-287 / :220-223 bound_pad): inputs that neither start at T0 nor sit on the
boundary.  Runs the UNMODIFIED reference (oracle/ref_runner.py) with the parameters of an existing golden case and
stores inputs + outputs under tests/golden/extra/.   usage: python tests/golden/make_pad_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_runner as rr  # noqa: E402
from tests import _golden as G       # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "extra")
CASES = {
    # name: (base golden case, requested times, seed)
    "pad_cube_d3_rk4": ("cube_d3_rk4", [0.3, 0.35, 0.9], 11),
    "pad_cube_d5_midpoint": ("cube_d5_alpha1_randbias", [0.12, 0.5, 0.55, 0.6, 1.0], 12),
    "pad_cone_d5": ("cone_d5_g2", [0.2, 0.6], 13),
}
# hourglass (src/dataset.py:127-152: per-path entry times, paths grouped by the length of their filled grids):
# name: (base case, requested times, seed, radii as fractions of r, paths put exactly on the boundary at the first time)
HOURGLASS = {
    "pad_hourglass_d5_early": ("hourglass_d5_g4_reentry", [0.1, 0.3, 0.45], 21, (0.0, 0.3), 0),
    "pad_hourglass_d5_late": ("hourglass_d5_g4_reentry", [0.62, 0.8, 1.0], 22, (0.05, 0.6), 0),
    "pad_hourglass_d5_late_onbdry": ("hourglass_d5_g4_reentry", [0.7, 0.75, 0.97], 23, (0.2, 0.68), 3),
}


def run(name, base, times, seed):
    case = G.load(base)
    p = dict(case["params"])
    over = {k: p[k] for k in p}
    over["domain"] = case["meta"].get("domain_class", "Hypercube")
    solver, funcs, params = rr.build(over, case["meta"]["funcs"], seed)
    with torch.no_grad():
        for q, w in zip(solver.u_net.parameters(), case["thu_list"]):
            q.copy_(torch.from_numpy(np.asarray(w)))
    d = p["dim"]
    g = torch.Generator().manual_seed(seed)
    N = 12
    scale = 0.6 if over["domain"] == "Hypercube" else 0.2      # well inside the cube / the cone at every requested time
    x = (torch.rand(N, d, generator=g) * 2 - 1) * scale
    t = torch.tensor(times, dtype=torch.float32)
    X = torch.cat((t.view(1, -1, 1).expand(N, -1, 1), x.unsqueeze(1).expand(-1, len(times), -1)), dim=2).contiguous()
    with torch.no_grad():
        u = solver.u_net(X)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), X=X.numpy(), u=u.numpy()[..., 0], base=np.array(base))
    print(name, "u[0] =", u[0, :, 0].numpy())


def run_hourglass(name, base, times, seed, radii, n_on):
    case = G.load(base)
    p = dict(case["params"])
    over = {k: p[k] for k in p}
    over["domain"] = case["meta"].get("domain_class")
    solver, funcs, params = rr.build(over, case["meta"]["funcs"], seed)
    with torch.no_grad():
        for q, w in zip(solver.u_net.parameters(), case["thu_list"]):
            q.copy_(torch.from_numpy(np.asarray(w)))
    net = solver.u_net.module if hasattr(solver.u_net, "module") else solver.u_net
    dom = net.domain
    d = p["dim"]
    g = torch.Generator().manual_seed(seed)
    N = 16
    dirs = torch.randn(N, d, generator=g, dtype=torch.float64)
    dirs /= dirs.norm(dim=1, keepdim=True)
    rad = (radii[0] + (radii[1] - radii[0]) * torch.rand(N, generator=g, dtype=torch.float64)) * dom.r
    x = dirs * rad.unsqueeze(1)
    for k in range(n_on):                                  # |x| = r t_0: func_w = 0 at the first requested time
        x[2 + 4 * k] = dirs[2 + 4 * k] * (dom.r * times[0])
    t = torch.tensor(times, dtype=torch.float64)
    X = torch.cat((t.view(1, -1, 1).expand(N, -1, 1), x.unsqueeze(1).expand(-1, len(times), -1)), dim=2).contiguous()
    path_i, idx, data = dom.bound_pad(X)
    with torch.no_grad():
        u = net(X)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), X=X.numpy(), u=u.numpy()[..., 0], base=np.array(base),
                        dom_times=dom.times.numpy(), n_groups=np.array(len(data)),
                        order=np.concatenate([q.numpy() for q in path_i]),
                        **{"grid%d" % k: q.numpy() for k, q in enumerate(data)},
                        **{"pos%d" % k: q.numpy() for k, q in enumerate(idx)})
    print(name, "groups", [len(q) for q in path_i], "grid lengths", [q.numel() for q in data], "u", tuple(u.shape))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "cube_cone"):
        for name, spec in CASES.items():
            run(name, *spec)
    if which in ("all", "hourglass"):
        for name, spec in HOURGLASS.items():
            run_hourglass(name, *spec)
