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


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name, spec in CASES.items():
        run(name, *spec)
