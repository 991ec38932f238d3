"""xnode-wan-pde-solver_b200 -- B200-native (sm_100a) implementation of the XNODE-WAN hot path:
the per-iteration Monte-Carlo weak-form loss and its parameter gradients, behind the reference's
own Python API (NODE_WAN_solver / NeuralODE / discriminator / loss / Comb_loader / func_eval).

The directory name carries a hyphen (it mirrors the upstream repository name); import it with
`importlib.import_module("xnode-wan-pde-solver_b200")` or through the alias package `xnode_wan_b200`.
"""
from . import _lib, hotpath, problems  # noqa: F401
from .aux import L_norm, rel_err  # noqa: F401
from .dataset import Comb_loader, Hypercube, NSphere_TCone, NSphere_THourglass  # noqa: F401
from .loss import CoefA, CoefB, CoefC, loss  # noqa: F401
from .paths import CollapsedPaths  # noqa: F401
from .model import LazyPrediction, NeuralODE, discriminator, init_weights  # noqa: F401
from .training import NODE_WAN_solver, func_eval  # noqa: F401
