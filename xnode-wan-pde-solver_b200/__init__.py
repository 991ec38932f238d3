"""xnode-wan-pde-solver_b200 -- B200-native (sm_100a) implementation of the XNODE-WAN hot path:
the per-iteration Monte-Carlo weak-form loss and its parameter gradients, behind the reference's
own Python API (NODE_WAN_solver / NeuralODE / discriminator / loss / Comb_loader / func_eval).

The directory name carries a hyphen (it mirrors the upstream repository name); import it with
`importlib.import_module("xnode-wan-pde-solver_b200")` or through the alias package `xnode_wan_b200`.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
