"""importable alias of the package directory `xnode-wan-pde-solver_b200/` (hyphens are not valid in
`import` statements).  `import xnode_wan_b200` returns that package object itself."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("xnode-wan-pde-solver_b200")
sys.modules[__name__] = _pkg
