"""Import shim: `import quda_b200` == the package in ./quda-qkxtm-multigrid_b200 (hyphenated name)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("quda-qkxtm-multigrid_b200")
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
api = importlib.import_module("quda-qkxtm-multigrid_b200.api")
