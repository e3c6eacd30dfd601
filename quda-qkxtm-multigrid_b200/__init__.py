"""B200-native twisted-mass Dslash + multigrid engine behind the QUDA C API.

The directory name follows the reference repository and is not a Python identifier; load it with
    import importlib; q = importlib.import_module("quda-qkxtm-multigrid_b200")
or through the `quda_b200.py` shim at the repository root.
"""
from .api import *  # noqa: F401,F403
from .api import lib, LIB_PATH, EXPORTS  # noqa: F401
