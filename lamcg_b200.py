"""Import shim: the package directory `2024-eumaster4hpc-student-challenge_b200/` is not a valid
Python identifier, so `import lamcg_b200` re-exports it."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("2024-eumaster4hpc-student-challenge_b200")
globals().update({k: getattr(_pkg, k) for k in dir(_pkg) if not k.startswith("__")})
package = _pkg
