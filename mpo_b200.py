"""Import alias for the hyphenated package directory `multimodal-path-omic_b200/`.

`import mpo_b200` returns the package itself (same module object), so `mpo_b200.mcat`, `mpo_b200.bagpass`, ...
are the real submodules and relative imports inside the package keep working.
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("multimodal-path-omic_b200")
sys.modules[__name__] = _pkg
