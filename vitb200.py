"""Import shim: `import vitb200` == the package in ./vit-of-pytorch_b200 (hyphenated directory)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("vit-of-pytorch_b200")
sys.modules[__name__] = _pkg
