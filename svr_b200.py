"""Import shim: ``import svr_b200`` == the package in ``single-view-3d-reconstruction_b200/``
(whose directory name, fixed by the project layout, is not a valid Python identifier)."""
import importlib
import sys
from pathlib import Path

_root = str(Path(__file__).resolve().parent)
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("single-view-3d-reconstruction_b200")
sys.modules[__name__] = _pkg
