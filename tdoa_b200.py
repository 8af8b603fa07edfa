"""Import shim: `import tdoa_b200` gives the package in ./tdoa-geolocation_b200/
(whose directory name is not a Python identifier)."""
import importlib
import sys
from pathlib import Path

_root = str(Path(__file__).resolve().parent)
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("tdoa-geolocation_b200")
sys.modules[__name__] = _pkg
