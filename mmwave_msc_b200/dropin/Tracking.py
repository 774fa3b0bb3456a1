"""`import Tracking` of the reference's scripts (offline_main.py, main.py, preprocessing.py) resolves here when this
directory is first on sys.path: the name is bound to the GPU package's module of the same name (one module object,
so edits to its attributes are seen by both)."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(1, _root)
import mmwave_msc_b200.Tracking as _m  # noqa: E402

sys.modules[__name__] = _m
