"""Importable alias of the package directory named after the reference repository."""
import importlib as _il
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)
PACKAGE_DIR_NAME = "ntire-2026-light-field-image-super-resolution-challenge---track-2-efficiency_b200"
_pkg = _il.import_module(PACKAGE_DIR_NAME)
_sys.modules[__name__] = _pkg
