"""Importable alias of the package directory `realtime-deformations_b200/` (hyphen is not a valid identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("realtime-deformations_b200")
build, capi, scenes = _pkg.build, _pkg.capi, _pkg.scenes
Sim, MpmError = capi.Sim, capi.MpmError
