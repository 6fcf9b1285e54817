"""Import shim: the package directory is ``happy-simulator_b200/`` (not a valid
Python identifier), so this module loads it under the importable name
``happysim_b200``.  ``import happysim_b200`` then behaves like a normal package
import (sub-modules such as ``happysim_b200.lowering`` resolve inside that
directory)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "happy-simulator_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
