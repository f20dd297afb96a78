"""Import shim: the package directory is named ``helix-db_b200`` (not a Python identifier).

``import helix_db_b200`` loads ``helix-db_b200/__init__.py`` under this module name.
"""
import importlib.util as _u
import sys as _sys
from pathlib import Path as _P

_pkg = _P(__file__).resolve().parent / "helix-db_b200"
_spec = _u.spec_from_file_location("helix_db_b200", _pkg / "__init__.py", submodule_search_locations=[str(_pkg)])
_mod = _u.module_from_spec(_spec)
_sys.modules["helix_db_b200"] = _mod
_spec.loader.exec_module(_mod)
