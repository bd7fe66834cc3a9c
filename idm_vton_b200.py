"""Importable alias for the `idm-vton_b200/` package directory (a hyphen is not a valid module name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "idm-vton_b200")
_spec = importlib.util.spec_from_file_location(
    "idm_vton_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["idm_vton_b200"] = _mod
_spec.loader.exec_module(_mod)
