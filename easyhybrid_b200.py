"""Importable alias for the package directory ``easyhybrid.jl_b200`` (a dot is not valid in
a Python package name): ``import easyhybrid_b200`` loads that directory as a package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "easyhybrid.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "easyhybrid_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["easyhybrid_b200"] = _mod
_spec.loader.exec_module(_mod)
