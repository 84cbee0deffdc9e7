"""Import shim: the package directory is named ``dojo.jl_b200`` (the name the
build contract fixes), which is not a valid Python identifier.  ``import
dojo_jl_b200`` loads that directory as a regular package under this name."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "dojo.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "dojo_jl_b200", os.path.join(_pkg_dir, "__init__.py"),
    submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dojo_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
