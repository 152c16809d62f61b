"""Import shim: `import pdmpflux_b200` loads the package in ./pdmpflux.jl_b200/ (whose directory name, fixed
by the repository layout, is not a valid Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pdmpflux.jl_b200")
_spec = importlib.util.spec_from_file_location("pdmpflux_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pdmpflux_b200"] = _mod
_spec.loader.exec_module(_mod)
