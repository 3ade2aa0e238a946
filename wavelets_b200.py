"""Import shim: `import wavelets_b200` loads the package that lives in ./wavelets.jl_b200/.

The package directory is named after the reference (Wavelets.jl -> wavelets.jl_b200), and a dot
is not a legal character in a Python module name, so this file loads it by path and installs it
in sys.modules under the importable name `wavelets_b200`.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "wavelets.jl_b200")
_spec = _ilu.spec_from_file_location(
    "wavelets_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["wavelets_b200"] = _mod
_spec.loader.exec_module(_mod)
