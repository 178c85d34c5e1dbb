"""Import shim: makes the folder `trixiparticles.jl_b200/` importable as the dotted module
`trixiparticles.jl_b200` (a directory name with a dot cannot be a regular package)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "trixiparticles.jl_b200")
_name = __name__ + ".jl_b200"
if _name not in _sys.modules:
    _spec = _ilu.spec_from_file_location(_name, _os.path.join(_real, "__init__.py"),
                                         submodule_search_locations=[_real])
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = _sys.modules[_name]
