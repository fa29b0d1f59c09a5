"""Import shim: the package directory is literally ``diffopt.jl_b200/`` (not importable by name)."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "diffopt.jl_b200")
_name = "diffopt_jl_b200"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_pkg, "__init__.py"),
                                                   submodule_search_locations=[_pkg])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
pkg = sys.modules[_name]


def submodule(name):
    import importlib
    return importlib.import_module(f"{_name}.{name}")


Context = pkg.Context
SingularException = pkg.SingularException
DiffOptB200Error = pkg.DiffOptB200Error
