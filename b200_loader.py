"""Import the package directory `interpolations.jl_b200/` (its name contains a dot, so the plain
`import` statement cannot reach it) under the module name `interpolations_jl_b200`."""
import importlib.util
import os
import sys

_NAME = "interpolations_jl_b200"
_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(_ROOT, "interpolations.jl_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG, "__init__.py"),
                                                  submodule_search_locations=[_PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
