"""Import helper: the package directory is literally ``gridapdistributed.jl_b200`` (a dotted name
is not a Python identifier), so it is registered under the importable alias
``gridapdistributed_jl_b200``."""
import importlib.util
import os
import sys

ALIAS = "gridapdistributed_jl_b200"
ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "gridapdistributed.jl_b200")


def load():
    if ALIAS in sys.modules:
        return sys.modules[ALIAS]
    spec = importlib.util.spec_from_file_location(
        ALIAS, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[ALIAS] = mod
    spec.loader.exec_module(mod)
    return mod
