"""Import alias: ``import pde_b200`` loads the package that lives in the (non-identifier)
directory ``neural-network-based-pde-solver_b200/`` next to this file."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "neural-network-based-pde-solver_b200")
_spec = importlib.util.spec_from_file_location("pde_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pde_b200"] = _mod
_spec.loader.exec_module(_mod)
