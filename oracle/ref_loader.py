"""TEST / BENCH INFRASTRUCTURE ONLY — loads the reference's own scripts from ``oracle/_ref/``.

``__graft_entry__.build()`` copies a handful of the reference's scripts (unmodified) from
``/root/reference`` into the git-ignored ``oracle/_ref/`` when that checkout is present, so that they
travel to the GPU box with the repository snapshot.  ``bench.py`` times them as the reference arm
(``--impl reference``, ``cpu_baseline.kind = "reference"``) and falls back to the torch port
``oracle/autograd_ref.py`` (``kind = "port"``) when the directory is absent.  Nothing on the product
path imports this module.

The scripts are standalone programs: importing them needs a stub ``matplotlib`` (absent from this image),
creates ``results/`` folders in the current directory and seeds the global RNG (KH_1D.py:15-18,
QHO_2D.py:12-23), so they are loaded from a temporary working directory.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
# file name in oracle/_ref  ->  path inside the reference checkout
SCRIPTS = {
    "Poisson_ND.py": "Poisson_Equations/Poisson_ND.py",
    "QHO_2D.py": "Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_2D.py",
    "IPW_1D_WAN.py": "Schrodinger_Equations/Infinite_Potential_Well/IPW_1D_WAN.py",
}
_CACHE = {}


def available(name="Poisson_ND.py"):
    return os.path.exists(os.path.join(REF_DIR, name))


def load(name="Poisson_ND.py"):
    """The reference script ``name`` as a module, or None when oracle/_ref does not hold it."""
    if name in _CACHE:
        return _CACHE[name]
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        return None
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].use = lambda *a, **k: None
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        import contextlib
        import torch
        state = torch.random.get_rng_state()
        spec = importlib.util.spec_from_file_location("pde_ref_" + name[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        with contextlib.redirect_stdout(sys.stderr):     # the scripts print at import time
            spec.loader.exec_module(mod)
        torch.random.set_rng_state(state)
    finally:
        os.chdir(cwd)
    _CACHE[name] = mod
    return mod
