#!/bin/bash
# Development aid (GPU): time config 2 with alternative builds of the library (PDE_B200_LIB), e.g.
#   tools/ab_time.sh neural-network-based-pde-solver_b200/libpde_b200_*.so
for lib in "$@"; do
  echo "== $lib"
  PDE_B200_LIB=$(realpath "$lib") python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
sys.argv = ["quick_time"]
import importlib.util
spec = importlib.util.spec_from_file_location("qt", "tools/quick_time.py"); qt = importlib.util.module_from_spec(spec); spec.loader.exec_module(qt)
import torch
qt.run(3, "pinn", "FBC", 1 << 22, iters=5)
qt.run(5, "drm", "RB", 1 << 20, iters=5)
PY
done
