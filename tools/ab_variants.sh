#!/bin/bash
# Development aid: time configs 2 and 3 (tools/quick_time.py lines) for each libpde_b200_<tag>.so given as argument.
for v in "$@"; do
  export PDE_B200_LIB=$PWD/neural-network-based-pde-solver_b200/libpde_b200_$v.so
  echo "== $v"; python tools/quick_time.py 2>&1 | grep -E "N=4194304|drm"
done
