#!/bin/bash
# Development aid: build libpde_b200_<tag>.so with extra -D flags for pde_tc.cu only (headline instantiation only, fast),
# reusing the other objects of the regular build.  Usage: tools/build_variant.sh <tag> [-DFLAG ...]
# Select it at run time with PDE_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../neural-network-based-pde-solver_b200/csrc"
tag=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
     -DPDE_TC_ONLY_CFG2 "$@" -c pde_tc.cu -o /tmp/pde_tc_$tag.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpde_b200_$tag.so pde_abi.o pde_f32.o pde_f64.o /tmp/pde_tc_$tag.o pde_step.o pde_comm.o -lcudart
echo built ../libpde_b200_$tag.so
