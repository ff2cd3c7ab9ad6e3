// pde_tc.h — Blackwell tensor-core (tcgen05 / TMEM) path for the headline network shapes.
// Returns PDE_ERR_UNSUPPORTED for anything it does not cover; the caller then uses the
// generic SIMT kernel.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/pde_b200.h"

namespace pde {
bool tc_supported(const pde_net* net, const pde_program* prog, long long n_points);
int tc_workspace_bytes(const pde_net* net, int order, long long n_points, size_t* bytes);
int tc_residual_loss_grad(const pde_net* net, const pde_envelope* env, const pde_program* prog, const void* X,
                          long long n_points, const void* seed, double inv_n, void* sums, void* grad,
                          void* energy_grad, void* workspace, size_t workspace_bytes, cudaStream_t stream);
}  // namespace pde
