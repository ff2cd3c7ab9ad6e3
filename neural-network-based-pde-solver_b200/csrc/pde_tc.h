// pde_tc.h — Blackwell tensor-core (tcgen05 / TMEM) path for the headline network shapes.
// Returns PDE_ERR_UNSUPPORTED for anything it does not cover; the caller then uses the
// generic SIMT kernel.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/pde_b200.h"

namespace pde {
// exchange fused into the reduction of the per-CTA partials (pde_residual_loss_grad_exchange); null: none
struct ExchangeReq {
  const pde_peers* peers;
  int64_t slot_elems;
  void* seq;
};
bool tc_supported(const pde_net* net, const pde_program* prog, long long n_points);
int tc_workspace_bytes(const pde_net* net, int order, long long n_points, size_t* bytes);
int tc_residual_loss_grad(const pde_net* net, const pde_envelope* env, const pde_program* prog, const void* X,
                          long long n_points, const void* seed, double inv_n, void* sums, void* grad,
                          void* energy_grad, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                          const ExchangeReq* ex = nullptr);
// network jets (orders 0 / 1) and their reverse sweep on the same kernel
bool tc_jets_supported(const pde_net* net, int order, long long n_points);
int tc_jets_forward(const pde_net* net, int order, const void* X, long long n_points, void* J, void* workspace,
                    size_t workspace_bytes, cudaStream_t stream);
int tc_jets_backward(const pde_net* net, int order, const void* X, long long n_points, const void* Jbar, void* grad,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);
// kernel-family override: -1 automatic, 0 generic SIMT kernel, 1 tensor-core kernel wherever its shapes allow
void tc_set_path_override(int v);
int tc_get_path_override();
}  // namespace pde
