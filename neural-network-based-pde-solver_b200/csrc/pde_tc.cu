// pde_tc.cu — tcgen05 path (placeholder until the tensor-core kernel lands: declines every shape).
#include "pde_tc.h"
namespace pde {
bool tc_supported(const pde_net*, const pde_program*, long long) { return false; }
int tc_workspace_bytes(const pde_net*, int, long long, size_t*) { return PDE_ERR_UNSUPPORTED; }
int tc_residual_loss_grad(const pde_net*, const pde_envelope*, const pde_program*, const void*, long long, const void*,
                          double, void*, void*, void*, void*, size_t, cudaStream_t) {
  return PDE_ERR_UNSUPPORTED;
}
}  // namespace pde
