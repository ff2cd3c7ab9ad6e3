// fp32 instantiations of the generic collocation kernels.
#include "pde_inst.cuh"
namespace pde { PDE_INSTANTIATE(float) }
