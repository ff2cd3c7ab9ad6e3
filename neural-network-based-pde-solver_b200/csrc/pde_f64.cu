// fp64 instantiations of the generic collocation kernels (validation path, 1e-10 parity).
#include "pde_inst.cuh"
namespace pde { PDE_INSTANTIATE(double) }
