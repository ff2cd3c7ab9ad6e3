// pde_launch.h — host-side launch shims between the C ABI (pde_abi.cu) and the per-dtype
// instantiation units (pde_f32.cu / pde_f64.cu).
#pragma once
#include <cuda_runtime.h>
#include "pde_simt.cuh"

namespace pde {

// Process-wide counters the host side reads through the C ABI (pde_launch_count, pde_last_kernel_path): every kernel
// this library enqueues is counted where it is launched, and the fused calls record which kernel family served them.
void count_launch(int k = 1);
void set_last_path(int path);   // 0 generic SIMT kernel, 1 tcgen05 kernel

struct KernelInfo {
  const void* fn;
  int regs;
};

// kernel handle for (dim, order); nullptr if out of range
template <typename T> KernelInfo net_kernel_info(int dim, int order);
template <typename T> cudaError_t launch_net(int dim, int order, int grid, int block, size_t smem, cudaStream_t st, const KArgs<T>& a);
template <typename T> cudaError_t launch_pack(cudaStream_t st, const PackArgs<T>& a);
template <typename T> cudaError_t launch_reduce(cudaStream_t st, const ReduceArgs<T>& a);

// WAN elementwise coupling on jets (pde_wan in include/pde_b200.h)
template <typename T>
struct WanArgs {
  int D;
  long long n;
  const T* X; const T* Ju; const T* Jv;
  const T* f; const T* beta; const T* energy; const T* seed;
  T alpha, beta_const, energy_const, w_lo, w_hi, eps_den, inv_n;
  EnvDev<T> env_u, env_v;
  T* Jbar_u; T* Jbar_v;
  double* psums;   // [blocks][8]
  T* sums;         // (5) out, written by the finishing kernel
  int blocks;
};
template <typename T> cudaError_t launch_wan(cudaStream_t st, const WanArgs<T>& a);

}  // namespace pde
