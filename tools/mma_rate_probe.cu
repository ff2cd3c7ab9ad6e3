// mma_rate_probe.cu — development probe (GPU): how long do back-to-back tcgen05.mma instructions of the shapes the
// loss kernel issues take when both operands come from shared memory, and is an M = 128 instruction whose A operand
// stacks two 64-row tiles (8-row groups at a uniform 1 KB stride) as cheap as an M = 64 one?  One CTA per SM, all
// 148 SMs busy, cycles measured with clock64 from the first issue to the arrival of the commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mma_rate_probe.bin tools/mma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../neural-network-based-pde-solver_b200/csrc/pde_tc_core.cuh"

using namespace pde::tc;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_k(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}

struct Args {
  long long* cycles;   // [grid][NV]
  long long* issue;    // [grid][NV] cycles until the last instruction had been issued
  float* raw;          // [128][64] raw lanes of the layout check (CTA 0)
  const float* A;      // [128][64]
  const float* B;      // [64][64]
  int reps;
};
constexpr int NV = 8;
constexpr int NT = 14;   // tiles of 8 KB

__global__ void __launch_bounds__(128, 1) probe(Args a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // tiles 0,1: A (128 rows, K-major, groups of 8 rows 1 KB apart); tile 2: B; the rest: small numbers
  for (int i = tid; i < NT * TILE_BYTES / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x2C002C00u;   // 2^-4
  __syncthreads();
  for (int i = tid; i < 128 * 64; i += 128) {
    int r = i / 64, c = i % 64;
    *reinterpret_cast<__half*>(sm + (r >> 6) * TILE_BYTES + tile_off(r & 63, c >> 3) + (c & 7) * 2) = __float2half(a.A[i]);
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    int r = i / 64, c = i % 64;
    *reinterpret_cast<__half*>(sm + 2 * TILE_BYTES + tile_off(r, c >> 3) + (c & 7) * 2) = __float2half(a.B[i]);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const uint32_t s0 = smem_u32(sm);
  uint32_t ph = 0;
  auto tile = [&](int t) { return s0 + t * TILE_BYTES; };

  // ---- layout check: D[128][64] = A[128][64] B^T, one M = 128 instruction per K step
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, 64, 0, 0);
    for (int ks = 0; ks < 4; ++ks) mma_bf16(tb, desc_kmajor(tile(0), ks), desc_kmajor(tile(2), ks), idesc, ks > 0);
    mma_commit(&bar);
  }
  mbar_wait(&bar, ph); ph ^= 1;
  tc_fence_after();
  if (blockIdx.x == 0) {
    for (int cb = 0; cb < 64; cb += 8) {
      float v[8];
      tmem_ld_32x32b_x8(taddr_of(tb, 32 * warp, cb), v);
      tmem_ld_wait();
      for (int i = 0; i < 8; ++i) a.raw[(32 * warp + lane) * 64 + cb + i] = v[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- rates: the issue loop is written like the kernel's (one warp, elect.sync, descriptor words = warp-uniform base +
  // compile-time offsets, so they live in uniform registers)
  constexpr uint32_t TD = TILE_BYTES >> 4;
  constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024, version 1, SWIZZLE_128B
  const uint32_t tb0 = __shfl_sync(0xffffffffu, tb, 0);
  const uint32_t kK = __shfl_sync(0xffffffffu, (s0 >> 4) | ((16u >> 4) << 16), 0), kM = __shfl_sync(0xffffffffu, (s0 >> 4) | ((8192u >> 4) << 16), 0);
  for (int v = 0; v < NV; ++v) {
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
      constexpr uint32_t i64 = make_idesc(64, 64, 0, 0), i128 = make_idesc(128, 64, 0, 0), i64n128 = make_idesc(64, 128, 0, 0);
      constexpr uint32_t w64 = make_idesc(64, 64, 1, 1), w128 = make_idesc(128, 64, 1, 1), s8 = make_idesc(64, 8, 1, 0);
      t0 = clock64();
      for (int r = 0; r < a.reps; ++r) {
        if (elect_one()) {
          if (v == 0) {          // today's forward K step: 5 channels x 3 terms, M = 64
#pragma unroll
            for (int c = 0; c < 5; ++c) {
              const uint32_t d = tb0 + ((16 * (c & 1)) << 16) + 64 * (c >> 1);
              mma_k(d, kK + 2 * c * TD, HI, kK + 10 * TD, HI, i64, 1u);
              mma_k(d, kK + (2 * c + 1) * TD, HI, kK + 10 * TD, HI, i64, 1u);
              mma_k(d, kK + 2 * c * TD, HI, kK + 11 * TD, HI, i64, 1u);
            }
          } else if (v == 1) {   // channel pairs stacked along M: 2 x 3 M = 128 + 3 M = 64
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              const uint32_t d = tb0 + 64 * p;
              mma_k(d, kK + 4 * p * TD, HI, kK + 10 * TD, HI, i128, 1u);
              mma_k(d, kK + (4 * p + 2) * TD, HI, kK + 10 * TD, HI, i128, 1u);
              mma_k(d, kK + 4 * p * TD, HI, kK + 11 * TD, HI, i128, 1u);
            }
            const uint32_t d = tb0 + 128;
            mma_k(d, kK + 8 * TD, HI, kK + 10 * TD, HI, i64, 1u);
            mma_k(d, kK + 9 * TD, HI, kK + 10 * TD, HI, i64, 1u);
            mma_k(d, kK + 8 * TD, HI, kK + 11 * TD, HI, i64, 1u);
          } else if (v == 2) {   // 15 M = 128 instructions
#pragma unroll
            for (int c = 0; c < 15; ++c) mma_k(tb0 + 64 * (c % 3), kK + ((2 * c) % 10) * TD, HI, kK + (10 + (c & 1)) * TD, HI, i128, 1u);
          } else if (v == 3) {   // wgrad: both MN-major, M = 64
#pragma unroll
            for (int c = 0; c < 15; ++c) mma_k(tb0 + 384, kM + (c % 10) * TD, HI, kM + ((c + 3) % 10) * TD, HI, w64, 1u);
          } else if (v == 4) {   // wgrad with (hi, lo) stacked along M = 128 (tiles 8 KB apart = LBO of the MN-major view)
#pragma unroll
            for (int c = 0; c < 15; ++c) mma_k(tb0 + 384, kM + (2 * (c % 5)) * TD, HI, kM + ((c + 3) % 10) * TD, HI, w128, 1u);
          } else if (v == 5) {   // first layer / bias: M = 64, N = 8, A MN-major
#pragma unroll
            for (int c = 0; c < 15; ++c) mma_k(tb0 + 448, kM + (c % 10) * TD, HI, kK + 12 * TD, HI, s8, 1u);
          } else if (v == 6) {   // same A tile every time
#pragma unroll
            for (int c = 0; c < 15; ++c) mma_k(tb0 + 64 * (c % 3), kK, HI, kK + 10 * TD, HI, i64, 1u);
          } else {               // M = 64, N = 128
#pragma unroll
            for (int c = 0; c < 15; ++c) mma_k(tb0 + 128 * (c & 1), kK + (c % 10) * TD, HI, kK + 10 * TD, HI, i64n128, 1u);
          }
        }
        __syncwarp();
      }
      if (elect_one()) mma_commit(&bar);
      __syncwarp();
      if (lane == 0) a.issue[blockIdx.x * NV + v] = clock64() - t0;
    }
    mbar_wait(&bar, ph); ph ^= 1;
    if (tid == 0) {
      t1 = clock64();
      a.cycles[blockIdx.x * NV + v] = t1 - t0;
    }
    tc_fence_after();
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  std::vector<float> A(128 * 64), B(64 * 64);
  srand(1);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 500.0f;
  for (auto& x : B) x = (rand() % 2001 - 1000) / 500.0f;
  float *dA, *dB, *dR;
  long long *dC, *dI;
  const int grid = 148, reps = 256;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dR, 128 * 64 * 4); cudaMalloc(&dC, grid * NV * 8); cudaMalloc(&dI, grid * NV * 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = NT * TILE_BYTES + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Args a{dC, dI, dR, dA, dB, reps};
  for (int it = 0; it < 2; ++it) {
    probe<<<grid, 128, smem>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
  }
  std::vector<float> R(128 * 64);
  std::vector<long long> C(grid * NV), I(grid * NV);
  cudaMemcpy(I.data(), dI, I.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(R.data(), dR, R.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
  auto h = [](float x) { return __half2float(__float2half(x)); };
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      double s = 0;
      for (int k = 0; k < 64; ++k) s += (double)h(A[m * 64 + k]) * h(B[n * 64 + k]);
      maxerr = fmax(maxerr, fabs(s - R[m * 64 + n]));
      maxref = fmax(maxref, fabs(s));
    }
  printf("M=128 layout (lane = row of the stacked A): max err %.3e (max ref %.3e) %s\n", maxerr, maxref, maxerr < 1e-3 * maxref ? "OK" : "FAIL");
  const char* names[NV] = {"fwd K step today: 15 x (M64,N64) K-major", "paired: 6 x (M128,N64) + 3 x (M64,N64)", "15 x (M128,N64) K-major",
                           "wgrad: 15 x (M64,N64) MN/MN", "wgrad stacked: 15 x (M128,N64) MN/MN", "15 x (M64,N8) A MN-major",
                           "15 x (M64,N64) same A tile", "15 x (M64,N128)"};
  for (int v = 0; v < NV; ++v) {
    double s = 0, mx = 0, si = 0;
    const int n = v == 1 ? 9 : 15;
    for (int b = 0; b < grid; ++b) { s += C[b * NV + v]; si += I[b * NV + v]; mx = fmax(mx, (double)C[b * NV + v]); }
    printf("%-45s %8.1f cycles per batch (max over CTAs %8.1f, issue alone %8.1f), %6.1f per instruction\n", names[v], s / grid / reps, mx / reps,
           si / grid / reps, s / grid / reps / n);
  }
  return 0;
}
