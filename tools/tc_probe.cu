// tc_probe.cu — development probe: checks the UMMA descriptor / TMEM layout assumptions of
// csrc/pde_tc_core.cuh on a real B200 with single 64x64x64 bf16 GEMMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe.bin tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../neural-network-based-pde-solver_b200/csrc/pde_tc_core.cuh"
#if PDE_TC_FP16
typedef __half op_t;
#define TO_OP(x) __float2half(x)
#define FROM_OP(x) __half2float(x)
#else
typedef __nv_bfloat16 op_t;
#define TO_OP(x) __float2bfloat16(x)
#define FROM_OP(x) __bfloat162float(x)
#endif

using namespace pde::tc;

// modes: 0 A K-major, B K-major | 1 A K-major, B MN-major | 2 A MN-major, B MN-major
//        3 = mode 0 but D at lane offset 16, column 64 | 4 = N=8 K-major B
//        5 = A MN-major, B MN-major N=16 slice at column offset 32 of the B tile (wgrad^T chunk)
//        6 = A K-major WITHOUT swizzle, one 2 KB block per K step: core matrix (8 rows x 16 B) of row group g, K half k at
//            256 g + 128 k (LBO 128, SBO 256); B as in mode 0
struct Args {
  const float* A;  // [64][64] logical "row-major as stored in the tile"
  const float* B;  // [64][64]
  float* D;        // [64][64] via assumed fragment mapping
  float* raw;      // [128][64] raw lanes via 32x32b
  int mode;
};

__global__ void __launch_bounds__(128, 1) probe(Args a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* tA = sm;
  unsigned char* tB = sm + TILE_BYTES;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // fill tiles: element (r, c) of the stored matrix
  for (int i = tid; i < 64 * 64; i += 128) {
    int r = i / 64, c = i % 64;
    op_t va = TO_OP(a.A[i]), vb = TO_OP(a.B[i]);
    *reinterpret_cast<op_t*>(tA + tile_off(r, c >> 3) + (c & 7) * 2) = va;
    *reinterpret_cast<op_t*>(tB + tile_off(r, c >> 3) + (c & 7) * 2) = vb;
  }
  if (a.mode == 6) {
    __syncthreads();
    for (int i = tid; i < 64 * 64; i += 128) {
      int r = i / 64, c = i % 64, ks = c >> 4, k = (c >> 3) & 1;
      *reinterpret_cast<op_t*>(tA + ks * 2048 + (r >> 3) * 256 + k * 128 + (r & 7) * 16 + (c & 7) * 2) = TO_OP(a.A[i]);
    }
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;

  // zero the TMEM region we will read raw (so garbage is visible as such): skip, just run MMA
  const int N = (a.mode == 4) ? 8 : (a.mode == 5 ? 16 : 64);
  const int lane_off = (a.mode == 3) ? 16 : 0;
  const int col_off = (a.mode == 3) ? 64 : 0;
  if (tid == 0) {
    const bool a_mn = (a.mode == 2 || a.mode == 5), b_mn = (a.mode == 1 || a.mode == 2 || a.mode == 5);
    const uint32_t idesc = make_idesc(64, N, a_mn, b_mn);
    const uint32_t d = taddr_of(tb, lane_off, col_off);
    for (int ks = 0; ks < 4; ++ks) {
      uint64_t ad = a_mn ? desc_mnmajor(smem_u32(tA), ks) : desc_kmajor(smem_u32(tA), ks);
      if (a.mode == 6) {   // no swizzle: layout type 0, LBO 128 (next core matrix along K), SBO 256 (next 8 rows)
        const uint32_t sa = smem_u32(tA) + ks * 2048;
        ad = static_cast<uint64_t>((sa & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(128 >> 4) << 16) | (static_cast<uint64_t>(256 >> 4) << 32) |
             (static_cast<uint64_t>(1) << 46);
      }
      uint64_t bd = b_mn ? desc_mnmajor(smem_u32(tB) + (a.mode == 5 ? 64 : 0), ks) : desc_kmajor(smem_u32(tB), ks);
      mma_bf16(d, ad, bd, idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();

  // fragment read: warp q reads its quarter's 16 rows
  {
    const int q = warp;
    for (int cb = 0; cb < N; cb += 8) {
      float v[4];
      tmem_ld_16x256b(taddr_of(tb, 32 * q + lane_off, col_off + cb), v);
      tmem_ld_wait();
      int r0 = 16 * q + lane / 4, c0 = cb + 2 * (lane % 4);
      a.D[r0 * 64 + c0] = v[0];
      a.D[r0 * 64 + c0 + 1] = v[1];
      a.D[(r0 + 8) * 64 + c0] = v[2];
      a.D[(r0 + 8) * 64 + c0 + 1] = v[3];
    }
  }
  // raw dump of all 128 lanes x 64 columns (from col_off)
  for (int cb = 0; cb < 64; cb += 8) {
    float v[8];
    tmem_ld_32x32b_x8(taddr_of(tb, 32 * warp, col_off + cb), v);
    tmem_ld_wait();
    for (int i = 0; i < 8; ++i) a.raw[(32 * warp + lane) * 64 + cb + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

static float bf(float x) { return FROM_OP(TO_OP(x)); }

int main() {
  std::vector<float> A(4096), B(4096);
  srand(1);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 500.0f;
  for (auto& x : B) x = (rand() % 2001 - 1000) / 500.0f;
  float *dA, *dB, *dD, *dR;
  cudaMalloc(&dA, 16384); cudaMalloc(&dB, 16384); cudaMalloc(&dD, 16384); cudaMalloc(&dR, 32768);
  cudaMemcpy(dA, A.data(), 16384, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), 16384, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES + 1024);
  int fails = 0;
  for (int mode = 0; mode <= 6; ++mode) {
    cudaMemset(dD, 0, 16384); cudaMemset(dR, 0, 32768);
    Args a{dA, dB, dD, dR, mode};
    probe<<<1, 128, 2 * TILE_BYTES + 1024>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
    std::vector<float> D(4096), R(8192);
    cudaMemcpy(D.data(), dD, 16384, cudaMemcpyDeviceToHost);
    cudaMemcpy(R.data(), dR, 32768, cudaMemcpyDeviceToHost);
    const int N = mode == 4 ? 8 : (mode == 5 ? 16 : 64);
    double maxerr = 0, maxref = 0;
    std::vector<double> ref(64 * 64, 0.0);
    for (int m = 0; m < 64; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < 64; ++k) {
          double av = (mode == 2 || mode == 5) ? bf(A[k * 64 + m]) : bf(A[m * 64 + k]);
          double bv = (mode == 5) ? bf(B[k * 64 + 32 + n]) : ((mode == 1 || mode == 2) ? bf(B[k * 64 + n]) : bf(B[n * 64 + k]));
          s += av * bv;
        }
        ref[m * 64 + n] = s;
        maxerr = fmax(maxerr, fabs(s - D[m * 64 + n]));
        maxref = fmax(maxref, fabs(s));
      }
    bool ok = maxerr < 1e-3 * maxref;
    printf("mode %d: max err %.3e (max ref %.3e) %s\n", mode, maxerr, maxref, ok ? "OK" : "FAIL");
    if (!ok) {
      ++fails;
      // where does ref[0][0..3], ref[1][0], ref[16][0], ref[8][0] show up in the raw dump?
      int probes[5][2] = {{0, 0}, {0, 1}, {1, 0}, {8, 0}, {16, 0}};
      for (auto& p : probes) {
        double want = ref[p[0] * 64 + p[1]];
        for (int l = 0; l < 128; ++l)
          for (int c = 0; c < 64; ++c)
            if (fabs(R[l * 64 + c] - want) < 1e-3 * fmax(1.0, fabs(want))) printf("  ref[%d][%d]=%.4f found at lane %d col %d\n", p[0], p[1], want, l, c);
      }
    }
  }
  printf(fails ? "PROBE FAILED (%d)\n" : "PROBE OK\n", fails);
  return fails ? 1 : 0;
}
