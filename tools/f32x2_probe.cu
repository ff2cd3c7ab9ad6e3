// f32x2_probe.cu — does packed fma.rn.f32x2 give more FP32 throughput than scalar FFMA on sm_100a?
#include <cstdio>
#include <cstdint>
__global__ void scalar_k(float* out, float a, float b, int iters) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
    x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t pk(float lo, float hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__global__ void packed_k(float* out, float a, float b, int iters) {
  float t = threadIdx.x;
  uint64_t x0 = pk(t, t + 1), x1 = pk(t + 2, t + 3), x2 = pk(t + 4, t + 5), x3 = pk(t + 6, t + 7);
  uint64_t x4 = pk(t + 8, t + 9), x5 = pk(t + 10, t + 11), x6 = pk(t + 12, t + 13), x7 = pk(t + 14, t + 15);
  const uint64_t A = pk(a, a), B = pk(b, b);
  for (int i = 0; i < iters; ++i) {
    x0 = fma2(x0, A, B); x1 = fma2(x1, A, B); x2 = fma2(x2, A, B); x3 = fma2(x3, A, B);
    x4 = fma2(x4, A, B); x5 = fma2(x5, A, B); x6 = fma2(x6, A, B); x7 = fma2(x7, A, B);
  }
  uint64_t s = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((uint32_t)s) + __uint_as_float((uint32_t)(s >> 32));
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 100000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); scalar_k<<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 148.0 * 8 * 256 * iters * 8 * 2;
    printf("scalar: %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
    cudaEventRecord(e0); packed_k<<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    fl = 148.0 * 8 * 256 * iters * 16 * 2;
    printf("packed: %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
