// pde_comm.cu — the exchange step of the data-parallel loss step as ONE kernel over NVLink peer
// memory.
//
// What is exchanged is tiny — [grad | dE | sums], 12 803 floats = 51 KB for config 2 — so the cost
// of the NCCL all-reduce is launch + protocol latency, not bandwidth.  pde_allreduce_oneshot does
// the whole exchange in a single thread block per GPU:
//
//   copy-in   : the rank's vector goes into its own peer-visible slot (double buffered by call parity)
//   barrier   : one system-scope release store of the call number into every peer's signal pad,
//               then an acquire spin until every peer's number has arrived in mine
//   reduce    : every rank pulls all slots over NVLink (NVSwitch: full bandwidth to every peer) and
//               adds them in rank order 0..W-1 — the same order on every rank, so the replicated
//               parameters stay bit-identical, and the result does not depend on arrival order
//
// No end-of-call barrier is needed: slot parity p is rewritten two calls later, and a rank can only
// pass the barrier of the call in between after every peer has finished reading parity p (it signals
// the next call number only from the kernel that follows on its stream).
// The call counter lives in device memory, so the kernel is CUDA-graph replayable.
//
// Peer buffers are plain cudaMalloc allocations shared with CUDA IPC handles (pde_peer_alloc /
// pde_peer_open); the handles travel through torch.distributed's object collectives once at set-up.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pde_b200.h"

namespace {

constexpr int SIGNAL_BYTES = 256;   // one uint32 per source rank, padded

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
template <typename T> __device__ __forceinline__ T ld_relaxed_sys(const T* p);
template <> __device__ __forceinline__ float ld_relaxed_sys<float>(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
template <> __device__ __forceinline__ double ld_relaxed_sys<double>(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

struct CommArgs {
  int rank, world;
  unsigned char* base[PDE_MAX_PEERS];   // peer-visible allocation of every rank: [signal pad | slot 0 | slot 1]
  long long n, slot_elems;
  void* buf;                            // local vector, reduced in place
  uint32_t* seq;                        // local: number of completed calls
  long long spin_limit;                 // clock64 ticks before giving up (a peer died): result is poisoned with NaN
};

template <typename T>
__global__ void __launch_bounds__(1024, 1) allreduce_oneshot_kernel(const CommArgs a) {
  __shared__ int failed;
  const uint32_t call = *a.seq + 1u;
  const int par = (int)(call & 1u);
  T* mine = reinterpret_cast<T*>(a.base[a.rank] + SIGNAL_BYTES) + (long long)par * a.slot_elems;
  T* buf = static_cast<T*>(a.buf);
  if (threadIdx.x == 0) failed = 0;
  for (long long i = threadIdx.x; i < a.n; i += blockDim.x) mine[i] = buf[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < a.world) {
    const int p = threadIdx.x;
    st_release_sys(reinterpret_cast<uint32_t*>(a.base[p]) + a.rank, call);
    const uint32_t* my_pad = reinterpret_cast<const uint32_t*>(a.base[a.rank]) + p;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(my_pad) - call) < 0) {
      if (clock64() - t0 > a.spin_limit) { failed = 1; break; }
    }
  }
  __syncthreads();
  const bool bad = failed != 0;
  for (long long i = threadIdx.x; i < a.n; i += blockDim.x) {
    T v = T(0);
    for (int p = 0; p < a.world; ++p) {
      const T* src = reinterpret_cast<const T*>(a.base[p] + SIGNAL_BYTES) + (long long)par * a.slot_elems;
      v += (p == a.rank) ? mine[i] : ld_relaxed_sys<T>(src + i);
    }
    buf[i] = bad ? (T)__longlong_as_double(0x7ff8000000000000ll) : v;
  }
  __syncthreads();
  if (threadIdx.x == 0) *a.seq = call;
}

}  // namespace

extern "C" {

int pde_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  if (!ptr || !handle64 || bytes == 0) return PDE_ERR_INVALID;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return PDE_ERR_CUDA; }
  if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaFree(p); cudaGetLastError(); return PDE_ERR_CUDA; }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); cudaGetLastError(); return PDE_ERR_CUDA; }
  static_assert(sizeof(h) == 64, "CUDA IPC handle size");
  memcpy(handle64, &h, 64);
  *ptr = p;
  return PDE_OK;
}

int pde_peer_open(const unsigned char* handle64, void** ptr) {
  if (!ptr || !handle64) return PDE_ERR_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return PDE_ERR_CUDA; }
  *ptr = p;
  return PDE_OK;
}

int pde_peer_close(void* ptr) {
  if (!ptr) return PDE_ERR_INVALID;
  if (cudaIpcCloseMemHandle(ptr) != cudaSuccess) { cudaGetLastError(); return PDE_ERR_CUDA; }
  return PDE_OK;
}

int pde_peer_free(void* ptr) {
  if (!ptr) return PDE_ERR_INVALID;
  if (cudaFree(ptr) != cudaSuccess) { cudaGetLastError(); return PDE_ERR_CUDA; }
  return PDE_OK;
}

int pde_peer_bytes(int32_t dtype, int64_t slot_elems, size_t* bytes) {
  if (!bytes || slot_elems < 1 || (dtype != PDE_F32 && dtype != PDE_F64)) return PDE_ERR_INVALID;
  const size_t es = dtype == PDE_F64 ? 8 : 4;
  *bytes = (size_t)SIGNAL_BYTES + 2 * (size_t)slot_elems * es;
  return PDE_OK;
}

int pde_allreduce_oneshot(const pde_peers* peers, int32_t dtype, void* buf, int64_t n, int64_t slot_elems, void* seq,
                          void* stream) {
  if (!peers || !buf || !seq) return PDE_ERR_INVALID;
  if (dtype != PDE_F32 && dtype != PDE_F64) return PDE_ERR_INVALID;
  if (peers->world < 1 || peers->world > PDE_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->world) return PDE_ERR_INVALID;
  if (n < 1 || n > slot_elems) return PDE_ERR_INVALID;
  CommArgs a;
  a.rank = peers->rank; a.world = peers->world;
  for (int p = 0; p < PDE_MAX_PEERS; ++p) {
    a.base[p] = p < peers->world ? static_cast<unsigned char*>(peers->base[p]) : nullptr;
    if (p < peers->world && !a.base[p]) return PDE_ERR_INVALID;
  }
  a.n = n; a.slot_elems = slot_elems; a.buf = buf; a.seq = static_cast<uint32_t*>(seq);
  a.spin_limit = 4000000000ll;   // ~2 s at 1.9 GHz
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == PDE_F32) allreduce_oneshot_kernel<float><<<1, 1024, 0, st>>>(a);
  else allreduce_oneshot_kernel<double><<<1, 1024, 0, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? PDE_OK : PDE_ERR_CUDA;
}

}  // extern "C"
