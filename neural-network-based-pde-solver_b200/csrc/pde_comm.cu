// pde_comm.cu — the exchange step of the data-parallel loss step as ONE kernel over NVLink peer
// memory.
//
// What is exchanged is tiny — [grad | dE | sums], 12 803 floats = 51 KB for config 2 — so the cost
// of the NCCL all-reduce is launch + protocol latency, not bandwidth.  pde_allreduce_oneshot does
// the whole exchange in one launch of a few independent thread blocks per GPU:
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

constexpr int SIGNAL_BYTES = 512;   // one uint32 per (source rank, block), then the local block-completion counter

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

__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
template <typename T> struct Vec16;            // 16-byte groups of T
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void add(uint4& acc, const uint4& v) {
    acc.x = __float_as_uint(__uint_as_float(acc.x) + __uint_as_float(v.x)); acc.y = __float_as_uint(__uint_as_float(acc.y) + __uint_as_float(v.y));
    acc.z = __float_as_uint(__uint_as_float(acc.z) + __uint_as_float(v.z)); acc.w = __float_as_uint(__uint_as_float(acc.w) + __uint_as_float(v.w));
  }
};
template <> struct Vec16<double> {
  static constexpr int N = 2;
  static __device__ __forceinline__ void add(uint4& acc, const uint4& v) {
    const double a0 = __hiloint2double(acc.y, acc.x) + __hiloint2double(v.y, v.x), a1 = __hiloint2double(acc.w, acc.z) + __hiloint2double(v.w, v.z);
    acc.x = __double2loint(a0); acc.y = __double2hiint(a0); acc.z = __double2loint(a1); acc.w = __double2hiint(a1);
  }
};

// G independent blocks of 512 threads, block b owning the b-th segment of the vector and its own signal
// word per source rank, so no grid-wide barrier is needed.  The peers' slots are pulled as 16-byte groups,
// U groups per thread, with every load issued before the first add: the pull costs one NVLink round trip,
// not one per element.
constexpr int COMM_THREADS = 512, COMM_MAX_BLOCKS = 8, COMM_U = 2;

template <typename T>
__global__ void __launch_bounds__(COMM_THREADS, 1) allreduce_oneshot_kernel(const CommArgs a) {
  constexpr int U = COMM_U, VN = Vec16<T>::N;
  __shared__ int failed;
  const uint32_t call = *a.seq + 1u;
  const int par = (int)(call & 1u);
  const int G = gridDim.x, b = blockIdx.x;
  T* mine = reinterpret_cast<T*>(a.base[a.rank] + SIGNAL_BYTES) + (long long)par * a.slot_elems;
  T* buf = static_cast<T*>(a.buf);
  const bool vec_ok = (reinterpret_cast<uintptr_t>(buf) & 15) == 0;   // slots are 16-byte aligned by construction
  const long long nv = vec_ok ? a.n / VN : 0;                          // 16-byte groups handled by the vector path
  const long long per = (nv + G - 1) / G;
  const long long v0 = b * per, v1 = (v0 + per < nv) ? v0 + per : nv;   // this block's groups
  const bool tail = (b == 0);                                          // block 0 also owns the scalar tail
  if (threadIdx.x == 0) failed = 0;
  for (long long i = v0 + threadIdx.x; i < v1; i += blockDim.x) reinterpret_cast<uint4*>(mine)[i] = reinterpret_cast<const uint4*>(buf)[i];
  if (tail)
    for (long long i = nv * VN + threadIdx.x; i < a.n; i += blockDim.x) mine[i] = buf[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < a.world) {
    const int p = threadIdx.x;
    st_release_sys(reinterpret_cast<uint32_t*>(a.base[p]) + a.rank * COMM_MAX_BLOCKS + b, call);
    const uint32_t* my_pad = reinterpret_cast<const uint32_t*>(a.base[a.rank]) + p * COMM_MAX_BLOCKS + b;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(my_pad) - call) < 0) {
      if (clock64() - t0 > a.spin_limit) { failed = 1; break; }
    }
  }
  __syncthreads();
  const bool bad = failed != 0;
  const T poison = (T)__longlong_as_double(0x7ff8000000000000ll);
  const unsigned char* src[PDE_MAX_PEERS];
#pragma unroll
  for (int p = 0; p < PDE_MAX_PEERS; ++p)
    src[p] = (p < a.world) ? a.base[p] + SIGNAL_BYTES + (size_t)par * a.slot_elems * sizeof(T) : nullptr;
  for (long long i0 = v0; i0 < v1; i0 += (long long)U * blockDim.x) {
    uint4 v[PDE_MAX_PEERS][U];
#pragma unroll
    for (int p = 0; p < PDE_MAX_PEERS; ++p) {
      if (p < a.world) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = i0 + (long long)u * blockDim.x + threadIdx.x;
          if (i < v1) v[p][u] = (p == a.rank) ? reinterpret_cast<const uint4*>(mine)[i] : ld_relaxed_sys_v4(src[p] + i * 16);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + (long long)u * blockDim.x + threadIdx.x;
      if (i < v1) {
        uint4 acc = v[0][u];                     // rank order 0, 1, ..., W-1 on every rank
#pragma unroll
        for (int p = 1; p < PDE_MAX_PEERS; ++p)
          if (p < a.world) Vec16<T>::add(acc, v[p][u]);
        if (bad) {
          T* o = buf + i * VN;
          for (int e = 0; e < VN; ++e) o[e] = poison;
        } else {
          reinterpret_cast<uint4*>(buf)[i] = acc;
        }
      }
    }
  }
  if (tail) {
    for (long long i = nv * VN + threadIdx.x; i < a.n; i += blockDim.x) {
      T v = T(0);
      for (int p = 0; p < a.world; ++p) v += (p == a.rank) ? mine[i] : ld_relaxed_sys<T>(reinterpret_cast<const T*>(src[p]) + i);
      buf[i] = bad ? poison : v;
    }
  }
  // the last block to finish advances the call counter: by then every block has read it
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* done = reinterpret_cast<uint32_t*>(a.base[a.rank]) + PDE_MAX_PEERS * COMM_MAX_BLOCKS;
    __threadfence();
    if (atomicAdd(done, 1u) == (uint32_t)(G - 1)) {
      *done = 0u;
      __threadfence();
      *a.seq = call;
    }
  }
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
  if (n < 1 || n > slot_elems || (slot_elems & 3) != 0) return PDE_ERR_INVALID;   // slots stay 16-byte aligned
  CommArgs a;
  a.rank = peers->rank; a.world = peers->world;
  for (int p = 0; p < PDE_MAX_PEERS; ++p) {
    a.base[p] = p < peers->world ? static_cast<unsigned char*>(peers->base[p]) : nullptr;
    if (p < peers->world && !a.base[p]) return PDE_ERR_INVALID;
  }
  a.n = n; a.slot_elems = slot_elems; a.buf = buf; a.seq = static_cast<uint32_t*>(seq);
  a.spin_limit = 4000000000ll;   // ~2 s at 1.9 GHz
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // every rank derives the same block count from n (block b of every rank pairs with block b of its peers)
  const long long groups = n / (dtype == PDE_F32 ? 4 : 2);
  long long g = (groups + COMM_U * COMM_THREADS - 1) / (COMM_U * COMM_THREADS);
  const int grid = (int)(g < 1 ? 1 : (g > COMM_MAX_BLOCKS ? COMM_MAX_BLOCKS : g));
  if (dtype == PDE_F32) allreduce_oneshot_kernel<float><<<grid, COMM_THREADS, 0, st>>>(a);
  else allreduce_oneshot_kernel<double><<<grid, COMM_THREADS, 0, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? PDE_OK : PDE_ERR_CUDA;
}

}  // extern "C"
