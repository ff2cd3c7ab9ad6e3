// pde_comm.cuh — device side of the flag-in-data push all-reduce over NVLink peer memory (see pde_comm.cu for the
// protocol).  Shared by the standalone kernel (pde_allreduce_oneshot) and by reduce_kernel, whose tail pushes every
// reduced gradient element straight to the peers (pde_residual_loss_grad_exchange: one launch fewer per step).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pde_b200.h"

namespace pde {
namespace comm {

constexpr int CTRL_BYTES = 512;     // local control words in front of the slots: [0] block-completion counter, [16] timeouts
constexpr int CTRL_ERR_WORD = 16;   // uint32 index of the error counter inside the control block

// (value, flag) travels as ONE 64-bit scalar access: single-copy atomic in the PTX memory model
__device__ __forceinline__ void st_pair_sys(void* p, uint32_t v, uint32_t flag) {
  const unsigned long long w = ((unsigned long long)flag << 32) | v;
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ uint2 ld_pair_sys(const void* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return make_uint2((uint32_t)w, (uint32_t)(w >> 32));
}

struct CommArgs {
  int rank, world;
  unsigned char* base[PDE_MAX_PEERS];   // peer-visible allocation of every rank: [ctrl | parity 0: [W][slot] pairs | parity 1]
  long long n_words, slot_words;        // 32-bit words of the vector / capacity of one slot
  int is_f64;
  void* buf;                            // local vector, reduced in place (standalone kernel only)
  uint32_t* seq;                        // local: number of completed calls
  long long spin_limit;                 // clock64 ticks before giving up (<= 0: never): result is poisoned with NaN
};

__device__ __forceinline__ unsigned char* slot_of(unsigned char* base, int par, int src, long long slot_words) {
  return base + CTRL_BYTES + ((size_t)par * PDE_MAX_PEERS + src) * (size_t)slot_words * 8;
}

// Element i of this rank's vector holds `mine`: push it to every peer, collect the peers' values and return the
// rank-ordered sum 0, 1, ..., W-1 (identical bits on every rank).  `call` = *seq + 1 read at kernel start.
template <typename T>
__device__ __forceinline__ T exchange_element(const CommArgs& a, long long i, T mine, uint32_t call) {
  constexpr int WPE = sizeof(T) / 4;             // 32-bit words per element
  const int par = (int)(call & 1u);
  uint32_t w[WPE];
  if (WPE == 1) {
    w[0] = __float_as_uint((float)mine);
  } else {
    const double d = (double)mine;
    w[0] = (uint32_t)__double2loint(d);
    w[WPE - 1] = (uint32_t)__double2hiint(d);
  }
#pragma unroll
  for (int p = 0; p < PDE_MAX_PEERS; ++p) {
    if (p < a.world && p != a.rank) {
      unsigned char* dst = slot_of(a.base[p], par, a.rank, a.slot_words) + (size_t)i * WPE * 8;
#pragma unroll
      for (int k = 0; k < WPE; ++k) st_pair_sys(dst + 8 * k, w[k], call);
    }
  }
  unsigned char* my_slots = slot_of(a.base[a.rank], par, 0, a.slot_words);
  uint32_t got[PDE_MAX_PEERS][WPE];
  uint32_t pending = 0;
#pragma unroll
  for (int p = 0; p < PDE_MAX_PEERS; ++p)
    if (p < a.world && p != a.rank) pending |= 1u << p;
  const long long t0 = clock64();
  bool bad = false;
  while (pending) {
#pragma unroll
    for (int p = 0; p < PDE_MAX_PEERS; ++p) {
      if (pending & (1u << p)) {
        const unsigned char* src = my_slots + (size_t)p * a.slot_words * 8 + (size_t)i * WPE * 8;
        uint2 v[WPE];
#pragma unroll
        for (int k = 0; k < WPE; ++k) v[k] = ld_pair_sys(src + 8 * k);
        bool ok = true;
#pragma unroll
        for (int k = 0; k < WPE; ++k) ok = ok && (v[k].y == call);
        if (ok) {
#pragma unroll
          for (int k = 0; k < WPE; ++k) got[p][k] = v[k].x;
          pending &= ~(1u << p);
        }
      }
    }
    if (pending && a.spin_limit > 0 && clock64() - t0 > a.spin_limit) { bad = true; break; }
  }
  T acc = T(0);
#pragma unroll
  for (int p = 0; p < PDE_MAX_PEERS; ++p) {
    if (p < a.world) {
      T v;
      if (p == a.rank) {
        v = mine;
      } else {
        if (WPE == 1) v = (T)__uint_as_float(got[p][0]); else v = (T)__hiloint2double(got[p][WPE - 1], got[p][0]);
      }
      acc = (p == 0) ? v : acc + v;
    }
  }
  if (bad) {
    acc = (T)__longlong_as_double(0x7ff8000000000000ll);
    atomicAdd(reinterpret_cast<uint32_t*>(a.base[a.rank]) + CTRL_ERR_WORD, 1u);   // seen by pde_exchange_errors()
  }
  return acc;
}

// End of a kernel that exchanged: the last block to finish advances the call counter (by then every block has read it).
// Call from all threads of the block after the block's last exchange_element.
__device__ __forceinline__ void finish_call(const CommArgs& a, uint32_t call) {
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* done = reinterpret_cast<uint32_t*>(a.base[a.rank]);
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      *done = 0u;
      __threadfence();
      *a.seq = call;
    }
  }
}

}  // namespace comm

// host side (pde_comm.cu): fill the device-side arguments from the ABI structs; 0 on success
int comm_fill_args(const pde_peers* peers, int32_t dtype, int64_t n, int64_t slot_elems, void* seq, comm::CommArgs* out);

}  // namespace pde
