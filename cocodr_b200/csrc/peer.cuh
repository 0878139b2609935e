// Device-side helpers of the peer-memory exchange (peer.cu): flag publication / polling over NVLink.
#pragma once
#include "cdr_common.cuh"

namespace cdr {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Last block of the grid publishes `epoch` in flag[rank] of every peer.  Call with all peer stores of the calling
// thread done; contains __syncthreads.
__device__ __forceinline__ void peer_signal_grid_done(const cdr_peer_args& pa, int flag_set) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    const unsigned int prev = atomicAdd(pa.done_counter, 1u);
    if (prev == total - 1) {
      __threadfence_system();
      const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(pa.epoch);
      for (int r = 0; r < pa.world; ++r) st_release_sys(pa.peer_flag[r] + flag_set * 8 + pa.rank, epoch);
      *pa.done_counter = 0u;
    }
  }
}

__device__ __forceinline__ void peer_wait_all(const uint32_t* flags, int world, uint32_t epoch) {
  // epochs only grow: (int)(flag - epoch) >= 0 also survives wrap-around
  for (int r = 0; r < world; ++r)
    while (static_cast<int>(ld_acquire_sys(flags + r) - epoch) < 0) {
    }
}

int peer_check(const cdr_peer_args* pa, const char* who);  // peer.cu

}  // namespace cdr
