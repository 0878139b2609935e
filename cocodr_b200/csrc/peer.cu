// Peer-memory exchange of the contrastive head (the ONE real exchange step of the training path is
// the all-gather of the passage CLS embeddings and, in backward, the reduce-scatter of their gradients).
// Instead of NCCL collectives, kernels write straight into the other ranks' HBM over NVLink (symmetric
// allocations: the same buffer exists on every rank and every rank holds all base pointers):
//
//   forward : the last LayerNorm kernel (elementwise.cu, cdr_ln_fwd_push) stores each passage CLS row into slot
//             `rank` of EVERY rank's gather buffer as it produces it, and raises flag[rank] on every peer when its
//             grid has finished;  cdr_peer_wait holds the consumer until all `world` flags carry the step's epoch.
//   backward: cdr_peer_scatter_rows sends row block r of d(P_all) to slot `rank` of rank r's receive buffer
//             (+ flags);  cdr_peer_reduce_slots waits for the world's contributions and sums the slots.
//
// Flags are monotonically increasing epochs (no reset, no ABA); data -> __threadfence_system -> st.release.sys
// on the writer, ld.acquire.sys spin on the reader.
#include "peer.cuh"

namespace cdr {

__global__ void peer_wait_kernel(const uint32_t* flags, int world, const uint32_t* epoch) {
  if (threadIdx.x == 0) peer_wait_all(flags, world, *reinterpret_cast<const volatile uint32_t*>(epoch));
}

__global__ void peer_next_epoch_kernel(uint32_t* epoch) { epoch[0] += 1u; }

// Wait for the pushes of this epoch, then copy the epoch's half of the double-buffered gather area into a private
// tensor: the consumer (and its saved-for-backward state) never aliases memory the peers write into, and a peer
// that is already one step ahead writes the OTHER half.
__global__ void __launch_bounds__(256)
peer_wait_fetch_kernel(const uint32_t* __restrict__ flags, int world, const uint32_t* __restrict__ epoch,
                       const float* __restrict__ gather, long long half_elems, float* __restrict__ out) {
  const uint32_t ep = *reinterpret_cast<const volatile uint32_t*>(epoch);
  if (threadIdx.x == 0) peer_wait_all(flags, world, ep);
  __syncthreads();
  const float* src = gather + (ep & 1u) * half_elems;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < half_elems; i += stride)
    *reinterpret_cast<float4*>(out + i) = *reinterpret_cast<const float4*>(src + i);
}

// src [world * rows, dim] fp32: row block r -> slot `rank` of rank r's receive buffer [world][rows, dim]
__global__ void __launch_bounds__(256)
peer_scatter_kernel(const float* __restrict__ src, long long per_block, cdr_peer_args pa) {
  const long long total = per_block * pa.world;  // floats, multiple of 4
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < total; i += stride) {
    const int r = static_cast<int>(i / per_block);
    const long long off = i - static_cast<long long>(r) * per_block;
    float* dst = static_cast<float*>(pa.peer_buf[r]) + static_cast<long long>(pa.rank) * per_block + off;
    *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src + i);
  }
  peer_signal_grid_done(pa, 1);
}

// out[i] = sum_r recv[r][i] once every rank's contribution for `epoch` has landed
__global__ void __launch_bounds__(256)
peer_reduce_kernel(const float* __restrict__ recv, const uint32_t* __restrict__ flags, int world, long long n,
                   const uint32_t* epoch, float* __restrict__ out) {
  if (threadIdx.x == 0) peer_wait_all(flags, world, *reinterpret_cast<const volatile uint32_t*>(epoch));
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 v = *reinterpret_cast<const float4*>(recv + static_cast<long long>(r) * n + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + i) = acc;
  }
}

int peer_check(const cdr_peer_args* pa, const char* who) {
  CDR_REQUIRE(pa != nullptr, "%s: null peer args", who);
  CDR_REQUIRE(pa->world >= 1 && pa->world <= 8 && pa->rank >= 0 && pa->rank < pa->world, "%s: bad world / rank", who);
  CDR_REQUIRE(pa->done_counter != nullptr && pa->epoch != nullptr, "%s: done_counter and epoch are required", who);
  for (int r = 0; r < pa->world; ++r)
    CDR_REQUIRE(pa->peer_buf[r] != nullptr && pa->peer_flag[r] != nullptr, "%s: null peer pointer %d", who, r);
  return CDR_OK;
}

}  // namespace cdr

using namespace cdr;

extern "C" {

int cdr_peer_next_epoch(uint32_t* epoch, void* stream) {
  CDR_REQUIRE(epoch != nullptr, "cdr_peer_next_epoch: null pointer");
  peer_next_epoch_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(epoch);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_peer_wait(const uint32_t* local_flags, int32_t world, const uint32_t* epoch, void* stream) {
  CDR_REQUIRE(local_flags != nullptr && epoch != nullptr && world >= 1 && world <= 8, "cdr_peer_wait: bad arguments");
  peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(local_flags, world, epoch);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_peer_wait_fetch(const uint32_t* local_flags, int32_t world, const uint32_t* epoch, const float* gather,
                        int64_t half_elems, float* out, void* stream) {
  CDR_REQUIRE(local_flags && epoch && gather && out && world >= 1 && world <= 8 && half_elems > 0 && half_elems % 4 == 0,
              "cdr_peer_wait_fetch: bad arguments (half_elems must be a multiple of 4)");
  long long blocks = (half_elems / 4 + 255) / 256;
  if (blocks > 2 * sm_count()) blocks = 2 * sm_count();
  peer_wait_fetch_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      local_flags, world, epoch, gather, half_elems, out);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_peer_scatter_rows(const float* src, int32_t rows, int32_t dim, const cdr_peer_args* pa, void* stream) {
  if (int rc = peer_check(pa, "cdr_peer_scatter_rows")) return rc;
  CDR_REQUIRE(src != nullptr && rows > 0 && dim > 0 && (static_cast<long long>(rows) * dim) % 4 == 0,
              "cdr_peer_scatter_rows: rows * dim must be a positive multiple of 4");
  const long long per_block = static_cast<long long>(rows) * dim;
  long long blocks = (per_block * pa->world / 4 + 255) / 256;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  peer_scatter_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, per_block, *pa);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_peer_reduce_slots(const float* recv, const uint32_t* local_flags, int32_t world, int64_t n,
                          const uint32_t* epoch, float* out, void* stream) {
  CDR_REQUIRE(recv && local_flags && epoch && out && world >= 1 && world <= 8 && n > 0 && n % 4 == 0,
              "cdr_peer_reduce_slots: bad arguments (n must be a multiple of 4)");
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 2 * sm_count()) blocks = 2 * sm_count();
  peer_reduce_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(recv, local_flags, world,
                                                                                                n, epoch, out);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

}  // extern "C"
