// K12 on the tensor cores: gram[G, G] += X X^T for the fp32 [G, P] matrix of per-group gradients
// (ANCE/model/dro_loss.py:235-237 normalises the rows and multiplies G G^T; the norms are the diagonal).
//
// The matrix is 4.25 GB at G = 50 / BERT-base and is read ONCE; the arithmetic (2 G^2 P = 106 GFLOP) is what made the
// fp32 CUDA-core kernel (heads.cu gram_kernel) take 6.7 ms -- 10 % of the HBM rate.  Here the products run as
// tcgen05.mma kind::tf32 (fp32 operands straight from shared memory, read as TF32, fp32 accumulation in TMEM), so the
// kernel is bound by the stream from HBM:
//   * persistent CTAs, CTA b owns column chunks b, b + grid, ... of GRT_CHUNK = 128 columns
//   * warp 0 lane 0: TMA producer -- per chunk 4 boxes {32 fp32 columns, 64 rows} (rows >= G are zero-filled, nothing
//     is fetched for them), 128B-swizzled, into a GRT_STAGES-deep ring
//   * warp 1 lane 0: MMA issuer -- the SAME shared-memory tile is both operands (A = B = X[:, chunk], K-major):
//     D[128, 64] += A[128, 8] B[64, 8]^T per K step.  UMMA_M is 128 (lane == row layout) although only 64 rows are
//     staged: accumulator rows 64..127 are computed from whatever follows the tile in shared memory and never read
//     (each accumulator row depends on its own A row only); 8 KB of slack behind the ring keeps those reads in bounds
//   * the CTA's whole slab accumulates into ONE TMEM tile; at the end warps 0 and 1 (TMEM lanes 0..63) read it and add
//     their G x G block into the global result with atomics
// TF32 keeps 10 mantissa bits of every operand: the Gram entries agree with fp32 to ~1e-3 relative, well inside what
// the cosine-similarity statistics of iDRO need (h_fun is checked against the fp32 oracle at 1e-2).
#include "cdr_common.cuh"
#include "tma_host.h"

namespace cdr {

constexpr int GRT_ROWS = 64;     // staged rows (max groups)
constexpr int GRT_BOX = 32;      // fp32 columns per TMA box = one 128-byte swizzle row
constexpr int GRT_CHUNK = 128;   // columns per pipeline stage
constexpr int GRT_STAGES = 6;
constexpr int GRT_BOX_BYTES = GRT_ROWS * GRT_BOX * 4;                 // 8 KB
constexpr int GRT_STAGE_BYTES = (GRT_CHUNK / GRT_BOX) * GRT_BOX_BYTES;  // 32 KB
constexpr int GRT_SMEM = GRT_STAGES * GRT_STAGE_BYTES + GRT_BOX_BYTES + 256 + 1024;  // ring | slack | barriers | align

// kind::tf32: a/b format 2 (TF32), fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(64, 1)
gram_tf32_kernel(const __grid_constant__ CUtensorMap tma_x, int G, long long n_chunks, float* __restrict__ gram) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + GRT_STAGES * GRT_STAGE_BYTES + GRT_BOX_BYTES);
  uint64_t* empty = full + GRT_STAGES;
  uint64_t* done = empty + GRT_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_x);
    for (int i = 0; i < GRT_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const bool has_work = static_cast<long long>(blockIdx.x) < n_chunks;
  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_expect_tx(&full[stage], GRT_STAGE_BYTES);
#pragma unroll
      for (int b = 0; b < GRT_CHUNK / GRT_BOX; ++b)
        tma_load_2d(smem + stage * GRT_STAGE_BYTES + b * GRT_BOX_BYTES, &tma_x, &full[stage],
                    static_cast<int>(c * GRT_CHUNK + b * GRT_BOX), 0);
      if (++stage == GRT_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_tf32(128, 64);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t base = smem_u32(smem + stage * GRT_STAGE_BYTES);
#pragma unroll
      for (int b = 0; b < GRT_CHUNK / GRT_BOX; ++b)
#pragma unroll
        for (int k = 0; k < GRT_BOX / 8; ++k) {  // UMMA_K = 8 for TF32: +32 bytes inside the swizzle row
          const uint64_t d = make_smem_desc(base + b * GRT_BOX_BYTES + k * 32, 16, 1024);
          tc_mma_tf32(tmem, d, d, idesc, first ? 0u : 1u);
          first = false;
        }
      tc_commit(&empty[stage]);
      if (++stage == GRT_STAGES) { stage = 0; phase ^= 1; }
    }
    tc_commit(done);
  }
  __syncwarp();
  if (has_work) {
    // warps 0 / 1 own TMEM lanes 0..31 / 32..63 = accumulator rows (groups)
    mbar_wait(done, 0);
    tc_fence_after();
    const int g = warp * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(trow + c * 32, v);
      tc_wait_ld();
      if (g < G) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < G) atomicAdd(gram + g * G + c * 32 + j, __uint_as_float(v[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

// Returns CDR_OK when the tensor-core kernel was launched, 1 when the operand layout does not allow it (caller falls
// back to the CUDA-core kernel), a negative error code otherwise.
int gram_tf32_launch(const float* x, int g, long long p, long long ldx, float* gram, cudaStream_t st) {
  if (g > GRT_ROWS || (reinterpret_cast<uintptr_t>(x) & 15) != 0 || (ldx % 4) != 0 || p < GRT_CHUNK) return 1;
  CUtensorMap tm;
  if (int rc = make_tma_2d_f32(&tm, x, static_cast<uint64_t>(p), static_cast<uint64_t>(g), static_cast<uint64_t>(ldx),
                               GRT_BOX, GRT_ROWS))
    return rc;
  static bool cfg = false;
  if (!cfg) {
    CDR_CUDA(cudaFuncSetAttribute(gram_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRT_SMEM));
    cfg = true;
  }
  const long long n_chunks = (p + GRT_CHUNK - 1) / GRT_CHUNK;
  const int grid = static_cast<int>(n_chunks < sm_count() ? n_chunks : sm_count());
  gram_tf32_kernel<<<grid, 64, GRT_SMEM, st>>>(tm, g, n_chunks, gram);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

}  // namespace cdr
