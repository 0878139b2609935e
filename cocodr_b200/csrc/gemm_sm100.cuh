// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = A[M,K] * B[N,K]^T      (fp16 operands, fp32 accumulation in TMEM)
//
// * one CTA per SM, static round-robin over (m_tile, n_tile, k_split) work items
// * warp 0  : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count)
// * warp 1  : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, cta_group::1)
// * warps 2-9: epilogue (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// * two TMEM accumulator stages (2 x BN columns) so the epilogue of tile i overlaps the MMAs of
//   tile i+1
// * operands may be K-major ([rows, K] row-major, the "NT" case) or MN-major ([K, rows] row-major)
//   -- the latter is what dgrad (B = W[N_out, K_in]) and wgrad (A = dY^T, B = X^T) need, so no
//   transposed copies of activations or weights are ever made.
//
// Shared-memory tile layouts (both are the canonical SWIZZLE_128B UMMA layouts, see
// cute/atom/mma_traits_sm100.hpp "make_umma_desc"):
//   K-major : [rows][64 halfs]  128 B per row, 8-row swizzle atoms of 1024 B  (SBO = 1024)
//             one 2-D TMA box {64, rows}.   K-step of 16 halfs = +32 B on the start address.
//   MN-major: [rows/64][64 k][64 halfs]: per 64-wide chunk of rows a [k][64] slab of 128 B rows
//             (SBO = 1024 between 8-k groups, LBO = 8192 between 64-row chunks)
//             one 2-D TMA box {64 rows, 64 k} per 64-row chunk over the global view {rows, K}.
//             K-step of 16 = +2048 B on the start address.
#pragma once
#include "cdr_common.cuh"

namespace cdr {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;  // 320

struct GemmParams {
  int M, N, K;          // problem size (K = reduction length)
  int split_k;          // >= 1
  int kb_per_split;     // k-blocks (of 64) per split
  int m_tiles, n_tiles;
  // epilogue
  void* out;            // fp16 or fp32 [M, ldo]
  void* out2;           // optional second output (pre-activation), fp16
  const float* bias;    // [N] fp32 or null
  const __half* aux;    // residual R or pre-activation Z, [M, ldaux] fp16
  long long ldo, ldaux;
  float alpha;
  // scan filter epilogue
  const float* thresh;          // [N] per-query admission threshold
  unsigned long long* cand;     // [N, cand_cap] packed (score, doc) keys
  int* cand_count;              // [N]
  int cand_cap;
  long long row_base;           // global doc index of row 0
  // debug overrides for descriptor probing (0 = default)
  int dbg_lbo, dbg_sbo;
};

// host-side entry shared by cdr_gemm and the scan (gemm.cu)
int gemm_run(const cdr_gemm_args& g, GemmParams p, cudaStream_t st);

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
};

// Order-preserving packing of (score desc, doc asc) into one u64 so that a plain descending sort of
// the keys yields the scan contract order.
__device__ __forceinline__ unsigned long long pack_score_doc(float s, unsigned int doc) {
  unsigned int u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned long long>(~doc);
}

// Epilogue for one thread: 32 consecutive columns [n0, n0+32) of row m.
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, const uint32_t (&acc)[32], int m, int n0) {
  if (m >= p.M) return;
  if constexpr (EPI == CDR_EPI_SCAN_FILTER) {
    // rows = documents, columns = queries.  Admit (score >= thresh[q]) into the per-query buffers.
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      const int q = n0 + j;
      if (q >= p.N) break;
      const float s = __uint_as_float(acc[j]);
      if (s >= __ldg(p.thresh + q)) {
        const int pos = atomicAdd(p.cand_count + q, 1);
        if (pos < p.cand_cap)
          p.cand[static_cast<long long>(q) * p.cand_cap + pos] =
              pack_score_doc(s, static_cast<unsigned int>(p.row_base + m));
      }
    }
    return;
  } else if constexpr (EPI == CDR_EPI_F32_ATOMIC || EPI == CDR_EPI_F32_STORE) {
    float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(m) * p.ldo + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (n0 + j >= p.N) break;
      float4 v = make_float4(__uint_as_float(acc[j]) * p.alpha, __uint_as_float(acc[j + 1]) * p.alpha,
                             __uint_as_float(acc[j + 2]) * p.alpha, __uint_as_float(acc[j + 3]) * p.alpha);
      if constexpr (EPI == CDR_EPI_F32_ATOMIC) {
        atomicAdd(reinterpret_cast<float4*>(o + j), v);
      } else {
        *reinterpret_cast<float4*>(o + j) = v;
      }
    }
    return;
  } else {
    __half* o = reinterpret_cast<__half*>(p.out) + static_cast<long long>(m) * p.ldo + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (n0 + j >= p.N) break;
      float v[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(acc[j + t]) * p.alpha;
      if (p.bias != nullptr) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if constexpr (EPI == CDR_EPI_BIAS_RESIDUAL) {
        const uint4 r = *reinterpret_cast<const uint4*>(p.aux + static_cast<long long>(m) * p.ldaux + n0 + j);
        const __half2* rh = reinterpret_cast<const __half2*>(&r);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __half22float2(rh[t]);
          v[2 * t] += f.x;
          v[2 * t + 1] += f.y;
        }
      }
      if constexpr (EPI == CDR_EPI_DGELU) {
        const uint4 z = *reinterpret_cast<const uint4*>(p.aux + static_cast<long long>(m) * p.ldaux + n0 + j);
        const __half2* zh = reinterpret_cast<const __half2*>(&z);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __half22float2(zh[t]);
          v[2 * t] *= gelu_erf_grad(f.x);
          v[2 * t + 1] *= gelu_erf_grad(f.y);
        }
      }
      if constexpr (EPI == CDR_EPI_BIAS_GELU) {
        if (p.out2 != nullptr) {
          uint4 zq;
          __half2* zh = reinterpret_cast<__half2*>(&zq);
#pragma unroll
          for (int t = 0; t < 4; ++t) zh[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
          *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out2) + static_cast<long long>(m) * p.ldo + n0 + j) = zq;
          // GELU is applied to the fp16-rounded pre-activation so that backward (which only has Z)
          // differentiates exactly the function forward evaluated.
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = __half22float2(zh[t]);
            v[2 * t] = f.x;
            v[2 * t + 1] = f.y;
          }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = gelu_erf(v[t]);
      }
      uint4 q;
      __half2* qh = reinterpret_cast<__half2*>(&q);
#pragma unroll
      for (int t = 0; t < 4; ++t) qh[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
      *reinterpret_cast<uint4*>(o + j) = q;
    }
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const GemmParams p) {
  using S = GemmSmem<BN>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * S::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_items = p.m_tiles * p.n_tiles * p.split_k;
  const int total_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], GEMM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int tile = item / p.split_k;
        const int ks = item - tile * p.split_k;
        const int m0 = (tile / p.n_tiles) * GEMM_BM;
        const int n0 = (tile % p.n_tiles) * BN;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, total_kb);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          void* da = smem_a + stage * S::A_BYTES;
          void* db = smem_b + stage * S::B_BYTES;
          if constexpr (A_MN) {
#pragma unroll
            for (int c = 0; c < GEMM_BM / 64; ++c)
              tma_load_2d(static_cast<uint8_t*>(da) + c * (GEMM_BK * 128), &tma_a, &full_bar[stage], m0 + c * 64,
                          kb * GEMM_BK);
          } else {
            tma_load_2d(da, &tma_a, &full_bar[stage], kb * GEMM_BK, m0);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(static_cast<uint8_t*>(db) + c * (GEMM_BK * 128), &tma_b, &full_bar[stage], n0 + c * 64,
                          kb * GEMM_BK);
          } else {
            tma_load_2d(db, &tma_b, &full_bar[stage], kb * GEMM_BK, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // descriptor geometry
      const uint32_t a_lbo = A_MN ? (p.dbg_lbo ? p.dbg_lbo : GEMM_BK * 128) : 16;
      const uint32_t b_lbo = B_MN ? (p.dbg_lbo ? p.dbg_lbo : GEMM_BK * 128) : 16;
      const uint32_t sbo = p.dbg_sbo ? p.dbg_sbo : 1024;
      constexpr uint32_t a_kstep = A_MN ? 2048 : 32;   // bytes per UMMA_K = 16
      constexpr uint32_t b_kstep = B_MN ? 2048 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int ks = item % p.split_k;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, total_kb);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * S::A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * S::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t ad = make_smem_desc(sa + k * a_kstep, a_lbo, sbo);
            const uint64_t bd = make_smem_desc(sb + k * b_kstep, b_lbo, sbo);
            tc_mma_f16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;             // 0..7
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int half = ew >> 2;            // which half of the column chunks this warp handles
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int tile = item / p.split_k;
      const int m0 = (tile / p.n_tiles) * GEMM_BM;
      const int n0 = (tile % p.n_tiles) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m = m0 + quad * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
      constexpr int CHUNKS = BN / 32;
#pragma unroll 1
      for (int c = half; c < CHUNKS; c += 2) {
        if (n0 + c * 32 >= p.N) break;
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tc_wait_ld();
        gemm_epilogue_chunk<EPI>(p, r, m, n0 + c * 32);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

}  // namespace cdr
