// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = A[M,K] * B[N,K]^T      (fp16 operands, fp32 accumulation in TMEM)
//
// * one CTA per SM, static round-robin over (m_tile, n_tile, k_split) work items
// * warp 0  : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count)
// * warp 1  : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16; CG = 2 pairs two CTAs of a
//             cluster on one 256-row tile with cta_group::2)
// * warps 2..: epilogue, GEMM_EW<EPI> warps (8; 16 for the scan filters): tcgen05.ld 32x32b -> registers -> TMEM
//             stage released -> fused epilogue -> global
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
#include <type_traits>

#include "cdr_common.cuh"
#include "dropout.cuh"

namespace cdr {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
// Epilogue warps (a multiple of 4: one group per TMEM lane quadrant).  With the accumulators released early and
// the aux operand prefetched, 8 warps beat 16 for the dense epilogues (32 KB less staging smem = one more TMA
// stage); the scan filter is bound by the round trip of its slot-reserving atomics and wants 16.
#ifndef CDR_GEMM_EPI_WARPS
#define CDR_GEMM_EPI_WARPS 8
#endif
#ifndef CDR_GEMM_GELU_EPI_WARPS  // the GELU epilogue (value + derivative, ~24 instructions per element) is issue-bound
#define CDR_GEMM_GELU_EPI_WARPS CDR_GEMM_EPI_WARPS
#endif
template <int EPI>
constexpr int GEMM_EW = (EPI == CDR_EPI_SCAN_FILTER || EPI == CDR_EPI_SCAN_FILTER_Q) ? 16
                        : (EPI == CDR_EPI_BIAS_GELU)                                  ? CDR_GEMM_GELU_EPI_WARPS
                                                                                      : CDR_GEMM_EPI_WARPS;
template <int EPI>
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EW<EPI>;

#ifdef CDR_GEMM_DEBUG
#define CDR_DBG(p, bit) ((p).dbg_flags & (bit))
#define CDR_DBG_OR(v, dflt) ((v) ? (v) : (dflt))
#else
#define CDR_DBG(p, bit) 0
#define CDR_DBG_OR(v, dflt) (dflt)
#endif

struct GemmParams {
  int M, N, K;          // problem size (K = reduction length)
  int split_k;          // >= 1
  int kb_per_split;     // k-blocks (of 64) per split
  int m_tiles, n_tiles;
  int cta_group;        // 1 or 2 (CTA pair, 256-row tiles)
  // epilogue
  void* out;            // fp16 or fp32 [M, ldo]
  void* out2;           // optional second output (pre-activation), fp16
  const float* bias;    // [N] fp32 or null
  const __half* aux;    // residual R or pre-activation Z, [M, ldaux] fp16
  long long ldo, ldaux;
  float alpha;
  float* colsum;        // optional fp32 [N]: += colsum_scale * column sums of the (fp32, pre-rounding) output
  float colsum_scale;
  // scan filter epilogue
  const float* thresh;          // [N] per-query admission threshold
  unsigned long long* cand;     // [N, cand_cap] packed (score, doc) keys
  int* cand_count;              // [N]
  int cand_cap;
  long long row_base;           // global doc index of row 0
  // descriptor probing / timing experiments: honoured by -DCDR_GEMM_DEBUG builds only (tools/build_variant.sh); the
  // release kernels ignore them, so no environment variable or argument can make a product kernel skip work
  int dbg_lbo, dbg_sbo;
  long long a_rows_alloc;  // same for A
  long long b_rows_alloc;  // rows of B that physically exist (>= N; 0 = N): lets TMA boxes read zero padding instead of going out of bounds
  int dbg_flags;  // CDR_GEMM_DEBUG builds: CDR_GEMM_DBG (environment) bit 0 = skip the epilogue math and stores
  cdr_dropout drop;  // CDR_EPI_BIAS_DROP_RESIDUAL
  // CDR_EPI_F32_GROUPED: split s reduces k-blocks [seg_kb[s], seg_kb[s+1]) (device array of split_k + 1 ascending
  // entries) into out + s * seg_out_stride -- the iDRO per-group wgrad as one launch
  const int* seg_kb;
  long long seg_out_stride;
};

// k-blocks of split `ks`: uniform ranges, or the per-group table of the grouped wgrad
template <int EPI>
__device__ __forceinline__ void gemm_split_range(const GemmParams& p, int ks, int total_kb, int& kb0, int& kb1) {
  if constexpr (EPI == CDR_EPI_F32_GROUPED) {
    kb0 = __ldg(p.seg_kb + ks);
    kb1 = min(__ldg(p.seg_kb + ks + 1), total_kb);
  } else {
    kb0 = ks * p.kb_per_split;
    kb1 = min(kb0 + p.kb_per_split, total_kb);
  }
}

// host-side entry shared by cdr_gemm and the scan (gemm.cu)
int gemm_run(const cdr_gemm_args& g, GemmParams p, cudaStream_t st);

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cta_group::2) per 256 x BN tile -- each CTA stages
// its own 128 rows of A and HALF of the B tile (BN/2 rows); the pair's UMMA reads both halves, so the
// operand bytes pulled through L2 per MMA cycle drop by a third and the ring gets deeper.
template <int BN, int CG, int EW>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_ROWS = BN / CG;
  static constexpr int B_BYTES = B_ROWS * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STG_BYTES = EW * 32 * 32 * 4;  // epilogue staging tiles
  static constexpr int BUDGET = 232448 - 1024 - BAR_BYTES - STG_BYTES;  // 227 KB per CTA
  static constexpr int STAGES = (BUDGET / STAGE_BYTES) > 8 ? 8 : (BUDGET / STAGE_BYTES);
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + STG_BYTES + 1024;  // +1024 alignment slack
  static_assert(TOTAL <= 232448, "shared memory budget");
};

// Order-preserving packing of (score desc, doc asc) into one u64 so that a plain descending sort of
// the keys yields the scan contract order.
__device__ __forceinline__ unsigned long long pack_score_doc(float s, unsigned int doc) {
  unsigned int u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned long long>(~doc);
}

// ------------------------------------------------------------------------------------------------
// Epilogue.  tcgen05.ld hands every thread one accumulator ROW (TMEM lane) -- the wrong shape for global
// memory: 32 threads would touch 32 different rows per instruction.  Each epilogue warp therefore
// transposes its 32 x 32 fp32 chunk through a private, XOR-swizzled 4 KB staging tile in shared memory and
// re-reads it so that 4 neighbouring threads own 8 consecutive columns each of ONE row: residual /
// pre-activation loads and output stores become 64-128 B contiguous per row (full 32 B sectors), and the
// per-column bias / thresholds live in registers for the whole chunk.
// ------------------------------------------------------------------------------------------------
constexpr int GEMM_STG_BYTES = 32 * 32 * 4;  // per epilogue warp

// row r, 16-byte piece j (0..7) of a [32][32] fp32 tile, conflict-free for both access shapes
__device__ __forceinline__ float4* stg_piece(float* stg, int r, int j) {
  return reinterpret_cast<float4*>(stg + r * 32 + ((j ^ (r & 7)) << 2));
}

// 8 consecutive columns [n, n+8) of row m; v = raw accumulators
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_apply(const GemmParams& p, float (&v)[8], int m, int n,
                                                    const float (&bias)[8], const uint4& auxq, const DropCtx& dc,
                                                    uint32_t keep = 0xffu) {
  if constexpr (EPI == CDR_EPI_SCAN_FILTER) {
    // rows = documents, columns = queries; bias[] holds the admission thresholds of the 8 queries
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int q = n + j;
      if (q < p.N && v[j] >= bias[j]) {
        const int pos = atomicAdd(p.cand_count + q, 1);
        if (pos < p.cand_cap)
          p.cand[static_cast<long long>(q) * p.cand_cap + pos] =
              pack_score_doc(v[j], static_cast<unsigned int>(p.row_base + m));
      }
    }
  } else if constexpr (EPI == CDR_EPI_F32_ATOMIC || EPI == CDR_EPI_F32_STORE) {
    float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(m) * p.ldo + n;
    const float4 lo = make_float4(v[0] * p.alpha, v[1] * p.alpha, v[2] * p.alpha, v[3] * p.alpha);
    const float4 hi = make_float4(v[4] * p.alpha, v[5] * p.alpha, v[6] * p.alpha, v[7] * p.alpha);
    if constexpr (EPI == CDR_EPI_F32_ATOMIC) {
      if (CDR_DBG(p, 2)) return;
      atomicAdd(reinterpret_cast<float4*>(o), lo);
      atomicAdd(reinterpret_cast<float4*>(o + 4), hi);
    } else {
      if (CDR_DBG(p, 2)) return;
      *reinterpret_cast<float4*>(o) = lo;
      *reinterpret_cast<float4*>(o + 4) = hi;
    }
  } else {
    // packed fp32 pairs (FFMA2): the epilogue warps are bound by fp32-pipe issue slots
    f32x2 w[4];
    {
      const f32x2 al = pk2(p.alpha);
#pragma unroll
      for (int t = 0; t < 4; ++t) w[t] = fma2(pk2(v[2 * t], v[2 * t + 1]), al, pk2(bias[2 * t], bias[2 * t + 1]));
    }
    if constexpr (EPI == CDR_EPI_BIAS_DROP_RESIDUAL) {  // HF BertSelfOutput / BertOutput: LN(x + dropout(dense(h)))
#pragma unroll
      for (int t = 0; t < 4; ++t)
        w[t] = mul2(w[t], pk2(((keep >> (2 * t)) & 1u) ? dc.scale : 0.f, ((keep >> (2 * t + 1)) & 1u) ? dc.scale : 0.f));
    }
    if constexpr (EPI == CDR_EPI_BIAS_RESIDUAL || EPI == CDR_EPI_BIAS_DROP_RESIDUAL) {
      const __half2* rh = reinterpret_cast<const __half2*>(&auxq);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(rh[t]);
        w[t] = add2(w[t], pk2(f.x, f.y));
      }
    }
    if constexpr (EPI == CDR_EPI_DGELU) {
      const __half2* gh = reinterpret_cast<const __half2*>(&auxq);  // gelu'(z) saved by the forward epilogue
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(gh[t]);
        w[t] = mul2(w[t], pk2(f.x, f.y));
      }
    }
    if constexpr (EPI == CDR_EPI_BIAS_GELU) {
      f32x2 d[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float z0, z1;
        upk2(w[t], z0, z1);
        gelu_erf_both2(z0, z1, w[t], d[t]);
      }
      if (p.out2 != nullptr && !CDR_DBG(p, 2)) {
        uint4 dq;
        __half2* dh = reinterpret_cast<__half2*>(&dq);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float d0, d1;
          upk2(d[t], d0, d1);
          dh[t] = __floats2half2_rn(d0, d1);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out2) + static_cast<long long>(m) * p.ldo + n) = dq;
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) upk2(w[t], v[2 * t], v[2 * t + 1]);
    uint4 q;
    __half2* qh = reinterpret_cast<__half2*>(&q);
#pragma unroll
    for (int t = 0; t < 4; ++t) qh[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
    if (CDR_DBG(p, 2)) {
      if (q.x == 0x12345678u && q.y == 0x9abcdef0u) reinterpret_cast<uint4*>(p.out)[0] = q;  // keep the math alive
      return;
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + static_cast<long long>(m) * p.ldo + n) = q;
  }
}

// Scan filter straight from the accumulator registers (thread = one document row, 32 query columns): nothing
// dense is stored, so the chunk needs no transposition.  Admission is two-phase so that a warp pays ONE atomic
// round trip per chunk instead of one per admitted score:
//   reserve: per column j a ballot of the rows that pass threshold[j]; lane j owns column j and reserves
//            popc(ballot) slots of that query's candidate buffer with a single atomicAdd;
//   commit : every passing (row, column) writes its packed key at base[j] + (rank of the row in the ballot).
// ws = 96 words of per-warp shared memory (thresholds | ballots | bases).
struct FilterTicket {
  uint32_t ballot;  // lane j: rows of this chunk admitted for column j
  int base;         // lane j: first reserved slot
};

__device__ __forceinline__ FilterTicket gemm_filter_reserve(const GemmParams& p, const uint32_t (&acc)[32],
                                                            uint32_t* ws, int lane, int m, int n0) {
  const int q = n0 + lane;
  ws[lane] = __float_as_uint(q < p.N ? __ldg(p.thresh + q) : INFINITY);
  __syncwarp();
  const bool row_ok = m < p.M;
  FilterTicket t;
  t.ballot = 0u;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const uint32_t b = __ballot_sync(0xffffffffu, row_ok && __uint_as_float(acc[j]) >= __uint_as_float(ws[j]));
    if (lane == j) t.ballot = b;
  }
  t.base = 0;
  if (t.ballot != 0u) t.base = atomicAdd(p.cand_count + q, __popc(t.ballot));
  __syncwarp();
  return t;
}

__device__ __forceinline__ void gemm_filter_commit(const GemmParams& p, const uint32_t (&acc)[32], uint32_t* ws,
                                                   int lane, int m, int n0, const FilterTicket& t) {
  ws[32 + lane] = t.ballot;
  ws[64 + lane] = static_cast<uint32_t>(t.base);
  __syncwarp();
  const unsigned int doc = static_cast<unsigned int>(p.row_base + m);
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const uint32_t b = ws[32 + j];  // warp-uniform
    if (b != 0u) {
      if ((b >> lane) & 1u) {
        const int pos = static_cast<int>(ws[64 + j]) + __popc(b & lt);
        if (pos < p.cand_cap)
          p.cand[static_cast<long long>(n0 + j) * p.cand_cap + pos] = pack_score_doc(__uint_as_float(acc[j]), doc);
      }
    }
  }
  __syncwarp();
}

// Scan filter with the QUERIES on the accumulator rows (thread = one query, 32 document columns per chunk): the
// threshold is a per-thread scalar and a thread reserves the slots of BOTH chunks it holds with one atomicAdd (one
// atomic round trip per tile and warp) -- no cross-lane traffic at all.
__device__ __forceinline__ uint32_t gemm_filter_rows_mask(const GemmParams& p, const uint32_t (&acc)[32], int n0, float th) {
  uint32_t mask = 0u;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (n0 + j < p.N && __uint_as_float(acc[j]) >= th) mask |= 1u << j;
  return mask;
}
__device__ __forceinline__ void gemm_filter_rows_commit(const GemmParams& p, const uint32_t (&acc)[32], int q, int n0,
                                                        uint32_t mask, int base) {
  unsigned long long* dst = p.cand + static_cast<long long>(q) * p.cand_cap;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if ((mask >> j) & 1u) {
      const int pos = base + __popc(mask & ((1u << j) - 1u));
      if (pos < p.cand_cap) dst[pos] = pack_score_doc(__uint_as_float(acc[j]), static_cast<unsigned int>(p.row_base + n0 + j));
    }
}

template <int EPI>
constexpr bool GEMM_EPI_HAS_AUX =
    (EPI == CDR_EPI_BIAS_RESIDUAL || EPI == CDR_EPI_DGELU || EPI == CDR_EPI_BIAS_DROP_RESIDUAL);

// The residual / saved-derivative operand of one 32 x 32 chunk, in the post-transpose thread mapping
// (thread -> 8 columns of rows (lane>>2) + 8i).  Issued BEFORE the accumulator is waited for, so the global
// latency overlaps the tile's MMAs.
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_load_aux(const GemmParams& p, uint4 (&auxq)[4], int lane, int m_base,
                                                       int n0) {
  if constexpr (GEMM_EPI_HAS_AUX<EPI>) {
    const int n = n0 + (lane & 3) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m_base + (lane >> 2) + 8 * i;
      auxq[i] = make_uint4(0u, 0u, 0u, 0u);
      if (n < p.N && m < p.M && !CDR_DBG(p, 4)) auxq[i] = *reinterpret_cast<const uint4*>(p.aux + static_cast<long long>(m) * p.ldaux + n);
    }
  }
}

// One warp, one 32 x 32 chunk: rows [m_base, m_base+32), columns [n0, n0+32).  acc = this thread's row.
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, const uint32_t (&acc)[32], float* stg,
                                                    int lane, int m_base, int n0, const uint4 (&auxq)[4],
                                                    const DropCtx& dc) {
  const bool no_stage = CDR_DBG(p, 16);  // smem-contention experiment: skip the transposition (wrong values, same math)
  if (!no_stage) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *stg_piece(stg, lane, j) = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                             __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
    __syncwarp();
  }
  const int seg = lane & 3;
  const int n = n0 + seg * 8;
  const bool col_ok = n < p.N;
  float bias[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) bias[t] = 0.f;
  if constexpr (EPI == CDR_EPI_SCAN_FILTER) {
#pragma unroll
    for (int t = 0; t < 8; ++t) bias[t] = (n + t < p.N) ? __ldg(p.thresh + n + t) : INFINITY;
  } else if constexpr (EPI != CDR_EPI_F32_ATOMIC && EPI != CDR_EPI_F32_STORE) {
    if (col_ok && p.bias != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
      bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
      bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
    }
  }
  float csum[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) csum[t] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 2) + 8 * i;
    const int m = m_base + r;
    float4 lo, hi;
    if (no_stage) {
      lo = make_float4(__uint_as_float(acc[8 * i]), __uint_as_float(acc[8 * i + 1]), __uint_as_float(acc[8 * i + 2]), __uint_as_float(acc[8 * i + 3]));
      hi = make_float4(__uint_as_float(acc[8 * i + 4]), __uint_as_float(acc[8 * i + 5]), __uint_as_float(acc[8 * i + 6]), __uint_as_float(acc[8 * i + 7]));
    } else {
      lo = *stg_piece(stg, r, 2 * seg);
      hi = *stg_piece(stg, r, 2 * seg + 1);
    }
    float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t keep = 0xffu;
    if constexpr (EPI == CDR_EPI_BIAS_DROP_RESIDUAL) {
      keep = (col_ok && m < p.M) ? drop_keep8(dc, drop_group(dc, m, n, p.N)) : 0u;
      if (dc.keep_bits != nullptr) {
        // the 4 lanes of a row hold the keep bytes of 32 consecutive columns: one 4-byte store per row (all lanes
        // shuffle; N % 32 == 0 is checked by the host, so a row's 4 bytes are all valid or all out of range)
        uint32_t b = keep;
        b |= __shfl_down_sync(0xffffffffu, b, 1) << 8;
        b |= __shfl_down_sync(0xffffffffu, b, 2) << 16;
        if (seg == 0 && col_ok && m < p.M)
          *reinterpret_cast<uint32_t*>(dc.keep_bits + static_cast<long long>(m) * (p.N >> 3) + (n >> 3)) = b;
      }
    }
    if (col_ok && m < p.M) {
      gemm_epilogue_apply<EPI>(p, v, m, n, bias, auxq[i], dc, keep);
      if constexpr (EPI == CDR_EPI_DGELU) {
#pragma unroll
        for (int t = 0; t < 8; ++t) csum[t] += v[t];
      }
    }
  }
  if constexpr (EPI == CDR_EPI_DGELU) {
    // fused bias gradient: column sums of this chunk.  The 8 lanes that share `seg` hold the same 8 columns
    // for different rows; a halving butterfly over lane bits 4,3,2 leaves every lane with ONE column total.
    if (p.colsum != nullptr) {
      float a4[4], a2[2];
      {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float keep = up ? csum[t + 4] : csum[t];
          const float send = up ? csum[t] : csum[t + 4];
          a4[t] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
      }
      {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const float keep = up ? a4[t + 2] : a4[t];
          const float send = up ? a4[t] : a4[t + 2];
          a2[t] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
      }
      const bool up = (lane & 4) != 0;
      const float keep = up ? a2[1] : a2[0];
      const float send = up ? a2[0] : a2[1];
      const float tot = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      const int col = n + ((lane & 16) ? 4 : 0) + ((lane & 8) ? 2 : 0) + ((lane & 4) ? 1 : 0);
      if (col < p.N) atomicAdd(p.colsum + col, tot * p.colsum_scale);
    }
  }
  __syncwarp();  // staging tile is rewritten by the next chunk
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CG>
__global__ void __launch_bounds__(GEMM_THREADS<EPI>, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const GemmParams p) {
  using S = GemmSmem<BN, CG, GEMM_EW<EPI>>;
  constexpr int GEMM_EPI_WARPS = GEMM_EW<EPI>;
  constexpr int STAGES = S::STAGES;
  constexpr int BNL = S::B_ROWS;  // B rows staged by this CTA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * S::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* stg_all = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + S::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int worker = blockIdx.x / CG;        // CTA (pair) index
  const int n_workers = gridDim.x / CG;
  const int tiles = p.m_tiles * p.n_tiles;
  const int total_items = tiles * p.split_k;
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
      mbar_init(&tmem_empty[i], GEMM_EPI_WARPS * CG);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      tmem_alloc_cg2(tmem_ptr, 2 * BN);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_ptr, 2 * BN);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();    // everything above overlapped the previous kernel's tail; operands are touched only from here on
  pdl_launch();

  if (warp == 0) {
    // ===================== TMA producer (every CTA stages its own A rows and its share of B) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = worker; item < total_items; item += n_workers) {
        // k-split major: the tiles that run concurrently walk the SAME k-range, so every A / B panel of a
        // split-K wgrad is pulled from HBM once and shared through L2
        const int ks = item / tiles;
        const int tile = item - ks * tiles;
        const int m0 = (tile / p.n_tiles) * (GEMM_BM * CG) + static_cast<int>(cta_rank) * GEMM_BM;
        const int n0 = (tile % p.n_tiles) * BN + static_cast<int>(cta_rank) * BNL;
        int kb0, kb1;
        gemm_split_range<EPI>(p, ks, total_kb, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* da = smem_a + stage * S::A_BYTES;
          uint8_t* db = smem_b + stage * S::B_BYTES;
          if constexpr (CG == 2) {
            // completion bytes of BOTH CTAs are posted on the leader's barrier
            const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            const bool skip_b = CDR_DBG(p, 8) && (kb & 1);  // feed experiment: odd k-blocks reuse stale B tiles
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], skip_b ? 2 * S::A_BYTES : 2 * S::STAGE_BYTES);
            if constexpr (A_MN) {
#pragma unroll
              for (int c = 0; c < GEMM_BM / 64; ++c)
                tma_load_2d_cg2(da + c * (GEMM_BK * 128), &tma_a, bar, m0 + c * 64, kb * GEMM_BK);
            } else {
              tma_load_2d_cg2(da, &tma_a, bar, kb * GEMM_BK, m0);
            }
            if (skip_b) {
            } else if constexpr (B_MN) {
#pragma unroll
              for (int c = 0; c < BNL / 64; ++c)
                tma_load_2d_cg2(db + c * (GEMM_BK * 128), &tma_b, bar, n0 + c * 64, kb * GEMM_BK);
            } else {
              tma_load_2d_cg2(db, &tma_b, bar, kb * GEMM_BK, n0);
            }
          } else {
            mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
            if constexpr (A_MN) {
#pragma unroll
              for (int c = 0; c < GEMM_BM / 64; ++c)
                tma_load_2d(da + c * (GEMM_BK * 128), &tma_a, &full_bar[stage], m0 + c * 64, kb * GEMM_BK);
            } else {
              tma_load_2d(da, &tma_a, &full_bar[stage], kb * GEMM_BK, m0);
            }
            if constexpr (B_MN) {
#pragma unroll
              for (int c = 0; c < BNL / 64; ++c)
                tma_load_2d(db + c * (GEMM_BK * 128), &tma_b, &full_bar[stage], n0 + c * 64, kb * GEMM_BK);
            } else {
              tma_load_2d(db, &tma_b, &full_bar[stage], kb * GEMM_BK, n0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair leader only) =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(GEMM_BM * CG, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // descriptor geometry
      const uint32_t a_lbo = A_MN ? CDR_DBG_OR(p.dbg_lbo, GEMM_BK * 128) : 16;
      const uint32_t b_lbo = B_MN ? CDR_DBG_OR(p.dbg_lbo, GEMM_BK * 128) : 16;
      const uint32_t sbo = CDR_DBG_OR(p.dbg_sbo, 1024);
      constexpr uint32_t a_kstep = A_MN ? 2048 : 32;   // bytes per UMMA_K = 16
      constexpr uint32_t b_kstep = B_MN ? 2048 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = worker; item < total_items; item += n_workers) {
        const int ks = item / tiles;
        int kb0, kb1;
        gemm_split_range<EPI>(p, ks, total_kb, kb0, kb1);
        if constexpr (EPI == CDR_EPI_F32_GROUPED) {
          if (kb1 <= kb0) continue;  // absent group: no accumulator is produced (the epilogue skips it too)
        }
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
            if constexpr (CG == 2) tc_mma_f16_cg2(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else tc_mma_f16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs retire
          if constexpr (CG == 2) tc_commit_cg2(&empty_bar[stage], 3); else tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (CG == 2) tc_commit_cg2(&tmem_full[acc], 3); else tc_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;             // 0..GEMM_EPI_WARPS-1
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int half = ew >> 2;            // which share of the column chunks this warp handles
    float* stg = stg_all + ew * (32 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t empty_remote = (CG == 2) ? mapa_shared(smem_u32(&tmem_empty[0]), 0) : 0u;
    // the grouped wgrad runs the fp32 accumulate-add epilogue on a per-group output base
    constexpr bool GROUPED = EPI == CDR_EPI_F32_GROUPED;
    constexpr int EPIM = GROUPED ? CDR_EPI_F32_ATOMIC : EPI;
    struct NoCopy {};
    std::conditional_t<GROUPED, GemmParams, NoCopy> pg;  // private copy (per-item `out`) for the grouped wgrad only
    if constexpr (GROUPED) pg = p;
    const GemmParams& pe = [&]() -> const GemmParams& {
      if constexpr (GROUPED) return pg; else return p;
    }();
    DropCtx dc{};
    if constexpr (EPI == CDR_EPI_BIAS_DROP_RESIDUAL) dc = drop_load(p.drop);
    for (int item = worker; item < total_items; item += n_workers) {
      const int tile = item % tiles;
      if constexpr (EPI == CDR_EPI_F32_GROUPED) {
        const int ks = item / tiles;
        int kb0, kb1;
        gemm_split_range<EPI>(p, ks, total_kb, kb0, kb1);
        if (kb1 <= kb0) continue;
        pg.out = reinterpret_cast<float*>(p.out) + static_cast<long long>(ks) * p.seg_out_stride;
      }
      const int m0 = (tile / p.n_tiles) * (GEMM_BM * CG) + static_cast<int>(cta_rank) * GEMM_BM;
      const int n0 = (tile % p.n_tiles) * BN;
      // this warp's chunks: columns (half + NH*j) * 32 of TMEM lane quadrant `quad`, handled two at a time
      constexpr int NH = GEMM_EPI_WARPS / 4;          // warps per lane quadrant
      constexpr int CPW = (BN / 32) / NH;             // chunks per warp per tile
      static_assert(CPW == 1 || CPW % 2 == 0, "chunks per epilogue warp");
      const int mb = m0 + quad * 32;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
      uint4 aux[CPW][4];
#pragma unroll
      for (int j = 0; j < CPW; ++j)  // all in flight while the MMAs of this tile finish
        gemm_epilogue_load_aux<EPIM>(p, aux[j], lane, mb, n0 + (half + NH * j) * 32);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll
      for (int b = 0; b < CPW; b += 2) {
        const int c0 = half + NH * b, c1 = half + NH * (b + 1);
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_row + c0 * 32, r0);
        if constexpr (CPW >= 2) tmem_ld_32x32(t_row + c1 * 32, r1);
        tc_wait_ld();
        if (b + 2 >= CPW) {
          // the accumulator stage is free as soon as it sits in registers: release it BEFORE the epilogue
          // math and the global stores, so the MMAs of the tile after next never wait for this warp's stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_cluster_relaxed(empty_remote + acc * 8);
            else mbar_arrive(&tmem_empty[acc]);
          }
        }
        if constexpr (EPI == CDR_EPI_SCAN_FILTER_Q) {
          const int q = mb + lane;
          const float th = q < p.M ? __ldg(p.thresh + q) : INFINITY;  // rows beyond the queries admit nothing
          const uint32_t m0 = gemm_filter_rows_mask(p, r0, n0 + c0 * 32, th);
          uint32_t m1 = 0u;
          if constexpr (CPW >= 2) m1 = gemm_filter_rows_mask(p, r1, n0 + c1 * 32, th);
          if ((m0 | m1) != 0u) {
            const int base = atomicAdd(p.cand_count + q, __popc(m0) + __popc(m1));
            gemm_filter_rows_commit(p, r0, q, n0 + c0 * 32, m0, base);
            if constexpr (CPW >= 2) gemm_filter_rows_commit(p, r1, q, n0 + c1 * 32, m1, base + __popc(m0));
          }
        } else if constexpr (EPI == CDR_EPI_SCAN_FILTER) {
          // both chunks reserve first (their atomics are in flight together), then both commit
          uint32_t* fws = reinterpret_cast<uint32_t*>(stg);
          const bool ok0 = n0 + c0 * 32 < p.N, ok1 = (CPW >= 2) && (n0 + c1 * 32 < p.N);
          FilterTicket t0{}, t1{};
          if (ok0) t0 = gemm_filter_reserve(p, r0, fws, lane, mb + lane, n0 + c0 * 32);
          if (ok1) t1 = gemm_filter_reserve(p, r1, fws + 96, lane, mb + lane, n0 + c1 * 32);
          if (ok0) gemm_filter_commit(p, r0, fws, lane, mb + lane, n0 + c0 * 32, t0);
          if (ok1) gemm_filter_commit(p, r1, fws + 96, lane, mb + lane, n0 + c1 * 32, t1);
        } else if (n0 + c0 * 32 < p.N && !CDR_DBG(p, 1)) {
          gemm_epilogue_chunk<EPIM>(pe, r0, stg, lane, mb, n0 + c0 * 32, aux[b], dc);
          if constexpr (CPW >= 2) {
            if (n0 + c1 * 32 < p.N) gemm_epilogue_chunk<EPIM>(pe, r1, stg, lane, mb, n0 + c1 * 32, aux[b + 1 < CPW ? b + 1 : b], dc);
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, 2 * BN); else tmem_dealloc(tmem_base, 2 * BN);
  }
}

}  // namespace cdr
