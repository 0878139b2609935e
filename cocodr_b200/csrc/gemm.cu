// cdr_gemm: host-side validation, TMA map construction and kernel dispatch for the tcgen05 GEMM.
#include <stdlib.h>

#include "gemm_sm100.cuh"
#include "tma_host.h"

namespace cdr {

// CTA-pair (cta_group::2) tiles are used whenever the problem has a 256-wide N tile and more than one
// 128-row M tile; CDR_GEMM_2CTA=0 in the environment forces the single-CTA kernel (A/B comparisons).
static bool use_2cta_default() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CDR_GEMM_2CTA");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CG>
static int launch_gemm_cg(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, EPI, CG>;
  constexpr int smem = GemmSmem<BN, CG, GEMM_EW<EPI>>::TOTAL;
  static bool configured = false;  // per-instantiation; attribute is sticky per device context
  if (!configured) {
    CDR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int items = p.m_tiles * p.n_tiles * p.split_k;
  const int workers = sm_count() / CG;
  const int grid = (items < workers ? items : workers) * CG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS<EPI>);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  CDR_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  return CDR_OK;
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  if constexpr (BN == 256) {
    if (p.cta_group == 2) return launch_gemm_cg<BN, A_MN, B_MN, EPI, 2>(ta, tb, p, st);
  }
  return launch_gemm_cg<BN, A_MN, B_MN, EPI, 1>(ta, tb, p, st);
}

template <int BN, bool A_MN, bool B_MN>
static int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  switch (epi) {
    case CDR_EPI_STORE_F16: return launch_gemm<BN, A_MN, B_MN, CDR_EPI_STORE_F16>(ta, tb, p, st);
    case CDR_EPI_BIAS_RESIDUAL: return launch_gemm<BN, A_MN, B_MN, CDR_EPI_BIAS_RESIDUAL>(ta, tb, p, st);
    case CDR_EPI_F32_STORE: return launch_gemm<BN, A_MN, B_MN, CDR_EPI_F32_STORE>(ta, tb, p, st);
    default: break;
  }
  if constexpr (!A_MN && !B_MN) {
    if (epi == CDR_EPI_BIAS_DROP_RESIDUAL) return launch_gemm<BN, false, false, CDR_EPI_BIAS_DROP_RESIDUAL>(ta, tb, p, st);
    if (epi == CDR_EPI_BIAS_GELU) return launch_gemm<BN, false, false, CDR_EPI_BIAS_GELU>(ta, tb, p, st);
    if (epi == CDR_EPI_SCAN_FILTER) return launch_gemm<BN, false, false, CDR_EPI_SCAN_FILTER>(ta, tb, p, st);
    if (epi == CDR_EPI_SCAN_FILTER_Q) return launch_gemm<BN, false, false, CDR_EPI_SCAN_FILTER_Q>(ta, tb, p, st);
  }
  if constexpr (!A_MN && B_MN) {
    if (epi == CDR_EPI_DGELU) return launch_gemm<BN, false, true, CDR_EPI_DGELU>(ta, tb, p, st);
  }
  if constexpr (A_MN && B_MN) {
    if (epi == CDR_EPI_F32_ATOMIC) return launch_gemm<BN, true, true, CDR_EPI_F32_ATOMIC>(ta, tb, p, st);
    if (epi == CDR_EPI_F32_GROUPED) return launch_gemm<BN, true, true, CDR_EPI_F32_GROUPED>(ta, tb, p, st);
  }
  set_error("cdr_gemm: epilogue %d not instantiated for layout a_major=%d b_major=%d", epi, (int)A_MN, (int)B_MN);
  return CDR_EINVAL;
}

template <int BN>
static int dispatch_layout(int a_mn, int b_mn, int epi, const CUtensorMap& ta, const CUtensorMap& tb,
                           const GemmParams& p, cudaStream_t st) {
  if (!a_mn && !b_mn) return dispatch_epi<BN, false, false>(epi, ta, tb, p, st);
  if (!a_mn && b_mn) return dispatch_epi<BN, false, true>(epi, ta, tb, p, st);
  if (a_mn && b_mn) return dispatch_epi<BN, true, true>(epi, ta, tb, p, st);
  set_error("cdr_gemm: layout a_major=1,b_major=0 not instantiated");
  return CDR_EINVAL;
}

int gemm_run(const cdr_gemm_args& g, GemmParams p, cudaStream_t st) {
  CDR_REQUIRE(g.a && g.b, "cdr_gemm: null operand");
  CDR_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "cdr_gemm: empty problem M=%lld N=%lld K=%lld", (long long)g.M,
              (long long)g.N, (long long)g.K);
  CDR_REQUIRE(g.N % 8 == 0 || g.epilogue == CDR_EPI_SCAN_FILTER || g.epilogue == CDR_EPI_SCAN_FILTER_Q,
              "cdr_gemm: N must be a multiple of 8 (N=%lld)",
              (long long)g.N);
  CDR_REQUIRE(g.M < (1ll << 31) && g.N < (1ll << 31) && g.K < (1ll << 31), "cdr_gemm: dimension overflow");
  const int a_mn = g.a_major, b_mn = g.b_major;
  if (a_mn) CDR_REQUIRE(g.M % 64 == 0, "cdr_gemm: MN-major A needs M %% 64 == 0 (M=%lld)", (long long)g.M);
  if (b_mn) CDR_REQUIRE(g.N % 64 == 0, "cdr_gemm: MN-major B needs N %% 64 == 0 (N=%lld)", (long long)g.N);
  const int BN = (g.N > 128) ? 256 : 128;
  p.M = (int)g.M; p.N = (int)g.N; p.K = (int)g.K;
  p.cta_group = (BN == 256 && g.M > GEMM_BM && use_2cta_default()) ? 2 : 1;
  p.m_tiles = (p.M + GEMM_BM * p.cta_group - 1) / (GEMM_BM * p.cta_group);
  p.n_tiles = (p.N + BN - 1) / BN;
  const int total_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  int split = g.split_k;
  if (g.epilogue == CDR_EPI_F32_GROUPED) {
    split = p.split_k;  // one split per group, ranges from p.seg_kb (cdr_gemm_grouped)
  } else if (split <= 0) {
    // auto: minimise rounds x (k-blocks per item + a fixed per-item cost for pipeline ramp and the exposed
    // part of the epilogue), i.e. prefer splits whose item count fills whole waves of the persistent grid
    const int tiles = p.m_tiles * p.n_tiles;
    const int workers = sm_count() / p.cta_group;
    long best_cost = -1;
    split = 1;
    for (int s = 1; s <= 64 && s * 8 <= total_kb; ++s) {
      const int kbps = (total_kb + s - 1) / s;
      const int s_eff = (total_kb + kbps - 1) / kbps;
      if (s_eff != s) continue;
      const long rounds = (static_cast<long>(tiles) * s + workers - 1) / workers;
      const long cost = rounds * (kbps + 4);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        split = s;
      }
    }
  }
  if (g.epilogue == CDR_EPI_F32_GROUPED) {
    p.kb_per_split = total_kb;
    p.split_k = split;
  } else {
    if (split > total_kb) split = total_kb;
    p.kb_per_split = (total_kb + split - 1) / split;
    p.split_k = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
    CDR_REQUIRE(p.split_k == 1 || g.epilogue == CDR_EPI_F32_ATOMIC, "cdr_gemm: split_k > 1 needs CDR_EPI_F32_ATOMIC");
  }
  p.alpha = g.alpha;
  p.colsum = g.colsum;
  p.colsum_scale = g.colsum_scale;
#ifdef CDR_GEMM_DEBUG  // experiment builds only (tools/build_variant.sh): descriptor overrides, work-skipping flags
  p.dbg_lbo = g.dbg_lbo; p.dbg_sbo = g.dbg_sbo;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("CDR_GEMM_DBG");
      dbg = e ? atoi(e) : 0;
    }
    p.dbg_flags = dbg;
  }
#endif

  CUtensorMap ta, tb;
  int rc;
  if (a_mn) rc = make_tma_2d_f16(&ta, g.a, g.M, g.K, g.lda, 64, GEMM_BK);
  else rc = make_tma_2d_f16(&ta, g.a, g.K, p.a_rows_alloc > g.M ? p.a_rows_alloc : g.M, g.lda, GEMM_BK, GEMM_BM);
  if (rc != CDR_OK) return rc;
  if (b_mn) rc = make_tma_2d_f16(&tb, g.b, g.N, g.K, g.ldb, 64, GEMM_BK);
  else rc = make_tma_2d_f16(&tb, g.b, g.K, p.b_rows_alloc > g.N ? p.b_rows_alloc : g.N, g.ldb, GEMM_BK, BN / p.cta_group);
  if (rc != CDR_OK) return rc;

  if (BN == 256) return dispatch_layout<256>(a_mn, b_mn, g.epilogue, ta, tb, p, st);
  return dispatch_layout<128>(a_mn, b_mn, g.epilogue, ta, tb, p, st);
}

}  // namespace cdr

extern "C" int cdr_gemm(const cdr_gemm_args* g, void* stream) {
  if (g == nullptr) {
    cdr::set_error("cdr_gemm: null args");
    return CDR_EINVAL;
  }
  CDR_REQUIRE(g->epilogue != CDR_EPI_SCAN_FILTER && g->epilogue != CDR_EPI_SCAN_FILTER_Q,
              "cdr_gemm: the scan filter epilogues are internal to cdr_scan_topk");
  CDR_REQUIRE(g->epilogue != CDR_EPI_F32_GROUPED, "cdr_gemm: the grouped epilogue is internal to cdr_gemm_grouped");
  CDR_REQUIRE(g->out != nullptr, "cdr_gemm: null output");
  const bool f32 = g->epilogue == CDR_EPI_F32_ATOMIC || g->epilogue == CDR_EPI_F32_STORE;
  CDR_REQUIRE(g->ldo % (f32 ? 4 : 8) == 0, "cdr_gemm: ldo must keep rows 16-byte aligned (ldo=%lld)", (long long)g->ldo);
  CDR_REQUIRE((reinterpret_cast<uintptr_t>(g->out) & 15) == 0, "cdr_gemm: out must be 16-byte aligned");
  if (g->epilogue == CDR_EPI_BIAS_DROP_RESIDUAL) {
    CDR_REQUIRE(g->drop.state != nullptr && g->drop.threshold > 0 && g->drop.threshold < 65536 && g->drop.row_mul >= 0,
                "cdr_gemm: CDR_EPI_BIAS_DROP_RESIDUAL needs drop.state and 0 < drop.threshold < 65536");
    CDR_REQUIRE(g->M * (g->drop.row_mul > 0 ? g->drop.row_mul : 1) * (g->N / 8) < (1ll << 32),
                "cdr_gemm: dropout group index overflows 32 bits");
    CDR_REQUIRE(g->drop.keep_bits == nullptr || (g->N % 32 == 0 && (reinterpret_cast<uintptr_t>(g->drop.keep_bits) & 3) == 0),
                "cdr_gemm: drop.keep_bits needs N %% 32 == 0 and a 4-byte aligned buffer");
  }
  if (g->epilogue == CDR_EPI_BIAS_RESIDUAL || g->epilogue == CDR_EPI_DGELU || g->epilogue == CDR_EPI_BIAS_DROP_RESIDUAL) {
    CDR_REQUIRE(g->aux != nullptr && g->ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(g->aux) & 15) == 0,
                "cdr_gemm: epilogue %d needs a 16-byte aligned aux operand", g->epilogue);
  }
  if (g->colsum != nullptr)
    CDR_REQUIRE(g->epilogue == CDR_EPI_DGELU, "cdr_gemm: the fused column sum is only available with CDR_EPI_DGELU");
  if (g->bias) CDR_REQUIRE((reinterpret_cast<uintptr_t>(g->bias) & 15) == 0, "cdr_gemm: bias must be 16-byte aligned");
  cdr::GemmParams p{};
  p.out = g->out;
  p.out2 = g->out2;
  p.bias = g->bias;
  p.aux = static_cast<const __half*>(g->aux);
  p.ldo = g->ldo;
  p.ldaux = g->ldaux;
  p.drop = g->drop;
  return cdr::gemm_run(*g, p, static_cast<cudaStream_t>(stream));
}

extern "C" int cdr_gemm_segments(const cdr_gemm_args* base, int32_t n_seg, const int64_t* row_begin,
                                 const int64_t* row_count, const int64_t* out_offset, void* stream) {
  if (base == nullptr || n_seg < 0 || (n_seg > 0 && (!row_begin || !row_count || !out_offset))) {
    cdr::set_error("cdr_gemm_segments: null argument");
    return CDR_EINVAL;
  }
  CDR_REQUIRE(base->a_major == 1 && base->b_major == 1,
              "cdr_gemm_segments: both operands must be MN-major (the segmented axis is K)");
  CDR_REQUIRE(base->epilogue == CDR_EPI_F32_ATOMIC || base->epilogue == CDR_EPI_F32_STORE,
              "cdr_gemm_segments: fp32 epilogues only");
  for (int32_t i = 0; i < n_seg; ++i) {
    CDR_REQUIRE(row_begin[i] >= 0 && row_count[i] >= 0 && out_offset[i] >= 0, "cdr_gemm_segments: negative entry %d", i);
    if (row_count[i] == 0) continue;
    cdr_gemm_args g = *base;
    g.a = static_cast<const __half*>(base->a) + row_begin[i] * base->lda;
    g.b = static_cast<const __half*>(base->b) + row_begin[i] * base->ldb;
    g.out = static_cast<float*>(base->out) + out_offset[i];
    g.K = row_count[i];
    const int rc = cdr_gemm(&g, stream);
    if (rc != CDR_OK) return rc;
  }
  return CDR_OK;
}

extern "C" int cdr_gemm_grouped(const cdr_gemm_args* base, int32_t n_groups, const int32_t* seg_kb,
                                int64_t out_group_stride, void* stream) {
  if (base == nullptr || seg_kb == nullptr) {
    cdr::set_error("cdr_gemm_grouped: null argument");
    return CDR_EINVAL;
  }
  CDR_REQUIRE(n_groups > 0 && n_groups <= 65536, "cdr_gemm_grouped: n_groups=%d out of range", n_groups);
  CDR_REQUIRE(base->a_major == 1 && base->b_major == 1,
              "cdr_gemm_grouped: both operands must be MN-major (the grouped axis is K)");
  CDR_REQUIRE(base->epilogue == CDR_EPI_F32_ATOMIC, "cdr_gemm_grouped: CDR_EPI_F32_ATOMIC only");
  CDR_REQUIRE(base->out != nullptr && (reinterpret_cast<uintptr_t>(base->out) & 15) == 0 && base->ldo % 4 == 0 &&
                  out_group_stride % 4 == 0 && out_group_stride >= 0,
              "cdr_gemm_grouped: out, ldo and the group stride must keep rows 16-byte aligned");
  cdr_gemm_args g = *base;
  g.epilogue = CDR_EPI_F32_GROUPED;
  cdr::GemmParams p{};
  p.out = g.out;
  p.ldo = g.ldo;
  p.split_k = n_groups;
  p.seg_kb = seg_kb;
  p.seg_out_stride = out_group_stride;
  return cdr::gemm_run(g, p, static_cast<cudaStream_t>(stream));
}

