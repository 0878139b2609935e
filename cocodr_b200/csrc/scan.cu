// K15 corpus scan: exact top-k inner-product search over HBM-resident fp16 document embeddings.
//
//   1. thresholds  : score a strided sample of S documents against every query (tcgen05 GEMM, fp32 out)
//                    and take the m-th largest sample score per query as the admission threshold
//                    (m chosen so that ~3k documents of the full corpus are expected above it)
//   2. filter pass : ONE pass over the corpus -- tcgen05 GEMM, documents on M, queries on N -- whose
//                    epilogue admits (score >= threshold[q]) into per-query candidate buffers as packed
//                    order-preserving (score, ~doc) 64-bit keys.  Scores never touch HBM.
//   3. select      : per query, bitonic sort of the candidates in shared memory, emit the first k.
//                    If >= k candidates were admitted (and the buffer did not overflow) these are
//                    exactly the global top-k in (score desc, doc asc) order; otherwise status[0]++.
//
// Corpora small enough to fit the candidate buffer skip (1) and admit everything.
#include "gemm_sm100.cuh"

namespace cdr {

constexpr int SCAN_CAP_MIN = 8192;     // candidate slots per query (>= 8k)
constexpr int SCAN_SAMPLE = 8192;      // sampled documents for the thresholds
constexpr int SCAN_SORT_MAX = 16384;   // bitonic sort capacity (128 KB of keys)

__host__ __device__ inline int scan_cap(int k) {
  int c = SCAN_CAP_MIN;
  while (c < 8 * k && c < SCAN_SORT_MAX) c <<= 1;
  return c;
}

__device__ __forceinline__ float unflip_score(unsigned int u) {
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  return __uint_as_float(u);
}

// descending bitonic sort of n (power of two) 64-bit keys in shared memory, whole block
__device__ void bitonic_sort_desc(unsigned long long* s, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = s[lo], b = s[hi];
        if ((a < b) == desc) {
          s[lo] = b;
          s[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

__host__ __device__ __forceinline__ int next_pow2(int v) {
  int n = 2;
  while (n < v) n <<= 1;
  return n;
}

__device__ __forceinline__ unsigned int flip_score(float s) {
  const unsigned int u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// thresh[q] = m-th largest of sample[q, 0:S]: MSB-first radix select (4 passes of 8 bits) on the
// order-preserving integer image of the scores -- no sort, the row is re-read from L2 each pass.
__global__ void __launch_bounds__(256)
scan_threshold_kernel(const float* __restrict__ sample, int S, int m, float* __restrict__ thresh) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int sel_prefix, sel_rank;
  const float* row = sample + static_cast<long long>(blockIdx.x) * S;
  unsigned int prefix = 0, mask = 0, rank = static_cast<unsigned int>(m);
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
      const unsigned int u = flip_score(row[i]);
      if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int r = rank;
      int bin = 255;
      for (; bin > 0; --bin) {
        if (hist[bin] >= r) break;
        r -= hist[bin];
      }
      sel_prefix = prefix | (static_cast<unsigned int>(bin) << shift);
      sel_rank = r;
    }
    __syncthreads();
    prefix = sel_prefix;
    rank = sel_rank;
    mask |= 0xFFu << shift;
  }
  if (threadIdx.x == 0) thresh[blockIdx.x] = unflip_score(prefix);
}

__global__ void scan_init_kernel(float* thresh, int* count, int n_q, int fill_thresh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_q) {
    count[i] = 0;
    if (fill_thresh) thresh[i] = -INFINITY;
  }
}

// one block per query: sort the admitted candidates, write the first k
__global__ void __launch_bounds__(1024)
scan_select_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ count, int cap, int k,
                   long long n_docs, long long doc_base, float* __restrict__ out_s, long long* __restrict__ out_i,
                   int* __restrict__ status) {
  extern __shared__ unsigned long long keys[];
  const int q = blockIdx.x;
  const int c_raw = count[q];
  const int c = min(c_raw, cap);
  if (threadIdx.x == 0 && (c_raw > cap || static_cast<long long>(c_raw) < min(static_cast<long long>(k), n_docs)))
    atomicAdd(status, 1);
  const int n = next_pow2(c);
  for (int i = threadIdx.x; i < n; i += blockDim.x) keys[i] = i < c ? cand[static_cast<long long>(q) * cap + i] : 0ull;
  bitonic_sort_desc(keys, n);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    float s = -INFINITY;
    long long id = -1;
    if (i < c) {
      const unsigned long long key = keys[i];
      s = unflip_score(static_cast<unsigned int>(key >> 32));
      id = static_cast<long long>(~static_cast<unsigned int>(key)) + doc_base;
    }
    out_s[static_cast<long long>(q) * k + i] = s;
    out_i[static_cast<long long>(q) * k + i] = id;
  }
}

// merge: n_in (score, id) candidates per query -> top k by (score desc, id asc); ids < 2^32, id < 0 = empty
__global__ void __launch_bounds__(1024)
topk_merge_kernel(const float* __restrict__ scores, const long long* __restrict__ ids, int n_in, int k,
                  float* __restrict__ out_s, long long* __restrict__ out_i) {
  extern __shared__ unsigned long long keys[];
  const int q = blockIdx.x;
  const int n = next_pow2(n_in);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    unsigned long long key = 0ull;
    if (i < n_in) {
      const long long id = ids[static_cast<long long>(q) * n_in + i];
      if (id >= 0) key = pack_score_doc(scores[static_cast<long long>(q) * n_in + i], static_cast<unsigned int>(id));
    }
    keys[i] = key;
  }
  bitonic_sort_desc(keys, n);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    float s = -INFINITY;
    long long id = -1;
    if (i < n_in && keys[i] != 0ull) {
      s = unflip_score(static_cast<unsigned int>(keys[i] >> 32));
      id = static_cast<long long>(~static_cast<unsigned int>(keys[i]));
    }
    out_s[static_cast<long long>(q) * k + i] = s;
    out_i[static_cast<long long>(q) * k + i] = id;
  }
}

struct ScanPlan {
  int cap, S, m;
  long long step;
  bool exhaustive;
  size_t off_thresh, off_count, off_sample, off_cand, total;
};

static ScanPlan scan_plan(long long n_docs, int n_q, int k) {
  ScanPlan p{};
  p.cap = scan_cap(k);
  p.exhaustive = n_docs <= p.cap;
  p.S = static_cast<int>(n_docs < SCAN_SAMPLE ? (n_docs / 8) * 8 : SCAN_SAMPLE);
  if (p.S < 8) p.S = 8;
  p.step = n_docs / p.S;
  if (p.step < 1) p.step = 1;
  // expected number of admitted documents: ~3k (at least 1024), i.e. the m-th largest of S samples
  double target = 3.0 * k;
  if (target < 1024.0) target = 1024.0;
  p.m = static_cast<int>(target * p.S / static_cast<double>(n_docs) + 0.999);
  if (p.m < 1) p.m = 1;
  if (p.m > p.S) p.m = p.S;
  auto align = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  size_t o = 0;
  p.off_thresh = o; o = align(o + sizeof(float) * n_q);
  p.off_count = o;  o = align(o + sizeof(int) * n_q);
  p.off_sample = o; o = align(o + (p.exhaustive ? 0 : sizeof(float) * static_cast<size_t>(n_q) * p.S));
  p.off_cand = o;   o = align(o + sizeof(unsigned long long) * static_cast<size_t>(n_q) * p.cap);
  p.total = o;
  return p;
}

template <typename K>
static int set_smem(K kern, int bytes) {
  CDR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return CDR_OK;
}

}  // namespace cdr

using namespace cdr;

extern "C" {

size_t cdr_scan_workspace_bytes(int64_t n_docs, int32_t n_q, int32_t k) {
  if (n_docs <= 0 || n_q <= 0 || k <= 0) return 0;
  return scan_plan(n_docs, n_q, k).total;
}

int64_t cdr_scan_exhaustive_docs(int32_t k) { return scan_cap(k > 0 ? k : 1); }

int cdr_scan_topk(const cdr_scan_args* a, void* stream) {
  CDR_REQUIRE(a != nullptr, "cdr_scan_topk: null args");
  CDR_REQUIRE(a->docs && a->queries && a->out_scores && a->out_ids && a->workspace && a->status,
              "cdr_scan_topk: null pointer");
  CDR_REQUIRE(a->n_docs > 0 && a->n_q > 0 && a->k > 0 && a->dim > 0, "cdr_scan_topk: empty problem");
  CDR_REQUIRE(a->n_docs < (1ll << 31), "cdr_scan_topk: n_docs per call must be < 2^31");
  CDR_REQUIRE(a->dim % 8 == 0 && a->ld_docs % 8 == 0 && a->ld_docs >= a->dim,
              "cdr_scan_topk: dim and ld_docs must be multiples of 8 (dim=%d ld=%lld)", a->dim, (long long)a->ld_docs);
  CDR_REQUIRE(a->k <= SCAN_SORT_MAX / 8 * 8 && a->k <= scan_cap(a->k), "cdr_scan_topk: k=%d too large (max %d)", a->k,
              SCAN_SORT_MAX);
  const ScanPlan pl = scan_plan(a->n_docs, a->n_q, a->k);
  CDR_REQUIRE(pl.cap >= a->k, "cdr_scan_topk: k=%d exceeds the candidate capacity %d", a->k, pl.cap);
  if (a->workspace_bytes < pl.total) {
    set_error("cdr_scan_topk: workspace %zu < required %zu bytes", a->workspace_bytes, pl.total);
    return CDR_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(a->workspace);
  float* thresh = reinterpret_cast<float*>(ws + pl.off_thresh);
  int* count = reinterpret_cast<int*>(ws + pl.off_count);
  float* sample = reinterpret_cast<float*>(ws + pl.off_sample);
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(ws + pl.off_cand);

  scan_init_kernel<<<(a->n_q + 255) / 256, 256, 0, st>>>(thresh, count, a->n_q, pl.exhaustive ? 1 : 0);
  CDR_LAUNCH_CHECK();

  if (!pl.exhaustive) {
    // sample[q, s] = <query q, doc s*step>
    cdr_gemm_args g{};
    g.a = a->queries; g.b = a->docs; g.out = sample;
    g.M = a->n_q; g.N = pl.S; g.K = a->dim;
    g.lda = a->dim; g.ldb = pl.step * a->ld_docs; g.ldo = pl.S;
    g.epilogue = CDR_EPI_F32_STORE; g.split_k = 1; g.alpha = 1.f;
    GemmParams p{};
    p.out = sample; p.ldo = pl.S;
    if (int rc = gemm_run(g, p, st)) return rc;
    scan_threshold_kernel<<<a->n_q, 256, 0, st>>>(sample, pl.S, pl.m, thresh);
    CDR_LAUNCH_CHECK();
  }
  {
    cdr_gemm_args g{};
    g.a = a->docs; g.b = a->queries;
    g.M = a->n_docs; g.N = a->n_q; g.K = a->dim;
    g.lda = a->ld_docs; g.ldb = a->dim;
    g.epilogue = CDR_EPI_SCAN_FILTER; g.split_k = 1; g.alpha = 1.f;
    GemmParams p{};
    p.thresh = thresh; p.cand = cand; p.cand_count = count; p.cand_cap = pl.cap; p.row_base = 0;
    if (int rc = gemm_run(g, p, st)) return rc;
  }
  static bool cfg_s = false;
  if (!cfg_s) {
    if (int rc = set_smem(scan_select_kernel, SCAN_SORT_MAX * 8)) return rc;
    cfg_s = true;
  }
  scan_select_kernel<<<a->n_q, 1024, pl.cap * 8, st>>>(cand, count, pl.cap, a->k, a->n_docs, a->doc_base,
                                                       a->out_scores, reinterpret_cast<long long*>(a->out_ids),
                                                       a->status);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_topk_merge(const float* scores, const int64_t* ids, int32_t n_q, int32_t n_in, int32_t k, float* out_scores,
                   int64_t* out_ids, void* stream) {
  CDR_REQUIRE(scores && ids && out_scores && out_ids, "cdr_topk_merge: null pointer");
  CDR_REQUIRE(n_q > 0 && n_in > 0 && k > 0, "cdr_topk_merge: empty problem");
  CDR_REQUIRE(n_in <= SCAN_SORT_MAX, "cdr_topk_merge: at most %d candidates per query (got %d)", SCAN_SORT_MAX, n_in);
  static bool cfg = false;
  if (!cfg) {
    if (int rc = set_smem(topk_merge_kernel, SCAN_SORT_MAX * 8)) return rc;
    cfg = true;
  }
  topk_merge_kernel<<<n_q, 1024, next_pow2(n_in) * 8, static_cast<cudaStream_t>(stream)>>>(
      scores, reinterpret_cast<const long long*>(ids), n_in, k, out_scores, reinterpret_cast<long long*>(out_ids));
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

}  // extern "C"
