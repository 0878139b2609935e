// ANN negative mining and query clustering on the device (SURVEY f-2): the steps on either side of the corpus scan
// in the reference's ANN data generation (ANCE/drivers/run_ann_data_gen.py:306-429, 497-570), which there run as
// Python loops over numpy arrays after a 27 GB round trip through pickles.
//
//   mine_negatives_kernel : per query, from the scan's top-k document rows:
//        rr      = 1 / rank of the positive passage among ALL k results, 0 if absent            (:523-534)
//        neg[]   = the first n_neg distinct passage ids != positive, walking the candidate list in `order`
//                  (identity = SelectTopK, a permutation = the shuffled branch)                  (:536-563)
//   kmeans_assign_kernel  : group id = argmax_c (q . c - |c|^2 / 2) from the tcgen05 score matrix  (faiss Kmeans /
//                           IndexFlatL2.search(q, 1) of :340-351; ties -> lowest centroid index)
//   kmeans_accum_kernel   : per-centroid vector sums and counts for the Lloyd update
#include "cdr_common.cuh"

namespace cdr {

constexpr int MINE_MAX_NEG = 256;

// one warp per query
__global__ void __launch_bounds__(128)
mine_negatives_kernel(const long long* __restrict__ I, int n_q, int k, const long long* __restrict__ doc_pid,
                      long long n_docs, const long long* __restrict__ pos_pid, const int* __restrict__ order, int n_sel,
                      int n_neg, float* __restrict__ rr, long long* __restrict__ neg, int* __restrict__ neg_count) {
  __shared__ long long chosen[4][MINE_MAX_NEG];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + warp;
  if (q >= n_q) return;
  const long long pos = pos_pid[q];
  const long long* row = I + static_cast<long long>(q) * k;
  // ---- reciprocal rank of the positive over all k results
  int first = k;
  for (int j0 = 0; j0 < k; j0 += 32) {
    const int j = j0 + lane;
    bool hit = false;
    if (j < k) {
      const long long d = row[j];
      hit = d >= 0 && d < n_docs && doc_pid[d] == pos;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m != 0u) {
      first = j0 + __ffs(m) - 1;
      break;
    }
  }
  if (lane == 0) rr[q] = first < k ? 1.0f / static_cast<float>(first + 1) : 0.f;
  // ---- negatives: walk the selected candidates in order, skip the positive and repeats
  long long* mine = chosen[warp];
  int cnt = 0;
  for (int s0 = 0; s0 < n_sel && cnt < n_neg; s0 += 32) {
    const int s = s0 + lane;
    long long pid = -1;
    if (s < n_sel) {
      const int j = order ? order[static_cast<long long>(q) * n_sel + s] : s;
      if (j >= 0 && j < k) {
        const long long d = row[j];
        if (d >= 0 && d < n_docs) pid = doc_pid[d];
      }
    }
    // sequential semantics over the 32 candidates of this chunk
    for (int t = 0; t < 32 && cnt < n_neg; ++t) {
      const long long p = __shfl_sync(0xffffffffu, pid, t);
      if (s0 + t >= n_sel) break;
      if (p < 0 || p == pos) continue;
      bool dup = false;
      for (int c = lane; c < cnt; c += 32) dup |= (mine[c] == p);
      if (__any_sync(0xffffffffu, dup)) continue;
      if (lane == 0) mine[cnt] = p;
      __syncwarp();
      ++cnt;
    }
  }
  __syncwarp();
  for (int c = lane; c < n_neg; c += 32) neg[static_cast<long long>(q) * n_neg + c] = c < cnt ? mine[c] : -1;
  if (lane == 0) neg_count[q] = cnt;
}

// scores [n, ld] fp32 = X C^T (first k columns valid); half_sq[c] = |c|^2 / 2
__global__ void __launch_bounds__(256)
kmeans_assign_kernel(const float* __restrict__ scores, long long ld, const float* __restrict__ half_sq, long long n,
                     int k, int* __restrict__ assign) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* row = scores + i * ld;
  float best = -INFINITY;
  int arg = 0;
  for (int c = 0; c < k; ++c) {
    const float v = row[c] - half_sq[c];
    if (v > best) {
      best = v;
      arg = c;
    }
  }
  assign[i] = arg;
}

// sums[g, :] += x[i, :] (fp16 -> fp32), counts[g] += 1 for the rows of this block; blockDim.x * 8 >= dim
__global__ void __launch_bounds__(128)
kmeans_accum_kernel(const __half* __restrict__ x, const int* __restrict__ assign, long long n, int dim, int k,
                    int rows_per_block, float* __restrict__ sums, float* __restrict__ counts) {
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
  const int c = threadIdx.x * 8;
  // runs of equal group ids are accumulated in registers before touching the global sums
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int cur = -1, run = 0;
  auto flush = [&]() {
    if (cur >= 0 && c < dim) {
#pragma unroll
      for (int t = 0; t < 8; ++t) atomicAdd(sums + static_cast<long long>(cur) * dim + c + t, acc[t]);
      if (threadIdx.x == 0) atomicAdd(counts + cur, static_cast<float>(run));
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = 0.f;
    run = 0;
  };
  for (long long r = r0; r < r1; ++r) {
    const int g = assign[r];
    if (g != cur) {
      flush();
      cur = (g >= 0 && g < k) ? g : -1;
    }
    if (cur >= 0 && c < dim) {
      const uint4 qv = *reinterpret_cast<const uint4*>(x + r * dim + c);
      const __half2* h = reinterpret_cast<const __half2*>(&qv);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(h[t]);
        acc[2 * t] += f.x;
        acc[2 * t + 1] += f.y;
      }
    }
    ++run;
  }
  flush();
}

}  // namespace cdr

using namespace cdr;

extern "C" {

int cdr_mine_negatives(const int64_t* I, int32_t n_q, int32_t k, const int64_t* doc_pid, int64_t n_docs,
                       const int64_t* pos_pid, const int32_t* order, int32_t n_sel, int32_t n_neg, float* rr,
                       int64_t* neg, int32_t* neg_count, void* stream) {
  CDR_REQUIRE(I && doc_pid && pos_pid && rr && neg && neg_count, "cdr_mine_negatives: null pointer");
  CDR_REQUIRE(n_q >= 0 && k > 0 && n_sel > 0 && n_sel <= k && n_neg > 0 && n_neg <= MINE_MAX_NEG,
              "cdr_mine_negatives: need 0 < n_sel <= k and 0 < n_neg <= %d", MINE_MAX_NEG);
  if (n_q == 0) return CDR_OK;
  mine_negatives_kernel<<<(n_q + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(I), n_q, k, reinterpret_cast<const long long*>(doc_pid), n_docs,
      reinterpret_cast<const long long*>(pos_pid), order, n_sel, n_neg, rr, reinterpret_cast<long long*>(neg), neg_count);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_kmeans_assign(const float* scores, int64_t ld, const float* half_sq, int64_t n, int32_t k, int32_t* assign,
                      void* stream) {
  CDR_REQUIRE(scores && half_sq && assign && n >= 0 && k > 0 && ld >= k, "cdr_kmeans_assign: bad arguments");
  if (n == 0) return CDR_OK;
  kmeans_assign_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      scores, ld, half_sq, n, k, assign);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_kmeans_accumulate(const void* x, const int32_t* assign, int64_t n, int32_t dim, int32_t k, float* sums,
                          float* counts, void* stream) {
  CDR_REQUIRE(x && assign && sums && counts && n >= 0 && k > 0, "cdr_kmeans_accumulate: bad arguments");
  CDR_REQUIRE(dim > 0 && dim % 8 == 0 && dim <= 1024, "cdr_kmeans_accumulate: dim must be a multiple of 8, <= 1024");
  if (n == 0) return CDR_OK;
  const int rows_per_block = 256;
  kmeans_accum_kernel<<<static_cast<unsigned>((n + rows_per_block - 1) / rows_per_block), 128, 0,
                        static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(x), assign, n, dim, k,
                                                             rows_per_block, sums, counts);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

}  // extern "C"
