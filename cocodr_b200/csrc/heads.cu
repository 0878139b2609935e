// Contrastive heads and DRO statistics (K8, K9/K9', K10, K12).  fp32 CUDA-core kernels: these ops are
// latency-bound (<= 1 GFLOP, <= 2 MB) except the Gram matrix, which streams the [G, P_last] gradient
// matrix once from HBM.  Logits stay fp32 because trained CLS dot products are ~217 with gaps ~0.3.
#include "cdr_common.cuh"

namespace cdr {

// ------------------------------------------------------------------------------ small fp32 GEMM
// C[m,n] (=|+=) alpha * sum_k A(m,k) * B(k,n),  A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].
// 64x64 tile, 256 threads, 4x4 register micro-tile, K step 16.
constexpr int SG_T = 64, SG_K = 16;

__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K,
             long long sam, long long sak, long long sbk, long long sbn, long long ldc, float alpha) {
  __shared__ float sA[SG_K][SG_T + 4];
  __shared__ float sB[SG_K][SG_T + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += SG_K) {
    // 64x16 elements per operand, 4 per thread; pick the thread->element map that is contiguous in
    // memory for the operand's layout
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int mm, kk;
      if (sak == 1) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      const int gm = m0 + mm, gk = k0 + kk;
      sA[kk][mm] = (gm < M && gk < K) ? A[gm * sam + gk * sak] : 0.f;
      int nn, kb;
      if (sbk == 1) { kb = e & 15; nn = e >> 4; } else { nn = e & 63; kb = e >> 6; }
      const int gn = n0 + nn, gkb = k0 + kb;
      sB[kb][nn] = (gn < N && gkb < K) ? B[gkb * sbk + gn * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) C[m * ldc + n] = alpha * acc[i][j];
    }
  }
}

// Same contract, 16 x 16 outputs per block (one per thread): for problems with a handful of 64 x 64 tiles and a long K
// (the 64 x 64 score matrix of the contrastive head is ONE such tile with K = 768) this spreads the work over 16x
// more blocks.  Deterministic (no split-K atomics): the loss must not depend on the launch.
__global__ void __launch_bounds__(256)
sgemm_small_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K,
                   long long sam, long long sak, long long sbk, long long sbn, long long ldc, float alpha) {
  __shared__ float sA[16][65];  // [m][k]
  __shared__ float sB[16][65];  // [n][k]
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 16, n0 = blockIdx.x * 16;
  const int tm = tid >> 4, tn = tid & 15;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 64) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;  // 16 x 64 elements per operand
      int r, kk;
      if (sak == 1) { kk = e & 63; r = e >> 6; } else { r = e & 15; kk = e >> 4; }
      sA[r][kk] = (m0 + r < M && k0 + kk < K) ? A[(m0 + r) * sam + (k0 + kk) * sak] : 0.f;
      if (sbk == 1) { kk = e & 63; r = e >> 6; } else { r = e & 15; kk = e >> 4; }
      sB[r][kk] = (n0 + r < N && k0 + kk < K) ? B[(k0 + kk) * sbk + (n0 + r) * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll 16
    for (int kk = 0; kk < 64; ++kk) acc = fmaf(sA[tm][kk], sB[tn][kk], acc);
    __syncthreads();
  }
  if (m0 + tm < M && n0 + tn < N) C[(m0 + tm) * ldc + n0 + tn] = alpha * acc;
}

static int sgemm(const float* A, const float* B, float* C, int M, int N, int K, long long sam, long long sak,
                 long long sbk, long long sbn, long long ldc, float alpha, cudaStream_t st) {
  if (((N + SG_T - 1) / SG_T) * ((M + SG_T - 1) / SG_T) < 16 && K >= 256) {
    dim3 g16((N + 15) / 16, (M + 15) / 16);
    sgemm_small_kernel<<<g16, 256, 0, st>>>(A, B, C, M, N, K, sam, sak, sbk, sbn, ldc, alpha);
    CDR_LAUNCH_CHECK();
    return CDR_OK;
  }
  dim3 grid((N + SG_T - 1) / SG_T, (M + SG_T - 1) / SG_T);
  sgemm_kernel<<<grid, 256, 0, st>>>(A, B, C, M, N, K, sam, sak, sbk, sbn, ldc, alpha);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int i = 0; i < nw; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

__device__ __forceinline__ int sim_target(int mode, int gi) { return mode == CDR_SIM_COCO ? (gi ^ 1) : gi; }

// one block per row: mask the diagonal (COCO), log-sum-exp, loss
__global__ void __launch_bounds__(256)
simmat_ce_row_kernel(float* __restrict__ S, float* __restrict__ loss, float* __restrict__ lse, int n_keys, int mode,
                     int row_offset, float loss_scale) {
  __shared__ float red[8];
  const int i = blockIdx.x, gi = row_offset + i;
  float* row = S + static_cast<long long>(i) * n_keys;
  if (mode == CDR_SIM_COCO && threadIdx.x == 0 && gi < n_keys) row[gi] = -INFINITY;
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n_keys; j += blockDim.x) mx = fmaxf(mx, row[j]);
  mx = block_reduce(mx, red, true);
  float s = 0.f;
  for (int j = threadIdx.x; j < n_keys; j += blockDim.x) s += expf(row[j] - mx);
  s = block_reduce(s, red, false);
  if (threadIdx.x == 0) {
    const float l = mx + logf(s);
    lse[i] = l;
    loss[i] = loss_scale * (l - row[sim_target(mode, gi)]);
  }
}

__global__ void __launch_bounds__(256)
simmat_ce_grad_kernel(const float* __restrict__ S, const float* __restrict__ lse, const float* __restrict__ dloss,
                      float* __restrict__ G, int n_keys, int mode, int row_offset, float loss_scale) {
  const int i = blockIdx.x, gi = row_offset + i;
  const float l = lse[i], g = dloss[i] * loss_scale;
  const int t = sim_target(mode, gi);
  const float* row = S + static_cast<long long>(i) * n_keys;
  float* out = G + static_cast<long long>(i) * n_keys;
  for (int j = threadIdx.x; j < n_keys; j += blockDim.x) out[j] = g * (expf(row[j] - l) - (j == t ? 1.f : 0.f));
}

// dk_own[i, :] = dloss[i] * (exp(-loss[i]) - 1) * q[i, :]   (softmax_i,own = exp(s_own - lse_i) = exp(-loss_i))
__global__ void __launch_bounds__(128)
own_key_grad_kernel(const float* __restrict__ q, const float* __restrict__ loss, const float* __restrict__ dloss, int n,
                    int dim, float* __restrict__ dk) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= n) return;
  const float c = dloss[i] * (expf(-loss[i]) - 1.f);
  const long long o = static_cast<long long>(i) * dim;
  for (int d = lane; d < dim; d += 32) dk[o + d] = c * q[o + d];
}

// ------------------------------------------------------------------------------ vocabulary CE (K14)
// MLM head loss (HF BertForMaskedLM: CE(logits.view(-1, V), labels.view(-1)), COCO/modeling.py:87-93) on
// the gathered masked rows only: logits[M, ld] fp32 straight from the decoder GEMM, bias[n_cols] fp32
// added here (padding columns carry -inf).  One block per row.
__global__ void __launch_bounds__(256)
vocab_ce_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ bias,
                    const long long* __restrict__ labels, float* __restrict__ loss, float* __restrict__ lse,
                    int n_cols, long long ld) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  const float* row = logits + static_cast<long long>(i) * ld;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n_cols; j += blockDim.x) mx = fmaxf(mx, row[j] + bias[j]);
  mx = block_reduce(mx, red, true);
  float s = 0.f;
  for (int j = threadIdx.x; j < n_cols; j += blockDim.x) s += __expf(row[j] + bias[j] - mx);
  s = block_reduce(s, red, false);
  if (threadIdx.x == 0) {
    const float l = mx + logf(s);
    const long long t = labels[i];
    lse[i] = l;
    loss[i] = (t >= 0 && t < n_cols) ? l - (row[t] + bias[t]) : 0.f;
  }
}

// dlogits[i, j] = fp16( scale * dloss[i] * (softmax_ij - [j == label_i]) )
__global__ void __launch_bounds__(256)
vocab_ce_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ bias,
                    const long long* __restrict__ labels, const float* __restrict__ lse,
                    const float* __restrict__ dloss, __half* __restrict__ dlogits, int n_cols, long long ld,
                    float scale) {
  const int i = blockIdx.x;
  const float* row = logits + static_cast<long long>(i) * ld;
  __half* out = dlogits + static_cast<long long>(i) * ld;
  const float l = lse[i], g = dloss[i] * scale;
  const long long t = labels[i];
  const bool valid = t >= 0 && t < n_cols;
  for (int j = threadIdx.x * 2; j < n_cols; j += blockDim.x * 2) {  // n_cols is even (padded to 64)
    const float2 v = *reinterpret_cast<const float2*>(row + j);
    const float2 b = *reinterpret_cast<const float2*>(bias + j);
    float p0 = valid ? __expf(v.x + b.x - l) : 0.f, p1 = valid ? __expf(v.y + b.y - l) : 0.f;
    if (j == t) p0 -= 1.f;
    if (j + 1 == t) p1 -= 1.f;
    *reinterpret_cast<__half2*>(out + j) = __floats2half2_rn(g * p0, g * p1);
  }
}

// dz = dt * gelu'(z), z given as the derivative tensor saved by the forward GELU epilogue (the MLM transform's
// GELU sits between a GEMM and a LayerNorm, so its backward is not a GEMM epilogue)
__global__ void __launch_bounds__(256)
dgelu_kernel(const __half* __restrict__ dt, const __half* __restrict__ z, __half* __restrict__ dz, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  const uint4 a = *reinterpret_cast<const uint4*>(dt + i);
  const uint4 b = *reinterpret_cast<const uint4*>(z + i);
  const __half2* ah = reinterpret_cast<const __half2*>(&a);
  const __half2* bh = reinterpret_cast<const __half2*>(&b);
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 x = __half22float2(ah[t]), y = __half22float2(bh[t]);
    oh[t] = __floats2half2_rn(x.x * y.x, x.y * y.y);
  }
  *reinterpret_cast<uint4*>(dz + i) = o;
}

// ------------------------------------------------------------------------------ pairwise NLL (K8)
__global__ void __launch_bounds__(128)
pair_nll_fwd_kernel(const float* __restrict__ q, const float* __restrict__ a, const float* __restrict__ b, int n,
                    int dim, float* __restrict__ loss, long long* __restrict__ accs, float* __restrict__ logits) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= n) return;
  const float* qi = q + static_cast<long long>(i) * dim;
  const float* ai = a + static_cast<long long>(i) * dim;
  const float* bi = b + static_cast<long long>(i) * dim;
  float lp = 0.f, ln = 0.f;
  for (int c = lane; c < dim; c += 32) {
    lp = fmaf(qi[c], ai[c], lp);
    ln = fmaf(qi[c], bi[c], ln);
  }
  lp = warp_sum(lp);
  ln = warp_sum(ln);
  if (lane == 0) {
    const float x = ln - lp;  // loss = log(1 + exp(x))
    loss[i] = fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
    accs[i] = ln > lp ? 1 : 0;  // argmax, first index on ties
    logits[2 * i] = lp;
    logits[2 * i + 1] = ln;
  }
}

__global__ void __launch_bounds__(128)
pair_nll_bwd_kernel(const float* __restrict__ q, const float* __restrict__ a, const float* __restrict__ b,
                    const float* __restrict__ logits, const float* __restrict__ dloss, int n, int dim,
                    float* __restrict__ dq, float* __restrict__ da, float* __restrict__ db) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= n) return;
  const float x = logits[2 * i + 1] - logits[2 * i];
  const float sg = dloss[i] / (1.f + expf(-x));  // dloss * sigmoid(l- - l+)
  const long long o = static_cast<long long>(i) * dim;
  for (int c = lane; c < dim; c += 32) {
    const float qv = q[o + c], av = a[o + c], bv = b[o + c];
    if (dq) dq[o + c] = sg * (bv - av);
    if (da) da[o + c] = -sg * qv;
    if (db) db[o + c] = sg * qv;
  }
}

// ------------------------------------------------------------------------------ group stats (K10)
__global__ void group_reduce_fwd_kernel(const float* __restrict__ loss, const long long* __restrict__ g, int n,
                                        int n_groups, float* __restrict__ sums, float* __restrict__ counts) {
  extern __shared__ float sh[];  // [2 * n_groups]
  for (int i = threadIdx.x; i < 2 * n_groups; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long long gi = g[i];
    if (gi >= 0 && gi < n_groups) {
      atomicAdd(&sh[gi], loss[i]);
      atomicAdd(&sh[n_groups + gi], 1.f);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_groups; i += blockDim.x) {
    sums[i] = sh[i];
    counts[i] = sh[n_groups + i];
  }
}

__global__ void group_reduce_bwd_kernel(const float* __restrict__ dsums, const long long* __restrict__ g, int n,
                                        int n_groups, float* __restrict__ dloss) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long gi = g[i];
  dloss[i] = (gi >= 0 && gi < n_groups) ? dsums[gi] : 0.f;
}

// ------------------------------------------------------------------------------ Gram (K12)
// gram[G,G] += X[:, slab] X[:, slab]^T.  Each block walks its column slab in chunks of GR_C columns
// staged in shared memory (coalesced 16-byte loads along the rows); thread (ti, tj) keeps a 4x4
// block of the 64x64 (padded) Gram in registers.
constexpr int GR_C = 64;    // columns per chunk
constexpr int GR_G = 64;    // max groups
__global__ void __launch_bounds__(256)
gram_kernel(const float* __restrict__ X, int G, long long P, long long ldx, float* __restrict__ gram,
            long long cols_per_block) {
  __shared__ float sx[GR_G][GR_C + 1];
  const int tid = threadIdx.x;
  const int ti = tid >> 4, tj = tid & 15;
  float acc[4][4] = {};
  const long long c_begin = blockIdx.x * cols_per_block;
  const long long c_end = min(P, c_begin + cols_per_block);
  for (long long c0 = c_begin; c0 < c_end; c0 += GR_C) {
    __syncthreads();
    for (int e = tid; e < GR_G * GR_C; e += 256) {
      const int gi = e / GR_C, cc = e % GR_C;
      const long long c = c0 + cc;
      sx[gi][cc] = (gi < G && c < c_end) ? X[gi * ldx + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int cc = 0; cc < GR_C; ++cc) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sx[ti * 4 + i][cc];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sx[tj * 4 + j][cc];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = ti * 4 + i, gj = tj * 4 + j;
      if (gi < G && gj < G) atomicAdd(&gram[gi * G + gj], acc[i][j]);
    }
}

}  // namespace cdr

using namespace cdr;

extern "C" {

int cdr_pair_nll_fwd(const float* q, const float* a, const float* b, int32_t n, int32_t dim, float* loss,
                     int64_t* accs, float* logits, void* stream) {
  CDR_REQUIRE(q && a && b && loss && accs && logits, "cdr_pair_nll_fwd: null pointer");
  CDR_REQUIRE(n >= 0 && dim > 0, "cdr_pair_nll_fwd: bad shape");
  if (n == 0) return CDR_OK;
  pair_nll_fwd_kernel<<<(n + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      q, a, b, n, dim, loss, reinterpret_cast<long long*>(accs), logits);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_pair_nll_bwd(const float* q, const float* a, const float* b, const float* logits, const float* dloss,
                     int32_t n, int32_t dim, float* dq, float* da, float* db, void* stream) {
  CDR_REQUIRE(q && a && b && logits && dloss, "cdr_pair_nll_bwd: null pointer");
  CDR_REQUIRE(n >= 0 && dim > 0, "cdr_pair_nll_bwd: bad shape");
  if (n == 0) return CDR_OK;
  pair_nll_bwd_kernel<<<(n + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(q, a, b, logits, dloss, n, dim, dq,
                                                                                 da, db);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

static int simmat_check(const cdr_simmat_args* a, const char* who) {
  CDR_REQUIRE(a != nullptr, "%s: null args", who);
  CDR_REQUIRE(a->q && a->k && a->scores && a->lse, "%s: null pointer", who);
  CDR_REQUIRE(a->n_rows > 0 && a->n_keys > 0 && a->dim > 0, "%s: empty problem", who);
  CDR_REQUIRE(a->mode == CDR_SIM_QP || a->mode == CDR_SIM_COCO, "%s: bad mode %d", who, a->mode);
  CDR_REQUIRE(a->row_offset >= 0 && a->row_offset + a->n_rows <= a->n_keys,
              "%s: rows [%d, %d) are not a subset of the %d keys", who, a->row_offset, a->row_offset + a->n_rows,
              a->n_keys);
  if (a->mode == CDR_SIM_COCO)
    CDR_REQUIRE(a->n_keys % 2 == 0, "%s: COCO pairing needs an even number of keys (got %d)", who, a->n_keys);
  return CDR_OK;
}

int cdr_simmat_ce_fwd(const cdr_simmat_args* a, void* stream) {
  if (int rc = simmat_check(a, "cdr_simmat_ce_fwd")) return rc;
  CDR_REQUIRE(a->loss != nullptr, "cdr_simmat_ce_fwd: null loss");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // S[i,j] = <q_i, k_j>
  if (int rc = sgemm(a->q, a->k, a->scores, a->n_rows, a->n_keys, a->dim, a->dim, 1, 1, a->dim, a->n_keys, 1.f, st))
    return rc;
  simmat_ce_row_kernel<<<a->n_rows, 256, 0, st>>>(a->scores, a->loss, a->lse, a->n_keys, a->mode, a->row_offset,
                                                  a->loss_scale);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_simmat_ce_bwd(const cdr_simmat_args* a, void* stream) {
  if (int rc = simmat_check(a, "cdr_simmat_ce_bwd")) return rc;
  CDR_REQUIRE(a->gmat && a->dloss, "cdr_simmat_ce_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  simmat_ce_grad_kernel<<<a->n_rows, 256, 0, st>>>(a->scores, a->lse, a->dloss, a->gmat, a->n_keys, a->mode,
                                                   a->row_offset, a->loss_scale);
  CDR_LAUNCH_CHECK();
  if (a->dq)  // dq[i,d] = sum_j G[i,j] k[j,d]
    if (int rc = sgemm(a->gmat, a->k, a->dq, a->n_rows, a->dim, a->n_keys, a->n_keys, 1, a->dim, 1, a->dim, 1.f, st))
      return rc;
  if (a->dk)  // dk[j,d] = sum_i G[i,j] q[i,d]
    if (int rc = sgemm(a->gmat, a->q, a->dk, a->n_keys, a->dim, a->n_rows, 1, a->n_keys, a->dim, 1, a->dim, 1.f, st))
      return rc;
  return CDR_OK;
}

int cdr_simmat_own_key_grad(const float* q, const float* loss, const float* dloss, int32_t n, int32_t dim, float* dk_own,
                            void* stream) {
  CDR_REQUIRE(q && loss && dloss && dk_own, "cdr_simmat_own_key_grad: null pointer");
  CDR_REQUIRE(n >= 0 && dim > 0, "cdr_simmat_own_key_grad: bad shape");
  if (n == 0) return CDR_OK;
  own_key_grad_kernel<<<(n + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(q, loss, dloss, n, dim, dk_own);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_vocab_ce_fwd(const float* logits, const float* bias, const int64_t* labels, float* loss, float* lse,
                     int32_t n_rows, int32_t n_cols, int64_t ld, void* stream) {
  CDR_REQUIRE(logits && bias && labels && loss && lse, "cdr_vocab_ce_fwd: null pointer");
  CDR_REQUIRE(n_rows >= 0 && n_cols > 0 && ld >= n_cols, "cdr_vocab_ce_fwd: bad shape");
  if (n_rows == 0) return CDR_OK;
  vocab_ce_fwd_kernel<<<n_rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, bias, reinterpret_cast<const long long*>(labels), loss, lse, n_cols, ld);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_vocab_ce_bwd(const float* logits, const float* bias, const int64_t* labels, const float* lse,
                     const float* dloss, void* dlogits, int32_t n_rows, int32_t n_cols, int64_t ld, float scale,
                     void* stream) {
  CDR_REQUIRE(logits && bias && labels && lse && dloss && dlogits, "cdr_vocab_ce_bwd: null pointer");
  CDR_REQUIRE(n_rows >= 0 && n_cols > 0 && n_cols % 2 == 0 && ld >= n_cols && ld % 2 == 0,
              "cdr_vocab_ce_bwd: bad shape (n_cols and ld must be even)");
  if (n_rows == 0) return CDR_OK;
  vocab_ce_bwd_kernel<<<n_rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, bias, reinterpret_cast<const long long*>(labels), lse, dloss, static_cast<__half*>(dlogits), n_cols, ld,
      scale);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_dgelu_f16(const void* dt, const void* z, void* dz, int64_t n, void* stream) {
  CDR_REQUIRE(dt && z && dz, "cdr_dgelu_f16: null pointer");
  CDR_REQUIRE(n >= 0 && n % 8 == 0, "cdr_dgelu_f16: n must be a multiple of 8");
  if (n == 0) return CDR_OK;
  dgelu_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(dt), static_cast<const __half*>(z), static_cast<__half*>(dz), n);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_group_reduce_fwd(const float* loss, const int64_t* g, int32_t n, int32_t n_groups, float* sums, float* counts,
                         void* stream) {
  CDR_REQUIRE(loss && g && sums && counts, "cdr_group_reduce_fwd: null pointer");
  CDR_REQUIRE(n >= 0 && n_groups > 0 && n_groups <= 4096, "cdr_group_reduce_fwd: bad shape n=%d groups=%d", n, n_groups);
  group_reduce_fwd_kernel<<<1, 256, 2 * n_groups * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      loss, reinterpret_cast<const long long*>(g), n, n_groups, sums, counts);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_group_reduce_bwd(const float* dsums, const int64_t* g, int32_t n, int32_t n_groups, float* dloss,
                         void* stream) {
  CDR_REQUIRE(dsums && g && dloss, "cdr_group_reduce_bwd: null pointer");
  if (n <= 0) return CDR_OK;
  group_reduce_bwd_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dsums, reinterpret_cast<const long long*>(g), n, n_groups, dloss);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_gram_f32(const float* x, int32_t g, int64_t p, int64_t ldx, float* gram, void* stream) {
  CDR_REQUIRE(x && gram, "cdr_gram_f32: null pointer");
  CDR_REQUIRE(g > 0 && g <= GR_G, "cdr_gram_f32: 1 <= groups <= %d (got %d)", GR_G, g);
  if (p <= 0) return CDR_OK;
  long long blocks = 4LL * sm_count();
  long long cpb = (p + blocks - 1) / blocks;
  cpb = ((cpb + GR_C - 1) / GR_C) * GR_C;
  blocks = (p + cpb - 1) / cpb;
  gram_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, g, p, ldx, gram, cpb);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

}  // extern "C"
