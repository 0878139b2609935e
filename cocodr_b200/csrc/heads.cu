// Contrastive heads and DRO statistics (K8, K9/K9', K10, K12).  fp32 CUDA-core kernels: these ops are
// latency-bound (<= 1 GFLOP, <= 2 MB) except the Gram matrix, which streams the [G, P_last] gradient
// matrix once from HBM.  Logits stay fp32 because trained CLS dot products are ~217 with gaps ~0.3.
// The similarity matrix of K9 / K9' is never materialised: score tiles live in registers / shared memory.
#include "cdr_common.cuh"

namespace cdr {

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

// ------------------------------------------------------------------------------ fused similarity matrix + CE (K9 / K9')
// S = q k^T is never written to memory (COCO/modeling.py:244-248 materialises it: matmul, fill_diagonal_, cross_entropy).
// All three kernels share one tile step: a block keeps SIM_TO "outer" vectors resident in shared memory ([16][dim] fp32)
// and streams "inner" vectors in tiles of SIM_TI = 64, SIM_DC = 128 dimensions at a time (the chunk loop is latency
// bound: global -> shared, barrier, FMAs, barrier; wide chunks mean few trips); thread (o, ic) of 256 holds the four
// products A[o][ic + 16 u] in registers.
//   forward  (rows outer, keys inner): online softmax over the key tiles of the block's key range -> per (row, split)
//            partial (max, sum) + the target logit; simmat_combine_kernel folds the splits into lse / loss
//   backward (rows outer -> dq, keys outer -> dk): the tile of S is recomputed, turned into
//            G = dloss * loss_scale * (exp(S - lse) - [key == target]) in shared memory, and the block accumulates
//            G (or G^T) times the SAME streamed inner vectors into its outer gradient rows (registers, 8 dims per
//            thread and chunk); splits of the inner range add into the zero-filled output with atomics
constexpr int SIM_TO = 16, SIM_TI = 64, SIM_DC = 128, SIM_THREADS = 256, SIM_DPT = SIM_DC / 16;
constexpr int SIM_MAX_DIM = 2048;

struct SimParams {
  const float* q;      // [n_rows, dim]
  const float* k;      // [n_keys, dim]
  const float* lse;    // [n_rows]   (backward)
  const float* dloss;  // [n_rows]   (backward)
  float* part;         // forward: [n_rows, n_splits, 2] partial (max, sum)
  float* tgt;          // forward: [n_rows] target logit
  float* out;          // backward: dq [n_rows, dim] or dk [n_keys, dim]
  int n_rows, n_keys, dim, mode, row_offset, n_splits, inner_per_split;
  float loss_scale;
};

// A[o][ic + 16 u] (u < 4) for the outer tile in sX and inner vectors Y[i0 .. i0 + 64)
__device__ __forceinline__ void sim_tile_products(const float* __restrict__ sX, float (*sY)[SIM_DC + 1],
                                                  const float* __restrict__ Y, int n_inner, int i0, int dim, int o,
                                                  int ic, float (&acc)[4]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = 0.f;
  for (int c0 = 0; c0 < dim; c0 += SIM_DC) {
    __syncthreads();  // the previous chunk (or the caller's use of sY) is finished
#pragma unroll 8
    for (int e = 0; e < SIM_TI * SIM_DC / SIM_THREADS; ++e) {
      const int idx = tid + e * SIM_THREADS;
      const int j = idx / SIM_DC, dd = idx % SIM_DC;
      sY[j][dd] = (i0 + j < n_inner && c0 + dd < dim) ? Y[static_cast<long long>(i0 + j) * dim + c0 + dd] : 0.f;
    }
    __syncthreads();
    const int nd = min(SIM_DC, dim - c0);
#pragma unroll 4
    for (int dd = 0; dd < nd; ++dd) {
      const float x = sX[o * dim + c0 + dd];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fmaf(x, sY[ic + 16 * u][dd], acc[u]);
    }
  }
}

__device__ __forceinline__ void sim_load_outer(float* sX, const float* __restrict__ X, int n_outer, int o0, int dim) {
  for (int e = threadIdx.x; e < SIM_TO * dim; e += SIM_THREADS) {
    const int o = e / dim, d = e - o * dim;
    sX[e] = (o0 + o < n_outer) ? X[static_cast<long long>(o0 + o) * dim + d] : 0.f;
  }
}

__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int s = 8; s > 0; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int s = 8; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

__global__ void __launch_bounds__(SIM_THREADS)
simmat_fused_fwd_kernel(const SimParams p) {
  extern __shared__ float sim_smem[];
  float* sX = sim_smem;                                                        // [16][dim]
  float (*sY)[SIM_DC + 1] = reinterpret_cast<float (*)[SIM_DC + 1]>(sim_smem + SIM_TO * p.dim);  // [64][33]
  const int o = threadIdx.x >> 4, ic = threadIdx.x & 15;
  const int r0 = blockIdx.x * SIM_TO, split = blockIdx.y;
  const int k_begin = split * p.inner_per_split, k_end = min(p.n_keys, k_begin + p.inner_per_split);
  sim_load_outer(sX, p.q, p.n_rows, r0, p.dim);
  const int row = r0 + o, gi = p.row_offset + row;
  const int t = sim_target(p.mode, gi);
  float m = -INFINITY, l = 0.f;
  for (int j0 = k_begin; j0 < k_end; j0 += SIM_TI) {
    float a[4];
    sim_tile_products(sX, sY, p.k, k_end, j0, p.dim, o, ic, a);
    float tmax = -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int key = j0 + ic + 16 * u;
      if (key >= k_end || (p.mode == CDR_SIM_COCO && key == gi)) a[u] = -INFINITY;
      if (key == t && key < k_end && row < p.n_rows) p.tgt[row] = a[u];
      tmax = fmaxf(tmax, a[u]);
    }
    tmax = half_warp_max(tmax);
    const float m_new = fmaxf(m, tmax);
    float s = 0.f;
    if (m_new > -INFINITY) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s += expf(a[u] - m_new);
      l = l * expf(m - m_new);
    }
    l += half_warp_sum(s);
    m = m_new;
  }
  if (ic == 0 && row < p.n_rows) {
    float* dst = p.part + (static_cast<long long>(row) * p.n_splits + split) * 2;
    dst[0] = m;
    dst[1] = l;
  }
}

__global__ void simmat_combine_kernel(const float* __restrict__ part, const float* __restrict__ tgt, int n_rows,
                                      int n_splits, float loss_scale, float* __restrict__ loss, float* __restrict__ lse) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const float* pr = part + static_cast<long long>(i) * n_splits * 2;
  float m = -INFINITY;
  for (int s = 0; s < n_splits; ++s) m = fmaxf(m, pr[2 * s]);
  float l = 0.f;
  for (int s = 0; s < n_splits; ++s)
    if (pr[2 * s] > -INFINITY) l += pr[2 * s + 1] * expf(pr[2 * s] - m);
  const float v = m + logf(l);
  lse[i] = v;
  loss[i] = loss_scale * (v - tgt[i]);
}

// KEYS_OUTER = false: outer = query rows (out = dq), inner = keys.  true: outer = keys (out = dk), inner = rows.
template <bool KEYS_OUTER>
__global__ void __launch_bounds__(SIM_THREADS)
simmat_fused_bwd_kernel(const SimParams p) {
  extern __shared__ float sim_smem[];
  float* sX = sim_smem;                                                                        // [16][dim]
  float (*sY)[SIM_DC + 1] = reinterpret_cast<float (*)[SIM_DC + 1]>(sim_smem + SIM_TO * p.dim);  // [64][33]
  float (*sG)[SIM_TI + 1] = reinterpret_cast<float (*)[SIM_TI + 1]>(sim_smem + SIM_TO * p.dim + SIM_TI * (SIM_DC + 1));
  const int o = threadIdx.x >> 4, ic = threadIdx.x & 15;
  const int o0 = blockIdx.x * SIM_TO, split = blockIdx.y;
  const float* X = KEYS_OUTER ? p.k : p.q;
  const float* Y = KEYS_OUTER ? p.q : p.k;
  const int n_outer = KEYS_OUTER ? p.n_keys : p.n_rows;
  const int n_inner = KEYS_OUTER ? p.n_rows : p.n_keys;
  const int i_begin = split * p.inner_per_split, i_end = min(n_inner, i_begin + p.inner_per_split);
  sim_load_outer(sX, X, n_outer, o0, p.dim);
  const int n_chunks = (p.dim + SIM_DC - 1) / SIM_DC;
  float acc2[SIM_MAX_DIM / SIM_DC][SIM_DPT];  // out[o][c * 128 + ic + 16 t]
#pragma unroll
  for (int c = 0; c < SIM_MAX_DIM / SIM_DC; ++c)
#pragma unroll
    for (int t = 0; t < SIM_DPT; ++t) acc2[c][t] = 0.f;
  for (int i0 = i_begin; i0 < i_end; i0 += SIM_TI) {
    float a[4];
    sim_tile_products(sX, sY, Y, i_end, i0, p.dim, o, ic, a);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int inner = i0 + ic + 16 * u;
      const int row = KEYS_OUTER ? inner : o0 + o;
      const int key = KEYS_OUTER ? o0 + o : inner;
      float gval = 0.f;
      if (row < p.n_rows && key < p.n_keys && inner < i_end) {
        const int gi = p.row_offset + row;
        const float sc = (p.mode == CDR_SIM_COCO && key == gi) ? -INFINITY : a[u];
        gval = __ldg(p.dloss + row) * p.loss_scale * (expf(sc - __ldg(p.lse + row)) - (key == sim_target(p.mode, gi) ? 1.f : 0.f));
      }
      sG[o][ic + 16 * u] = gval;
    }
    // out[o][d] += sum_i G[o][i] * Y[i][d], the inner tile streamed once more in 32-dim chunks
    for (int c = 0; c < n_chunks; ++c) {
      __syncthreads();  // sG complete (first chunk) / previous chunk consumed
#pragma unroll 8
      for (int e = 0; e < SIM_TI * SIM_DC / SIM_THREADS; ++e) {
        const int idx = threadIdx.x + e * SIM_THREADS;
        const int j = idx / SIM_DC, dd = idx % SIM_DC;
        sY[j][dd] = (i0 + j < i_end && c * SIM_DC + dd < p.dim) ? Y[static_cast<long long>(i0 + j) * p.dim + c * SIM_DC + dd] : 0.f;
      }
      __syncthreads();
      float sacc[SIM_DPT];
#pragma unroll
      for (int t = 0; t < SIM_DPT; ++t) sacc[t] = 0.f;
#pragma unroll 4
      for (int j = 0; j < SIM_TI; ++j) {
        const float gv = sG[o][j];
#pragma unroll
        for (int t = 0; t < SIM_DPT; ++t) sacc[t] = fmaf(gv, sY[j][ic + 16 * t], sacc[t]);
      }
      // (acc2 is indexed by the chunk loop: keep the loop bounded by a compile-time maximum so it stays in registers)
#pragma unroll
      for (int cc = 0; cc < SIM_MAX_DIM / SIM_DC; ++cc)
        if (cc == c) {
#pragma unroll
          for (int t = 0; t < SIM_DPT; ++t) acc2[cc][t] += sacc[t];
        }
    }
  }
  if (o0 + o < n_outer) {
    float* dst = p.out + static_cast<long long>(o0 + o) * p.dim;
#pragma unroll
    for (int c = 0; c < SIM_MAX_DIM / SIM_DC; ++c)
#pragma unroll
      for (int t = 0; t < SIM_DPT; ++t) {
        const int d = c * SIM_DC + ic + 16 * t;
        if (c < n_chunks && d < p.dim) {
          if (p.n_splits > 1) atomicAdd(dst + d, acc2[c][t]);
          else dst[d] = acc2[c][t];
        }
      }
  }
}

// ---- small problems (the in-batch heads of the training step: 64 x 64 .. 512 x 512 scores): ONE BLOCK PER VECTOR.
// The tiled kernels above give a 64-row problem 4 blocks that each walk their chunks serially (105 us per launch, 0.19 ms
// of an 11 ms step); here every query row (forward, dq) or key (dk) gets its own block: warp w forms the dot products
// with inner vectors w, w + 8, ... (lanes stride the dimension: coalesced 128-byte reads), the block turns them into
// the loss or into G, and thread d accumulates out[d] = sum_i G_i Y[i][d] over the same inner vectors.  fp32 throughout.
constexpr int SIMS_THREADS = 256;
constexpr int SIMS_MAX_INNER = 2048;

// s[i] = <x, Y_i> for i < n_inner (x in shared memory); called by the whole block
__device__ __forceinline__ void sims_dots(const float* __restrict__ sx, const float* __restrict__ Y, int n_inner, int dim,
                                          float* __restrict__ s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < n_inner; i += SIMS_THREADS / 32) {
    const float* y = Y + static_cast<long long>(i) * dim;
    float acc = 0.f;
    for (int d = lane; d < dim; d += 32) acc = fmaf(sx[d], __ldg(y + d), acc);
    acc = warp_sum(acc);
    if (lane == 0) s[i] = acc;
  }
}

__device__ __forceinline__ float sims_block_reduce(float v, float* red, bool is_max) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = red[0];
#pragma unroll
  for (int w = 1; w < SIMS_THREADS / 32; ++w) t = is_max ? fmaxf(t, red[w]) : t + red[w];
  return t;
}

__global__ void __launch_bounds__(SIMS_THREADS)
simmat_small_fwd_kernel(const SimParams p, float* __restrict__ loss, float* __restrict__ lse) {
  extern __shared__ float sims_smem[];
  float* sx = sims_smem;          // [dim]
  float* s = sims_smem + p.dim;   // [n_keys]
  __shared__ float red[SIMS_THREADS / 32];
  const int row = blockIdx.x, gi = p.row_offset + row;
  for (int d = threadIdx.x; d < p.dim; d += SIMS_THREADS) sx[d] = p.q[static_cast<long long>(row) * p.dim + d];
  __syncthreads();
  sims_dots(sx, p.k, p.n_keys, p.dim, s);
  __syncthreads();
  float m = -INFINITY;
  for (int j = threadIdx.x; j < p.n_keys; j += SIMS_THREADS)
    if (!(p.mode == CDR_SIM_COCO && j == gi)) m = fmaxf(m, s[j]);
  m = sims_block_reduce(m, red, true);
  float l = 0.f;
  for (int j = threadIdx.x; j < p.n_keys; j += SIMS_THREADS)
    if (!(p.mode == CDR_SIM_COCO && j == gi)) l += expf(s[j] - m);
  l = sims_block_reduce(l, red, false);
  if (threadIdx.x == 0) {
    const float v = m + logf(l);
    lse[row] = v;
    loss[row] = p.loss_scale * (v - s[sim_target(p.mode, gi)]);
  }
}

// KEYS_OUTER = false: block = query row, out = dq[row];  true: block = key, out = dk[key]
template <bool KEYS_OUTER>
__global__ void __launch_bounds__(SIMS_THREADS)
simmat_small_bwd_kernel(const SimParams p) {
  extern __shared__ float sims_smem[];
  float* sx = sims_smem;          // [dim]
  float* s = sims_smem + p.dim;   // [n_inner]: scores, then G
  const int o = blockIdx.x;
  const float* X = KEYS_OUTER ? p.k : p.q;
  const float* Y = KEYS_OUTER ? p.q : p.k;
  const int n_inner = KEYS_OUTER ? p.n_rows : p.n_keys;
  for (int d = threadIdx.x; d < p.dim; d += SIMS_THREADS) sx[d] = X[static_cast<long long>(o) * p.dim + d];
  __syncthreads();
  sims_dots(sx, Y, n_inner, p.dim, s);
  __syncthreads();
  for (int i = threadIdx.x; i < n_inner; i += SIMS_THREADS) {
    const int row = KEYS_OUTER ? i : o, key = KEYS_OUTER ? o : i;
    const int gi = p.row_offset + row;
    const float sc = (p.mode == CDR_SIM_COCO && key == gi) ? -INFINITY : s[i];
    s[i] = __ldg(p.dloss + row) * p.loss_scale * (expf(sc - __ldg(p.lse + row)) - (key == sim_target(p.mode, gi) ? 1.f : 0.f));
  }
  __syncthreads();
  for (int d = threadIdx.x; d < p.dim; d += SIMS_THREADS) {
    float a0 = 0.f, a1 = 0.f;
    int i = 0;
    for (; i + 1 < n_inner; i += 2) {
      a0 = fmaf(s[i], __ldg(Y + static_cast<long long>(i) * p.dim + d), a0);
      a1 = fmaf(s[i + 1], __ldg(Y + static_cast<long long>(i + 1) * p.dim + d), a1);
    }
    if (i < n_inner) a0 = fmaf(s[i], __ldg(Y + static_cast<long long>(i) * p.dim + d), a0);
    p.out[static_cast<long long>(o) * p.dim + d] = a0 + a1;
  }
}

static bool sim_small(const cdr_simmat_args* a) {
  return a->n_rows <= SIMS_MAX_INNER && a->n_keys <= SIMS_MAX_INNER &&
         static_cast<long long>(a->n_rows) * a->n_keys <= 1024 * 1024;
}

static int sim_splits(int n_outer, int n_inner) {
  const int tiles = (n_outer + SIM_TO - 1) / SIM_TO;
  const int max_splits = (n_inner + SIM_TI - 1) / SIM_TI;
  int s = (2 * sm_count() + tiles - 1) / tiles;
  if (s > max_splits) s = max_splits;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
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

int gram_tf32_launch(const float* x, int g, long long p, long long ldx, float* gram, cudaStream_t st);  // gram.cu

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
  CDR_REQUIRE(a->q && a->k && a->lse, "%s: null pointer", who);
  CDR_REQUIRE(a->n_rows > 0 && a->n_keys > 0 && a->dim > 0, "%s: empty problem", who);
  CDR_REQUIRE(a->dim <= SIM_MAX_DIM, "%s: dim %d > %d", who, a->dim, SIM_MAX_DIM);
  CDR_REQUIRE(a->mode == CDR_SIM_QP || a->mode == CDR_SIM_COCO, "%s: bad mode %d", who, a->mode);
  CDR_REQUIRE(a->row_offset >= 0 && a->row_offset + a->n_rows <= a->n_keys,
              "%s: rows [%d, %d) are not a subset of the %d keys", who, a->row_offset, a->row_offset + a->n_rows,
              a->n_keys);
  if (a->mode == CDR_SIM_COCO)
    CDR_REQUIRE(a->n_keys % 2 == 0, "%s: COCO pairing needs an even number of keys (got %d)", who, a->n_keys);
  return CDR_OK;
}

size_t cdr_simmat_workspace_bytes(int32_t n_rows, int32_t n_keys) {
  if (n_rows <= 0 || n_keys <= 0) return 0;
  return static_cast<size_t>(n_rows) * (2 * static_cast<size_t>(sim_splits(n_rows, n_keys)) + 1) * sizeof(float);
}

static size_t sim_smem_bytes(int dim, bool bwd) {
  return sizeof(float) * (static_cast<size_t>(SIM_TO) * dim + SIM_TI * (SIM_DC + 1) + (bwd ? SIM_TO * (SIM_TI + 1) : 0));
}

static void sim_params(const cdr_simmat_args* a, SimParams& p) {
  p.q = a->q; p.k = a->k; p.lse = a->lse; p.dloss = a->dloss;
  p.n_rows = a->n_rows; p.n_keys = a->n_keys; p.dim = a->dim; p.mode = a->mode; p.row_offset = a->row_offset;
  p.loss_scale = a->loss_scale;
}

int cdr_simmat_ce_fwd(const cdr_simmat_args* a, void* stream) {
  if (int rc = simmat_check(a, "cdr_simmat_ce_fwd")) return rc;
  CDR_REQUIRE(a->loss != nullptr && a->scores != nullptr, "cdr_simmat_ce_fwd: null loss / workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SimParams p{};
  sim_params(a, p);
  if (sim_small(a)) {
    simmat_small_fwd_kernel<<<a->n_rows, SIMS_THREADS, sizeof(float) * (a->dim + a->n_keys), st>>>(p, a->loss, a->lse);
    CDR_LAUNCH_CHECK();
    return CDR_OK;
  }
  p.n_splits = sim_splits(a->n_rows, a->n_keys);
  p.inner_per_split = ((a->n_keys + p.n_splits - 1) / p.n_splits + SIM_TI - 1) / SIM_TI * SIM_TI;
  p.n_splits = (a->n_keys + p.inner_per_split - 1) / p.inner_per_split;
  p.part = a->scores;  // workspace: [n_rows, n_splits, 2] partials | [n_rows] target logits
  p.tgt = a->scores + static_cast<size_t>(a->n_rows) * p.n_splits * 2;
  const size_t smem = sim_smem_bytes(a->dim, false);
  static bool cfg = false;
  if (!cfg) {
    CDR_CUDA(cudaFuncSetAttribute(simmat_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CDR_CUDA(cudaFuncSetAttribute(simmat_fused_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CDR_CUDA(cudaFuncSetAttribute(simmat_fused_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cfg = true;
  }
  simmat_fused_fwd_kernel<<<dim3((a->n_rows + SIM_TO - 1) / SIM_TO, p.n_splits), SIM_THREADS, smem, st>>>(p);
  CDR_LAUNCH_CHECK();
  simmat_combine_kernel<<<(a->n_rows + 127) / 128, 128, 0, st>>>(p.part, p.tgt, a->n_rows, p.n_splits, a->loss_scale,
                                                                a->loss, a->lse);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_simmat_ce_bwd(const cdr_simmat_args* a, void* stream) {
  if (int rc = simmat_check(a, "cdr_simmat_ce_bwd")) return rc;
  CDR_REQUIRE(a->dloss != nullptr, "cdr_simmat_ce_bwd: null dloss");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = sim_smem_bytes(a->dim, true);
  static bool cfg = false;
  if (!cfg) {
    CDR_CUDA(cudaFuncSetAttribute(simmat_fused_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CDR_CUDA(cudaFuncSetAttribute(simmat_fused_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cfg = true;
  }
  for (int pass = 0; pass < 2; ++pass) {
    float* out = pass == 0 ? a->dq : a->dk;
    if (out == nullptr) continue;
    const int n_outer = pass == 0 ? a->n_rows : a->n_keys, n_inner = pass == 0 ? a->n_keys : a->n_rows;
    SimParams p{};
    sim_params(a, p);
    p.out = out;
    if (sim_small(a)) {
      const size_t sm = sizeof(float) * (a->dim + n_inner);
      if (pass == 0) simmat_small_bwd_kernel<false><<<n_outer, SIMS_THREADS, sm, st>>>(p);
      else simmat_small_bwd_kernel<true><<<n_outer, SIMS_THREADS, sm, st>>>(p);
      CDR_LAUNCH_CHECK();
      continue;
    }
    p.n_splits = sim_splits(n_outer, n_inner);
    p.inner_per_split = ((n_inner + p.n_splits - 1) / p.n_splits + SIM_TI - 1) / SIM_TI * SIM_TI;
    p.n_splits = (n_inner + p.inner_per_split - 1) / p.inner_per_split;
    if (p.n_splits > 1) CDR_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * static_cast<size_t>(n_outer) * a->dim, st));
    const dim3 grid((n_outer + SIM_TO - 1) / SIM_TO, p.n_splits);
    if (pass == 0) simmat_fused_bwd_kernel<false><<<grid, SIM_THREADS, smem, st>>>(p);
    else simmat_fused_bwd_kernel<true><<<grid, SIM_THREADS, smem, st>>>(p);
    CDR_LAUNCH_CHECK();
  }
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
  {  // tensor-core path (gram.cu: tcgen05 kind::tf32, HBM-bound); layouts it cannot take fall through to the fp32 kernel
    const int rc = gram_tf32_launch(x, g, p, ldx, gram, static_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;
  }
  long long blocks = 4LL * sm_count();
  long long cpb = (p + blocks - 1) / blocks;
  cpb = ((cpb + GR_C - 1) / GR_C) * GR_C;
  blocks = (p + cpb - 1) / cpb;
  gram_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, g, p, ldx, gram, cpb);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

}  // extern "C"
