// Memory-bound kernels of the encoder: embedding gather + LayerNorm (K1), LayerNorm fwd/bwd over
// fp16 activations (the LN halves of K4/K6), column sums for bias gradients, fp32->fp16 casts.
// fp32 statistics, 16-byte vector accesses, hidden % 8 == 0, hidden <= 2048.  LayerNorm over hidden <= 1024 runs as
// staged kernels (ln_fwd_staged_kernel / ln_bwd_staged_kernel: 8-row tiles brought in by cp.async.bulk into a
// 3-stage mbarrier ring, a warp per row for the statistics, a thread per 8 columns for the column sums); wider rows
// and the fp32 [CLS]-gradient variant keep the one-warp-per-row kernels.
#include <stdlib.h>

#include "cdr_common.cuh"
#include "dropout.cuh"
#include "peer.cuh"

namespace cdr {

constexpr int LN_WARPS = 4;
constexpr int MAX_VPL = 8;  // vectors (of 8 elements) per lane -> hidden <= 2048

__device__ __forceinline__ void load8_h(const __half* p, float (&v)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = __half22float2(h[t]);
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
}
__device__ __forceinline__ void store8_h(__half* p, const float (&v)[8]) {
  uint4 q;
  __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
  for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
  *reinterpret_cast<uint4*>(p) = q;
}
__device__ __forceinline__ void load8_f(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8_f(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void atomic_add8(float* p, const float (&v)[8]) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  atomicAdd(reinterpret_cast<float4*>(p + 4), make_float4(v[4], v[5], v[6], v[7]));
}

// Row statistics from register-resident values (two-pass: mean, then centred variance).
template <int VPL>
__device__ __forceinline__ void row_stats(const float (&x)[VPL][8], int nvec, int lane, int hidden, float eps,
                                          float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    if (lane + 32 * i < nvec)
#pragma unroll
      for (int t = 0; t < 8; ++t) s += x[i][t];
  mean = warp_sum(s) / hidden;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    if (lane + 32 * i < nvec)
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float d = x[i][t] - mean;
        q += d * d;
      }
  rstd = rsqrtf(warp_sum(q) / hidden + eps);
}

// ------------------------------------------------------------------------------------------ fwd
// MODE 0: x = fp16 activations.  MODE 1: x = word[id] + pos[l] + type0 (fp32 tables).
template <int VPL, int MODE>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const __half* __restrict__ x, const int64_t* __restrict__ ids, const float* __restrict__ word,
              const float* __restrict__ pos, const float* __restrict__ type0, const float* __restrict__ gamma,
              const float* __restrict__ beta, __half* __restrict__ y, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, float* __restrict__ cls_out, int rows, int hidden, int seq_len,
              int vocab, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = hidden >> 3;
  float v[VPL][8];
  if constexpr (MODE == 0) {
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (lane + 32 * i < nvec) load8_h(x + static_cast<long long>(row) * hidden + 8 * (lane + 32 * i), v[i]);
  } else {
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const int l = row % seq_len;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (lane + 32 * i < nvec) {
        const int c = 8 * (lane + 32 * i);
        float a[8], b[8], t[8];
        load8_f(word + id * hidden + c, a);
        load8_f(pos + static_cast<long long>(l) * hidden + c, b);
        load8_f(type0 + c, t);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[i][k] = (a[k] + t[k]) + b[k];  // HF order: (word + type) + pos
      }
  }
  float mean, rstd;
  row_stats<VPL>(v, nvec, lane, hidden, eps, mean, rstd);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const bool is_cls = cls_out != nullptr && (row % seq_len) == 0;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    if (lane + 32 * i < nvec) {
      const int c = 8 * (lane + 32 * i);
      float g[8], b[8], o[8];
      load8_f(gamma + c, g);
      load8_f(beta + c, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (v[i][k] - mean) * rstd * g[k] + b[k];
      store8_h(y + static_cast<long long>(row) * hidden + c, o);
      if (is_cls) store8_f(cls_out + static_cast<long long>(row / seq_len) * hidden + c, o);
    }
}

// ------------------------------------------------------------------------------------------ bwd
// Block (l, split): position l, sequences split, split+gridDim.y, ...  (rows = seq*seq_len + l), so
// that the position-embedding gradient of MODE 1 reduces inside the block.  For MODE 0 the same
// mapping is used (any row partition works for the column sums).
// Column sums (dgamma, dbeta, dcol = sum of dx) are reduced warp-registers -> smem -> atomics.
template <int VPL, int MODE>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(const __half* __restrict__ dy, const __half* __restrict__ x, const int64_t* __restrict__ ids,
              const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type0,
              const float* __restrict__ gamma, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const float* __restrict__ dy_cls, __half* __restrict__ dx, float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ dcol, float* __restrict__ dword, float* __restrict__ dpos,
              int n_seq, int hidden, int seq_len, int vocab, int pad_id, float in_scale, float out_scale) {
  extern __shared__ float red[];  // [3][hidden]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int l = blockIdx.x;
  const int nvec = hidden >> 3;
  float a_dg[VPL][8], a_db[VPL][8], a_dc[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) a_dg[i][k] = a_db[i][k] = a_dc[i][k] = 0.f;
  float g[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    if (lane + 32 * i < nvec) load8_f(gamma + 8 * (lane + 32 * i), g[i]);

  for (int seq = blockIdx.y * LN_WARPS + warp; seq < n_seq; seq += gridDim.y * LN_WARPS) {
    const long long row = static_cast<long long>(seq) * seq_len + l;
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[VPL][8], gy[VPL][8];
    long long id = 0;
    if constexpr (MODE == 1) {
      id = ids[row];
      id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (lane + 32 * i < nvec) {
        const int c = 8 * (lane + 32 * i);
        float xv[8], d[8];
        if constexpr (MODE == 0) {
          load8_h(x + row * hidden + c, xv);
        } else {
          float a[8], b[8], t[8];
          load8_f(word + id * hidden + c, a);
          load8_f(pos + static_cast<long long>(l) * hidden + c, b);
          load8_f(type0 + c, t);
#pragma unroll
          for (int k = 0; k < 8; ++k) xv[k] = (a[k] + t[k]) + b[k];
        }
        if (dy != nullptr) {
          load8_h(dy + row * hidden + c, d);
          if constexpr (MODE == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) d[k] *= in_scale;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) d[k] = 0.f;
        }
        if (dy_cls != nullptr && l == 0) {  // fp32 gradient of the CLS embedding output
          float e[8];
          load8_f(dy_cls + static_cast<long long>(seq) * hidden + c, e);
#pragma unroll
          for (int k = 0; k < 8; ++k) d[k] += e[k] * in_scale;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          xh[i][k] = (xv[k] - mean) * rstd;
          gy[i][k] = d[k] * g[i][k];
          s1 += gy[i][k];
          s2 += gy[i][k] * xh[i][k];
          a_dg[i][k] += d[k] * xh[i][k];
          a_db[i][k] += d[k];
        }
      }
    const float c1 = warp_sum(s1) / hidden;
    const float c2 = warp_sum(s2) / hidden;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (lane + 32 * i < nvec) {
        const int c = 8 * (lane + 32 * i);
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          o[k] = rstd * (gy[i][k] - c1 - xh[i][k] * c2);
          a_dc[i][k] += o[k];
        }
        if constexpr (MODE == 0) {
          store8_h(dx + row * hidden + c, o);
        } else {
          if (id != pad_id) {
            float w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) w[k] = o[k] * out_scale;
            atomic_add8(dword + id * hidden + c, w);
          }
        }
      }
  }
  // block reduction of the three column accumulators, one at a time through smem
  auto reduce_to = [&](float (&acc)[VPL][8], float* dst, long long dst_off) {
    if (dst == nullptr) return;
    __syncthreads();
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (lane + 32 * i < nvec)
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&red[8 * (lane + 32 * i) + k], acc[i][k]);
    __syncthreads();
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) atomicAdd(dst + dst_off + i, red[i] * out_scale);
  };
  reduce_to(a_dg, dgamma, 0);
  reduce_to(a_db, dbeta, 0);
  if constexpr (MODE == 0) {
    reduce_to(a_dc, dcol, 0);
  } else {
    reduce_to(a_dc, dcol, 0);                                       // token_type_embeddings row 0
    reduce_to(a_dc, dpos, static_cast<long long>(l) * hidden);      // position row l
  }
}

// ------------------------------------------------------------------------------------------ bwd, split
// The fused kernel above keeps 3 x H/32 column accumulators per lane, which caps occupancy and
// serialises rows.  For the common case (fp16 dy only) the work is split in two bandwidth-shaped passes:
//   ln_bwd_dx_kernel   : warp per row, dx only (+ the two per-row means c1, c2)          read 4 B, write 2 B / elem
//   ln_bwd_cols_kernel : thread per 8 columns over a slab of rows -> dgamma, dbeta, dcol  read 4 B / elem
// using  sum_r dx[r,c] = g[c] * sum_r rstd_r dy[r,c] - sum_r rstd_r c1_r - sum_r rstd_r c2_r xhat[r,c].
template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_dx_kernel(const __half* __restrict__ dy, const __half* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ mean_in, const float* __restrict__ rstd_in, __half* __restrict__ dx,
                 float2* __restrict__ rowc, int rows, int hidden) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = hidden >> 3;
  const float mean = mean_in[row], rstd = rstd_in[row];
  float xh[VPL][8], gy[VPL][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    if (lane + 32 * i < nvec) {
      const int c = 8 * (lane + 32 * i);
      float xv[8], d[8], g[8];
      load8_h(x + static_cast<long long>(row) * hidden + c, xv);
      load8_h(dy + static_cast<long long>(row) * hidden + c, d);
      load8_f(gamma + c, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        xh[i][k] = (xv[k] - mean) * rstd;
        gy[i][k] = d[k] * g[k];
        s1 += gy[i][k];
        s2 += gy[i][k] * xh[i][k];
      }
    }
  const float c1 = warp_sum(s1) / hidden;
  const float c2 = warp_sum(s2) / hidden;
  if (lane == 0) rowc[row] = make_float2(c1, c2);
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    if (lane + 32 * i < nvec) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = rstd * (gy[i][k] - c1 - xh[i][k] * c2);
      store8_h(dx + static_cast<long long>(row) * hidden + 8 * (lane + 32 * i), o);
    }
}

constexpr int LNC_Y = 8;  // row lanes per block: cross-lane reduction in smem before the global atomics
__global__ void __launch_bounds__(128 * LNC_Y)
ln_bwd_cols_kernel(const __half* __restrict__ dy, const __half* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                   const float2* __restrict__ rowc, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   float* __restrict__ dcol, int rows, int hidden, int rows_per_block, float out_scale) {
  __shared__ float red[LNC_Y][128 * 8];
  __shared__ float red_c[LNC_Y];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = (blockIdx.x * blockDim.x + tx) * 8;
  const bool active = c < hidden;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, rows);
  float dg[8], db[8], s3[8], s4[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) dg[k] = db[k] = s3[k] = s4[k] = 0.f;
  float csum = 0.f;
  if (active) {
#pragma unroll 4
    for (int r = r0 + ty; r < r1; r += LNC_Y) {
      float d[8], xv[8];
      load8_h(dy + static_cast<long long>(r) * hidden + c, d);
      load8_h(x + static_cast<long long>(r) * hidden + c, xv);
      const float m = mean_in[r], rs = rstd_in[r];
      const float2 cc = rowc[r];
      const float rc2 = rs * cc.y;
      csum = fmaf(rs, cc.x, csum);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (xv[k] - m) * rs;
        dg[k] = fmaf(d[k], xh, dg[k]);
        db[k] += d[k];
        s3[k] = fmaf(rs, d[k], s3[k]);
        s4[k] = fmaf(rc2, xh, s4[k]);
      }
    }
  }
  if (tx == 0) red_c[ty] = csum;
  float g[8];
  if (active) load8_f(gamma + c, g);
  // dcol needs the block-wide sum of csum first
  __syncthreads();
  float ctot = 0.f;
#pragma unroll
  for (int y = 0; y < LNC_Y; ++y) ctot += red_c[y];
  auto reduce_emit = [&](float (&acc)[8], float* dst, bool is_dcol) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) red[ty][tx * 8 + k] = acc[k];
    __syncthreads();
    if (ty == 0 && active && dst != nullptr) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float t = 0.f;
#pragma unroll
        for (int y = 0; y < LNC_Y; ++y) t += red[y][tx * 8 + k];
        o[k] = (is_dcol ? (t - ctot) : t) * out_scale;
      }
      atomic_add8(dst + c, o);
    }
  };
#pragma unroll
  for (int k = 0; k < 8; ++k) s3[k] = active ? g[k] * s3[k] - s4[k] : 0.f;  // dcol before the -C term
  reduce_emit(dg, dgamma, false);
  reduce_emit(db, dbeta, false);
  reduce_emit(s3, dcol, true);
}

// ------------------------------------------------------------------------------------------ staged, persistent
// Bandwidth-shaped LayerNorm kernels.  A persistent block streams tiles of LNS_ROWS whole rows through a ring of
// shared-memory stages filled by 1-D bulk copies (cp.async.bulk + mbarrier: the rows of a tile are contiguous in
// global memory, so one copy per tensor per tile), which keeps ~50-70 KB per block in flight independent of
// occupancy.  Each row is read from HBM exactly once:
//   forward : warp per row -> statistics, normalise, store.
//   backward: phase A, warp per row -> dx and the per-row scalars; phase B, thread per 8 columns over the rows of
//             the SAME staged tile -> dgamma / dbeta / bias-gradient partial sums kept in registers across all
//             tiles of the block, one round of atomics per block at the end.
struct LnPush {          // optional fused all-gather of the CLS rows (cdr_ln_fwd_push); world == 0: off
  cdr_peer_args pa;
  int first_seq, n_push;
};

constexpr int LNS_ROWS = 8;      // rows per tile == warps per block
constexpr int LNS_THREADS = 256;
constexpr int LNS_STAGES = 3;

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int VPL>
__global__ void __launch_bounds__(LNS_THREADS, 2)
ln_fwd_staged_kernel(const __half* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __half* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     float* __restrict__ cls_out, int rows, int hidden, int seq_len, float eps, const LnPush push) {
  extern __shared__ __align__(128) uint8_t lns_smem[];
  __shared__ uint64_t full[LNS_STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nvec = hidden >> 3;
  const uint32_t row_bytes = static_cast<uint32_t>(hidden) * 2u;
  const uint32_t tile_bytes = row_bytes * LNS_ROWS;
  const int n_tiles = (rows + LNS_ROWS - 1) / LNS_ROWS;
  if (tid == 0) {
    for (int i = 0; i < LNS_STAGES; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_launch();
  auto issue = [&](int tile, int stage) {
    const int r0 = tile * LNS_ROWS;
    const uint32_t bytes = static_cast<uint32_t>(min(LNS_ROWS, rows - r0)) * row_bytes;
    mbar_expect_tx(&full[stage], bytes);
    bulk_load(lns_smem + stage * tile_bytes, x + static_cast<long long>(r0) * hidden, bytes, &full[stage]);
  };
  if (tid == 0)
    for (int i = 0; i < LNS_STAGES; ++i) {
      const int t = blockIdx.x + i * gridDim.x;
      if (t < n_tiles) issue(t, i);
    }
  float g[VPL][8], b[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    if (lane + 32 * i < nvec) {
      load8_f(gamma + 8 * (lane + 32 * i), g[i]);
      load8_f(beta + 8 * (lane + 32 * i), b[i]);
    }
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int stage = it % LNS_STAGES;
    mbar_wait(&full[stage], (it / LNS_STAGES) & 1);
    const int row = tile * LNS_ROWS + warp;
    if (row < rows) {
      const __half* xs = reinterpret_cast<const __half*>(lns_smem + stage * tile_bytes) + warp * hidden;
      float v[VPL][8];
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (lane + 32 * i < nvec) load8_h(xs + 8 * (lane + 32 * i), v[i]);
      float mean, rstd;
      row_stats<VPL>(v, nvec, lane, hidden, eps, mean, rstd);
      if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
      }
      const bool is_cls = cls_out != nullptr && (row % seq_len) == 0;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (lane + 32 * i < nvec) {
          const int c = 8 * (lane + 32 * i);
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = (v[i][k] - mean) * rstd * g[i][k] + b[i][k];
          store8_h(y + static_cast<long long>(row) * hidden + c, o);
          if (is_cls) store8_f(cls_out + static_cast<long long>(row / seq_len) * hidden + c, o);
          if (push.pa.world > 0 && (row % seq_len) == 0 && row / seq_len >= push.first_seq) {
            // fused all-gather: the CLS row goes straight into slot `rank` of every rank's gather buffer (NVLink)
            const long long parity = *reinterpret_cast<const volatile uint32_t*>(push.pa.epoch) & 1u;  // double buffer
            const long long slot = parity * push.pa.world * push.n_push + static_cast<long long>(push.pa.rank) * push.n_push +
                                   (row / seq_len - push.first_seq);
            for (int r = 0; r < push.pa.world; ++r)
              store8_f(static_cast<float*>(push.pa.peer_buf[r]) + slot * hidden + c, o);
          }
        }
    }
    __syncthreads();  // every warp is done with this stage
    if (tid == 0) {
      const int nt = tile + LNS_STAGES * gridDim.x;
      if (nt < n_tiles) issue(nt, stage);
    }
  }
  if (push.pa.world > 0) peer_signal_grid_done(push.pa, 0);
}

// DROP: the LayerNorm input was x + dropout(d) (HF BertSelfOutput / BertOutput).  dx stays the gradient of the residual
// branch; dxm = mask . dx / (1 - p) is the gradient of the dense output d (operand of its dgrad / wgrad GEMMs) and dcol
// becomes the column sum of dxm (the dense bias gradient) -- phase B recomputes dx from the staged tile and the row
// statistics and applies the keep bits phase A left in shared memory.
template <int VPL, bool DROP>
__global__ void __launch_bounds__(LNS_THREADS, 2)
ln_bwd_staged_kernel(const __half* __restrict__ dy, const __half* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, __half* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcol, int rows,
                     int hidden, float out_scale, __half* __restrict__ dxm, const cdr_dropout drop) {
  extern __shared__ __align__(128) uint8_t lns_smem[];  // [stage][dy tile | x tile]
  __shared__ uint64_t full[LNS_STAGES];
  __shared__ float4 rowstat[LNS_ROWS];  // mean, rstd, rstd*c1, rstd*c2 of the rows of the current tile
  __shared__ uint8_t keepb[DROP ? LNS_ROWS : 1][DROP ? 128 : 1];  // DROP: keep bits of (row of the tile, column group)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nvec = hidden >> 3;
  const uint32_t row_bytes = static_cast<uint32_t>(hidden) * 2u;
  const uint32_t tile_bytes = row_bytes * LNS_ROWS;
  const int n_tiles = (rows + LNS_ROWS - 1) / LNS_ROWS;
  if (tid == 0) {
    for (int i = 0; i < LNS_STAGES; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_launch();
  auto issue = [&](int tile, int stage) {
    const int r0 = tile * LNS_ROWS;
    const uint32_t bytes = static_cast<uint32_t>(min(LNS_ROWS, rows - r0)) * row_bytes;
    uint8_t* dst = lns_smem + stage * 2 * tile_bytes;
    mbar_expect_tx(&full[stage], 2 * bytes);
    bulk_load(dst, dy + static_cast<long long>(r0) * hidden, bytes, &full[stage]);
    bulk_load(dst + tile_bytes, x + static_cast<long long>(r0) * hidden, bytes, &full[stage]);
  };
  if (tid == 0)
    for (int i = 0; i < LNS_STAGES; ++i) {
      const int t = blockIdx.x + i * gridDim.x;
      if (t < n_tiles) issue(t, i);
    }
  float g[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    if (lane + 32 * i < nvec) load8_f(gamma + 8 * (lane + 32 * i), g[i]);
  // phase-B ownership: thread -> 8 columns (cgp) of the rows [rh * 4, rh * 4 + 4) of each tile
  const int cgp = tid % 128, rh = tid / 128;
  const bool colthread = cgp < nvec;
  float a_dg[8], a_db[8], a_s3[8], a_s4[8], a_c = 0.f;  // DROP: a_s3 accumulates the masked dx directly
#pragma unroll
  for (int k = 0; k < 8; ++k) a_dg[k] = a_db[k] = a_s3[k] = a_s4[k] = 0.f;
  DropCtx dc{};
  float gcol[8];
  if constexpr (DROP) {
    dc = drop_load(drop);
#pragma unroll
    for (int k = 0; k < 8; ++k) gcol[k] = 0.f;
    if (colthread) load8_f(gamma + 8 * cgp, gcol);
  }

  // per-row statistics of the NEXT tile are fetched one iteration ahead (their global latency would otherwise
  // sit on the critical path of every tile)
  float mean_nx = 0.f, rstd_nx = 0.f;
  if (blockIdx.x * LNS_ROWS + warp < rows) {
    mean_nx = mean_in[blockIdx.x * LNS_ROWS + warp];
    rstd_nx = rstd_in[blockIdx.x * LNS_ROWS + warp];
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int stage = it % LNS_STAGES;
    const float mean = mean_nx, rstd = rstd_nx;
    {
      const long long nrow = static_cast<long long>(tile + gridDim.x) * LNS_ROWS + warp;
      if (nrow < rows) {
        mean_nx = mean_in[nrow];
        rstd_nx = rstd_in[nrow];
      }
    }
    mbar_wait(&full[stage], (it / LNS_STAGES) & 1);
    const __half* dys = reinterpret_cast<const __half*>(lns_smem + stage * 2 * tile_bytes);
    const __half* xs = reinterpret_cast<const __half*>(lns_smem + stage * 2 * tile_bytes + tile_bytes);
    const int r0 = tile * LNS_ROWS;
    {  // ---- phase A: one row per warp
      const int row = r0 + warp;
      if (row < rows) {
        float xh[VPL][8], gy[VPL][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          if (lane + 32 * i < nvec) {
            float xv[8], d[8];
            load8_h(xs + warp * hidden + 8 * (lane + 32 * i), xv);
            load8_h(dys + warp * hidden + 8 * (lane + 32 * i), d);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              xh[i][k] = (xv[k] - mean) * rstd;
              gy[i][k] = d[k] * g[i][k];
              s1 += gy[i][k];
              s2 += gy[i][k] * xh[i][k];
            }
          }
        const float c1 = warp_sum(s1) / hidden;
        const float c2 = warp_sum(s2) / hidden;
        if (lane == 0) rowstat[warp] = make_float4(mean, rstd, rstd * c1, rstd * c2);
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          if (lane + 32 * i < nvec) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = rstd * (gy[i][k] - c1 - xh[i][k] * c2);
            store8_h(dx + static_cast<long long>(row) * hidden + 8 * (lane + 32 * i), o);
            if constexpr (DROP) {
              const uint32_t keep = drop_keep8(dc, drop_group(dc, row, 8 * (lane + 32 * i), hidden));
              keepb[warp][lane + 32 * i] = static_cast<uint8_t>(keep);
              drop_apply8(dc, keep, o);
              store8_h(dxm + static_cast<long long>(row) * hidden + 8 * (lane + 32 * i), o);
            }
          }
      } else if (lane == 0) {
        rowstat[warp] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
    if (colthread) {  // ---- phase B: column partial sums from the same staged tile
#pragma unroll
      for (int rr = 0; rr < LNS_ROWS / 2; ++rr) {
        const int r = rh * (LNS_ROWS / 2) + rr;
        if (r0 + r < rows) {
          const float4 st = rowstat[r];
          float d[8], xv[8];
          load8_h(dys + r * hidden + 8 * cgp, d);
          load8_h(xs + r * hidden + 8 * cgp, xv);
          uint32_t keep = 0u;
          if constexpr (DROP) keep = keepb[r][cgp];
          else a_c += st.z;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float xh = (xv[k] - st.x) * st.y;
            a_dg[k] = fmaf(d[k], xh, a_dg[k]);
            a_db[k] += d[k];
            if constexpr (DROP) {  // dx of this element again (what phase A stored), summed where it was kept
              const float dxv = fmaf(st.y * gcol[k], d[k], -st.z) - st.w * xh;
              a_s3[k] += ((keep >> k) & 1u) ? dxv : 0.f;
            } else {
              a_s3[k] = fmaf(st.y, d[k], a_s3[k]);
              a_s4[k] = fmaf(st.w, xh, a_s4[k]);
            }
          }
        }
      }
    }
    __syncthreads();  // stage and rowstat are free
    if (tid == 0) {
      const int nt = tile + LNS_STAGES * gridDim.x;
      if (nt < n_tiles) issue(nt, stage);
    }
  }
  // ---- block totals: the two row halves meet in shared memory, then one atomic per column and tensor
  float* red = reinterpret_cast<float*>(lns_smem);  // all bulk copies have been consumed
  __shared__ float red_c[2];
  __syncthreads();
  if (cgp == 0) red_c[rh] = a_c;
  if (colthread && rh == 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[(0 * nvec + cgp) * 8 + k] = a_dg[k];
      red[(1 * nvec + cgp) * 8 + k] = a_db[k];
      red[(2 * nvec + cgp) * 8 + k] = a_s3[k];
      red[(3 * nvec + cgp) * 8 + k] = a_s4[k];
    }
  }
  __syncthreads();
  if (colthread && rh == 0) {
    const float ctot = red_c[0] + red_c[1];
    float gg[8], o[8];
    load8_f(gamma + 8 * cgp, gg);
    if (dgamma != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (a_dg[k] + red[(0 * nvec + cgp) * 8 + k]) * out_scale;
      atomic_add8(dgamma + 8 * cgp, o);
    }
    if (dbeta != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (a_db[k] + red[(1 * nvec + cgp) * 8 + k]) * out_scale;
      atomic_add8(dbeta + 8 * cgp, o);
    }
    if (dcol != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if constexpr (DROP)
          o[k] = (a_s3[k] + red[(2 * nvec + cgp) * 8 + k]) * dc.scale * out_scale;
        else
          o[k] = (gg[k] * (a_s3[k] + red[(2 * nvec + cgp) * 8 + k]) - (a_s4[k] + red[(3 * nvec + cgp) * 8 + k]) - ctot) *
                 out_scale;
      }
      atomic_add8(dcol + 8 * cgp, o);
    }
  }
}


// ------------------------------------------------------------------------------------------ staged, row-owner backward
// Second generation of the staged backward: a warp owns a row end to end AND the column partial sums of the rows it
// has processed (dgamma, dbeta, bias gradient: 3 x hidden / 32 accumulators per lane), so every staged element is read
// from shared memory once and no value is recomputed (the two-phase kernel above re-reads the tile and re-derives
// xhat / dx per column thread, which made it issue-bound at half of the HBM rate).  One block of 8 warps per SM, a
// deeper ring (all the shared memory of the SM in flight), cross-warp reduction of the accumulators once per block.
// Round 2: per-stage "empty" mbarriers instead of a __syncthreads per tile, arithmetic on packed fp32 pairs.
constexpr int LNR_THREADS = 256;

template <int VPL, bool DROP>
__global__ void __launch_bounds__(LNR_THREADS, 1)
ln_bwd_rows_kernel(const __half* __restrict__ dy, const __half* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ mean_in, const float* __restrict__ rstd_in, __half* __restrict__ dx,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcol, int rows, int hidden,
                   float out_scale, __half* __restrict__ dxm, const cdr_dropout drop, int stages) {
  extern __shared__ __align__(128) uint8_t lns_smem[];  // [stage][dy tile | x tile]
  __shared__ uint64_t full[8], empty[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nvec = hidden >> 3;
  const uint32_t row_bytes = static_cast<uint32_t>(hidden) * 2u;
  const uint32_t tile_bytes = row_bytes * LNS_ROWS;
  const int n_tiles = (rows + LNS_ROWS - 1) / LNS_ROWS;
  if (tid == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], LNS_ROWS);  // one arrival per row-owner warp: its row of the stage sits in registers
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_launch();
  auto issue = [&](int tile, int stage) {
    const int r0 = tile * LNS_ROWS;
    const uint32_t bytes = static_cast<uint32_t>(min(LNS_ROWS, rows - r0)) * row_bytes;
    uint8_t* dst = lns_smem + stage * 2 * tile_bytes;
    mbar_expect_tx(&full[stage], 2 * bytes);
    bulk_load(dst, dy + static_cast<long long>(r0) * hidden, bytes, &full[stage]);
    bulk_load(dst + tile_bytes, x + static_cast<long long>(r0) * hidden, bytes, &full[stage]);
  };
  if (tid == 0)
    for (int i = 0; i < stages; ++i) {
      const int t = blockIdx.x + i * gridDim.x;
      if (t < n_tiles) issue(t, i);
    }
  // column partial sums and gamma as packed fp32 pairs (FFMA2 / FADD2: half the fp32-pipe instructions)
  f32x2 g[VPL][4], a_dg[VPL][4], a_db[VPL][4], a_dc[VPL][4];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float gv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) gv[k] = 0.f;
    if (lane + 32 * i < nvec) load8_f(gamma + 8 * (lane + 32 * i), gv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      g[i][k] = pk2(gv[2 * k], gv[2 * k + 1]);
      a_dg[i][k] = a_db[i][k] = a_dc[i][k] = pk2(0.f);
    }
  }
  DropCtx dc{};
  if constexpr (DROP) dc = drop_load(drop);
  float mean_nx = 0.f, rstd_nx = 0.f;
  if (blockIdx.x * LNS_ROWS + warp < rows) {
    mean_nx = mean_in[blockIdx.x * LNS_ROWS + warp];
    rstd_nx = rstd_in[blockIdx.x * LNS_ROWS + warp];
  }
  // No block-wide barrier in the tile loop: a warp releases its share of a stage (mbarrier arrive) as soon as its row
  // is in registers, and thread 0 refills the stage of the PREVIOUS tile at the top of its next iteration -- it waits
  // for rows that were read a whole tile ago, so the warps drift apart instead of meeting after every tile.
  int stage = 0, it = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    if (tid == 0 && it > 0) {
      const int ps = (it - 1) % stages;
      const int nt = tile + (stages - 1) * gridDim.x;  // (tile - gridDim.x) + stages * gridDim.x
      if (nt < n_tiles) {
        mbar_wait(&empty[ps], ((it - 1) / stages) & 1);
        issue(nt, ps);
      }
    }
    const float mean = mean_nx, rstd = rstd_nx;
    {
      const long long nrow = static_cast<long long>(tile + gridDim.x) * LNS_ROWS + warp;
      if (nrow < rows) {
        mean_nx = mean_in[nrow];
        rstd_nx = rstd_in[nrow];
      }
    }
    mbar_wait(&full[stage], phase);
    const __half* dys = reinterpret_cast<const __half*>(lns_smem + stage * 2 * tile_bytes) + warp * hidden;
    const __half* xs = reinterpret_cast<const __half*>(lns_smem + stage * 2 * tile_bytes + tile_bytes) + warp * hidden;
    const int row = tile * LNS_ROWS + warp;
    const bool live = row < rows;
    f32x2 xh[VPL][4], d[VPL][4];
    f32x2 s1p = pk2(0.f), s2p = pk2(0.f);
    uint32_t kb[VPL];  // DROP with stored masks (written by the forward GEMM epilogue): one byte per 8 elements
#pragma unroll
    for (int i = 0; i < VPL; ++i) kb[i] = 0u;
    if constexpr (DROP) {
      if (live && dc.keep_bits != nullptr) {
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          if (lane + 32 * i < nvec) kb[i] = __ldg(dc.keep_bits + static_cast<long long>(row) * nvec + lane + 32 * i);
      }
    }
    if (live) {
      const f32x2 nmean = pk2(-mean), rs = pk2(rstd);
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (lane + 32 * i < nvec) {
          float xv[8], dv[8];
          load8_h(xs + 8 * (lane + 32 * i), xv);
          load8_h(dys + 8 * (lane + 32 * i), dv);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            xh[i][k] = mul2(add2(pk2(xv[2 * k], xv[2 * k + 1]), nmean), rs);
            d[i][k] = pk2(dv[2 * k], dv[2 * k + 1]);
            const f32x2 gy = mul2(d[i][k], g[i][k]);
            s1p = add2(s1p, gy);
            s2p = fma2(gy, xh[i][k], s2p);
            a_dg[i][k] = fma2(d[i][k], xh[i][k], a_dg[i][k]);
            a_db[i][k] = add2(a_db[i][k], d[i][k]);
          }
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);  // this warp's row has been read: its share of the stage is free
    if (live) {
      float s1a, s1b, s2a, s2b;
      upk2(s1p, s1a, s1b);
      upk2(s2p, s2a, s2b);
      const float c1 = warp_sum(s1a + s1b) / hidden;
      const float c2 = warp_sum(s2a + s2b) / hidden;
      const f32x2 nc1 = pk2(-c1), nc2 = pk2(-c2), rs = pk2(rstd);
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (lane + 32 * i < nvec) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 4; ++k)  // rstd * (dy * gamma - c1 - xhat * c2)
            upk2(mul2(rs, fma2(xh[i][k], nc2, fma2(d[i][k], g[i][k], nc1))), o[2 * k], o[2 * k + 1]);
          store8_h(dx + static_cast<long long>(row) * hidden + 8 * (lane + 32 * i), o);
          if constexpr (DROP) {
            const uint32_t keep = dc.keep_bits != nullptr ? kb[i]
                                                          : drop_keep8(dc, drop_group(dc, row, 8 * (lane + 32 * i), hidden));
            drop_apply8(dc, keep, o);
            store8_h(dxm + static_cast<long long>(row) * hidden + 8 * (lane + 32 * i), o);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) a_dc[i][k] = add2(a_dc[i][k], pk2(o[2 * k], o[2 * k + 1]));  // DROP: the dropped (and rescaled) gradient
        }
    }
    if (++stage == stages) { stage = 0; phase ^= 1; }
  }
  // ---- block totals: warp partials meet in shared memory (all bulk copies have been consumed), one atomic per column
  float* red = reinterpret_cast<float*>(lns_smem);  // [8 warps][hidden]
  auto reduce_emit = [&](f32x2 (&acc)[VPL][4], float* dst) {
    if (dst == nullptr) return;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (lane + 32 * i < nvec) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) upk2(acc[i][k], v[2 * k], v[2 * k + 1]);
        store8_f(red + warp * hidden + 8 * (lane + 32 * i), v);
      }
    __syncthreads();
    for (int c = tid; c < hidden; c += LNR_THREADS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < LNS_ROWS; ++w) t += red[w * hidden + c];
      atomicAdd(dst + c, t * out_scale);
    }
  };
  reduce_emit(a_dg, dgamma);
  reduce_emit(a_db, dbeta);
  reduce_emit(a_dc, dcol);
}

static int lnr_stages(int hidden) {
  const size_t stage = static_cast<size_t>(2) * LNS_ROWS * hidden * 2;
  int s = static_cast<int>((200 * 1024) / stage);
  if (s > 8) s = 8;
  if (s < 2) s = 2;
  return s;
}

template <bool DROP>
static int launch_ln_bwd_rows(const __half* dy, const __half* x, const float* gamma, const float* mean, const float* rstd,
                              __half* dx, float* dgamma, float* dbeta, float* dcol, int rows, int hidden, float out_scale,
                              __half* dxm, const cdr_dropout& drop, cudaStream_t st) {
  const int stages = lnr_stages(hidden);
  const size_t smem = static_cast<size_t>(stages) * 2 * LNS_ROWS * hidden * 2;
  const int n_tiles = (rows + LNS_ROWS - 1) / LNS_ROWS;
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  const int vpl = (hidden / 8 + 31) / 32;
#define LNR_BWD(V)                                                                                                   \
  do {                                                                                                               \
    static bool cfg = false;                                                                                         \
    if (!cfg) {                                                                                                      \
      CDR_CUDA(cudaFuncSetAttribute(ln_bwd_rows_kernel<V, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                    200 * 1024));                                                                    \
      cfg = true;                                                                                                    \
    }                                                                                                                \
    CDR_CUDA(launch_pdl(ln_bwd_rows_kernel<V, DROP>, dim3(grid), dim3(LNR_THREADS), smem, st, dy, x, gamma, mean,     \
                        rstd, dx, dgamma, dbeta, dcol, rows, hidden, out_scale, dxm, drop, stages));                 \
  } while (0)
  if (vpl <= 1) LNR_BWD(1);
  else if (vpl <= 2) LNR_BWD(2);
  else if (vpl <= 3) LNR_BWD(3);
  else LNR_BWD(4);
#undef LNR_BWD
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

static bool ln_bwd_rows_enabled() {  // CDR_LN_BWD=phases selects the first-generation two-phase kernel (A/B runs)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CDR_LN_BWD");
    v = (e != nullptr && e[0] == 'p') ? 0 : 1;
  }
  return v != 0;
}

static int lns_grid(int rows, size_t smem_per_block) {
  const int n_tiles = (rows + LNS_ROWS - 1) / LNS_ROWS;
  int per_sm = static_cast<int>((200 * 1024) / (smem_per_block + 1024));
  if (per_sm > 2) per_sm = 2;  // __launch_bounds__(256, 2): two resident blocks per SM
  if (per_sm < 1) per_sm = 1;
  const int g = sm_count() * per_sm;
  return n_tiles < g ? n_tiles : g;
}

// out[c] += scale * sum_r x[r, c]      (fp16 in, fp32 accumulate; bias gradients)
__global__ void __launch_bounds__(256)
colsum_kernel(const __half* __restrict__ x, float* __restrict__ out, int rows, int cols, long long ld, float scale,
              int rows_per_block) {
  // thread handles 8 columns; blockDim.x threads cover cols in strides; blockIdx.y picks the row slab
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, rows);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int r = r0; r < r1; ++r) {
    float v[8];
    load8_h(x + static_cast<long long>(r) * ld + c, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] *= scale;
  atomic_add8(out + c, acc);
}

// out = dropout(x) on a contiguous fp16 [rows, cols] tensor, one Philox group (8 elements, 16 bytes) per thread step
__global__ void __launch_bounds__(256)
dropout_kernel(const __half* __restrict__ x, __half* __restrict__ out, long long groups, int groups_per_row,
               const cdr_dropout drop) {
  const DropCtx dc = drop_load(drop);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long gi = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; gi < groups; gi += stride) {
    const long long m = gi / groups_per_row;
    const int cg = static_cast<int>(gi - m * groups_per_row);
    float v[8];
    load8_h(x + gi * 8, v);
    drop_apply8(dc, drop_keep8(dc, static_cast<uint32_t>(m * dc.row_mul * groups_per_row + cg)), v);
    store8_h(out + gi * 8, v);
  }
}

__global__ void cast_f32_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n) {
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float v[8];
    load8_f(src + i, v);
    store8_h(dst + i, v);
  } else {
    for (; i < n; ++i) dst[i] = __float2half_rn(src[i]);
  }
}

// One launch for a whole table of fp32 -> fp16 casts (and fp32 -> fp32 copies): the per-step refresh of
// every fp16 weight shadow of the encoder.  blockIdx.y = table entry, blockIdx.x strides its elements.
__global__ void __launch_bounds__(256)
cast_multi_kernel(const cdr_cast_item* __restrict__ items) {
  const cdr_cast_item it = items[blockIdx.y];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < it.n; i += stride) {
    if (i + 8 <= it.n) {
      float v[8];
      load8_f(it.src + i, v);
      if (it.dst_f32) store8_f(static_cast<float*>(it.dst) + i, v);
      else store8_h(static_cast<__half*>(it.dst) + i, v);
    } else {
      for (long long j = i; j < it.n; ++j) {
        if (it.dst_f32) static_cast<float*>(it.dst)[j] = it.src[j];
        else static_cast<__half*>(it.dst)[j] = __float2half_rn(it.src[j]);
      }
    }
  }
}

template <int MODE>
static int launch_ln_fwd(const __half* x, const int64_t* ids, const float* word, const float* pos, const float* type0,
                         const float* gamma, const float* beta, __half* y, float* mean, float* rstd, float* cls_out,
                         int rows, int hidden, int seq_len, int vocab, float eps, cudaStream_t st) {
  const int nvec = hidden / 8;
  const int vpl = (nvec + 31) / 32;
  const dim3 grid((rows + LN_WARPS - 1) / LN_WARPS), block(LN_WARPS * 32);
#define LN_FWD(V)                                                                                             \
  ln_fwd_kernel<V, MODE><<<grid, block, 0, st>>>(x, ids, word, pos, type0, gamma, beta, y, mean, rstd, cls_out, \
                                                 rows, hidden, seq_len, vocab, eps)
  if (vpl <= 1) LN_FWD(1);
  else if (vpl <= 2) LN_FWD(2);
  else if (vpl <= 3) LN_FWD(3);
  else if (vpl <= 4) LN_FWD(4);
  else LN_FWD(8);
#undef LN_FWD
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

template <int MODE>
static int launch_ln_bwd(const __half* dy, const __half* x, const int64_t* ids, const float* word, const float* pos,
                         const float* type0, const float* gamma, const float* mean, const float* rstd,
                         const float* dy_cls, __half* dx, float* dgamma, float* dbeta, float* dcol, float* dword,
                         float* dpos, int n_seq, int hidden, int seq_len, int vocab, int pad_id, float in_scale,
                         float out_scale, cudaStream_t st) {
  const int nvec = hidden / 8;
  const int vpl = (nvec + 31) / 32;
  // aim for ~4 waves of blocks: grid.x = seq_len positions, grid.y = sequence splits
  int ysplit = (4 * sm_count() + seq_len - 1) / seq_len;
  const int max_split = (n_seq + LN_WARPS - 1) / LN_WARPS;
  if (ysplit > max_split) ysplit = max_split;
  if (ysplit < 1) ysplit = 1;
  const dim3 grid(seq_len, ysplit), block(LN_WARPS * 32);
  const size_t smem = sizeof(float) * hidden;
#define LN_BWD(V)                                                                                                 \
  ln_bwd_kernel<V, MODE><<<grid, block, smem, st>>>(dy, x, ids, word, pos, type0, gamma, mean, rstd, dy_cls, dx,    \
                                                    dgamma, dbeta, dcol, dword, dpos, n_seq, hidden, seq_len, vocab, \
                                                    pad_id, in_scale, out_scale)
  if (vpl <= 1) LN_BWD(1);
  else if (vpl <= 2) LN_BWD(2);
  else if (vpl <= 3) LN_BWD(3);
  else if (vpl <= 4) LN_BWD(4);
  else LN_BWD(8);
#undef LN_BWD
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

static int check_hidden(int hidden) {
  CDR_REQUIRE(hidden > 0 && hidden % 8 == 0 && hidden <= 32 * 8 * MAX_VPL, "hidden must be a multiple of 8 and <= %d (got %d)",
              32 * 8 * MAX_VPL, hidden);
  return CDR_OK;
}

}  // namespace cdr

using namespace cdr;

extern "C" {

int cdr_embed_ln_fwd(const int64_t* ids, const float* word, const float* pos, const float* type0, const float* gamma,
                     const float* beta, void* out, float* mean, float* rstd, int32_t n_seq, int32_t seq_len,
                     int32_t hidden, int32_t vocab, float eps, void* stream) {
  if (int rc = check_hidden(hidden)) return rc;
  CDR_REQUIRE(ids && word && pos && type0 && gamma && beta && out, "cdr_embed_ln_fwd: null pointer");
  if (n_seq <= 0 || seq_len <= 0) return CDR_OK;
  return launch_ln_fwd<1>(nullptr, ids, word, pos, type0, gamma, beta, static_cast<__half*>(out), mean, rstd, nullptr,
                          n_seq * seq_len, hidden, seq_len, vocab, eps, static_cast<cudaStream_t>(stream));
}

int cdr_embed_ln_bwd(const void* dy, const int64_t* ids, const float* word, const float* pos, const float* type0,
                     const float* gamma, const float* mean, const float* rstd, float* dword, float* dpos, float* dtype0,
                     float* dgamma, float* dbeta, int32_t n_seq, int32_t seq_len, int32_t hidden, int32_t vocab,
                     int32_t pad_id, float in_scale, float out_scale, void* stream) {
  if (int rc = check_hidden(hidden)) return rc;
  CDR_REQUIRE(dy && ids && word && pos && type0 && gamma && mean && rstd && dword && dpos && dtype0 && dgamma && dbeta,
              "cdr_embed_ln_bwd: null pointer");
  if (n_seq <= 0 || seq_len <= 0) return CDR_OK;
  return launch_ln_bwd<1>(static_cast<const __half*>(dy), nullptr, ids, word, pos, type0, gamma, mean, rstd, nullptr,
                          nullptr, dgamma, dbeta, dtype0, dword, dpos, n_seq, hidden, seq_len, vocab, pad_id, in_scale,
                          out_scale, static_cast<cudaStream_t>(stream));
}

static int ln_fwd_impl(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                       float* cls_out, int32_t n_seq, int32_t seq_len, int32_t hidden, float eps, const LnPush& push,
                       void* stream) {
  if (int rc = check_hidden(hidden)) return rc;
  CDR_REQUIRE(x && gamma && beta && y, "cdr_ln_fwd: null pointer");
  if (n_seq <= 0 || seq_len <= 0) return CDR_OK;
  CDR_REQUIRE(push.pa.world == 0 || (hidden <= 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0),
              "cdr_ln_fwd_push: needs hidden <= 1024 and a 16-byte aligned input");
  if (hidden <= 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int rows = n_seq * seq_len;
    const size_t smem = static_cast<size_t>(LNS_STAGES) * LNS_ROWS * hidden * 2;
    const int vpl = (hidden / 8 + 31) / 32;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = lns_grid(rows, smem);
#define LNS_FWD(V)                                                                                                   \
  do {                                                                                                               \
    static bool cfg = false;                                                                                         \
    if (!cfg) {                                                                                                      \
      CDR_CUDA(cudaFuncSetAttribute(ln_fwd_staged_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); \
      cfg = true;                                                                                                    \
    }                                                                                                                \
    CDR_CUDA(launch_pdl(ln_fwd_staged_kernel<V>, dim3(grid), dim3(LNS_THREADS), smem, st,                             \
                        static_cast<const __half*>(x), gamma, beta, static_cast<__half*>(y), mean, rstd, cls_out,     \
                        rows, hidden, seq_len, eps, push));                                                          \
  } while (0)
    if (vpl <= 1) LNS_FWD(1);
    else if (vpl <= 2) LNS_FWD(2);
    else if (vpl <= 3) LNS_FWD(3);
    else LNS_FWD(4);
#undef LNS_FWD
    CDR_LAUNCH_CHECK();
    return CDR_OK;
  }
  return launch_ln_fwd<0>(static_cast<const __half*>(x), nullptr, nullptr, nullptr, nullptr, gamma, beta,
                          static_cast<__half*>(y), mean, rstd, cls_out, n_seq * seq_len, hidden, seq_len, 0, eps,
                          static_cast<cudaStream_t>(stream));
}

int cdr_ln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, float* cls_out,
               int32_t n_seq, int32_t seq_len, int32_t hidden, float eps, void* stream) {
  LnPush push{};
  return ln_fwd_impl(x, gamma, beta, y, mean, rstd, cls_out, n_seq, seq_len, hidden, eps, push, stream);
}

int cdr_ln_fwd_push(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                    float* cls_out, int32_t n_seq, int32_t seq_len, int32_t hidden, float eps, int32_t first_seq,
                    const cdr_peer_args* peers, void* stream) {
  if (int rc = peer_check(peers, "cdr_ln_fwd_push")) return rc;
  CDR_REQUIRE(first_seq >= 0 && first_seq < n_seq, "cdr_ln_fwd_push: first_seq out of range");
  LnPush push{};
  push.pa = *peers;
  push.first_seq = first_seq;
  push.n_push = n_seq - first_seq;
  return ln_fwd_impl(x, gamma, beta, y, mean, rstd, cls_out, n_seq, seq_len, hidden, eps, push, stream);
}

int cdr_ln_bwd(const void* dy, const float* dy_cls, const void* x, const float* gamma, const float* mean,
               const float* rstd, void* dx, float* dgamma, float* dbeta, float* dbias, float* row_ws, int32_t n_seq,
               int32_t seq_len, int32_t hidden, float in_scale, float out_scale, void* stream) {
  if (int rc = check_hidden(hidden)) return rc;
  CDR_REQUIRE((dy || dy_cls) && x && gamma && mean && rstd && dx, "cdr_ln_bwd: null pointer");
  if (n_seq <= 0 || seq_len <= 0) return CDR_OK;
  if (dy != nullptr && dy_cls == nullptr && hidden <= 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    // one staged pass: every row of dy and x is read from HBM once
    const int rows = n_seq * seq_len;
    if (ln_bwd_rows_enabled())
      return launch_ln_bwd_rows<false>(static_cast<const __half*>(dy), static_cast<const __half*>(x), gamma, mean, rstd,
                                       static_cast<__half*>(dx), dgamma, dbeta, dbias, rows, hidden, out_scale, nullptr,
                                       cdr_dropout{}, static_cast<cudaStream_t>(stream));
    const size_t smem = static_cast<size_t>(LNS_STAGES) * 2 * LNS_ROWS * hidden * 2;
    const int vpl = (hidden / 8 + 31) / 32;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = lns_grid(rows, smem);
#define LNS_BWD(V)                                                                                                   \
  do {                                                                                                               \
    static bool cfg = false;                                                                                         \
    if (!cfg) {                                                                                                      \
      CDR_CUDA(cudaFuncSetAttribute(ln_bwd_staged_kernel<V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                    112 * 1024));                                                                    \
      cfg = true;                                                                                                    \
    }                                                                                                                \
    CDR_CUDA(launch_pdl(ln_bwd_staged_kernel<V, false>, dim3(grid), dim3(LNS_THREADS), smem, st,                      \
                        static_cast<const __half*>(dy), static_cast<const __half*>(x), gamma, mean, rstd,             \
                        static_cast<__half*>(dx), dgamma, dbeta, dbias, rows, hidden, out_scale,                     \
                        static_cast<__half*>(nullptr), cdr_dropout{}));                                              \
  } while (0)
    if (vpl <= 1) LNS_BWD(1);
    else if (vpl <= 2) LNS_BWD(2);
    else if (vpl <= 3) LNS_BWD(3);
    else LNS_BWD(4);
#undef LNS_BWD
    CDR_LAUNCH_CHECK();
    return CDR_OK;
  }
  if (row_ws != nullptr && dy != nullptr && dy_cls == nullptr) {
    // split path: dx pass + column-sum pass
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rows = n_seq * seq_len;
    const int nvec = hidden / 8, vpl = (nvec + 31) / 32;
    const __half* dyh = static_cast<const __half*>(dy);
    const __half* xh = static_cast<const __half*>(x);
    float2* rowc = reinterpret_cast<float2*>(row_ws);
    const dim3 grid((rows + LN_WARPS - 1) / LN_WARPS), block(LN_WARPS * 32);
#define LN_DX(V) ln_bwd_dx_kernel<V><<<grid, block, 0, st>>>(dyh, xh, gamma, mean, rstd, static_cast<__half*>(dx), rowc, rows, hidden)
    if (vpl <= 1) LN_DX(1);
    else if (vpl <= 2) LN_DX(2);
    else if (vpl <= 3) LN_DX(3);
    else if (vpl <= 4) LN_DX(4);
    else LN_DX(8);
#undef LN_DX
    CDR_LAUNCH_CHECK();
    if (dgamma || dbeta || dbias) {
      const int threads = nvec < 128 ? ((nvec + 31) / 32) * 32 : 128;
      const int gx = (nvec + threads - 1) / threads;
      int rpb = (rows * gx + sm_count() - 1) / sm_count();  // ~one block per SM
      rpb = ((rpb + LNC_Y - 1) / LNC_Y) * LNC_Y;
      if (rpb < 4 * LNC_Y) rpb = 4 * LNC_Y;
      const int gy = (rows + rpb - 1) / rpb;
      ln_bwd_cols_kernel<<<dim3(gx, gy), dim3(threads, LNC_Y), 0, st>>>(dyh, xh, gamma, mean, rstd, rowc, dgamma, dbeta,
                                                                      dbias, rows, hidden, rpb, out_scale);
      CDR_LAUNCH_CHECK();
    }
    return CDR_OK;
  }
  return launch_ln_bwd<0>(static_cast<const __half*>(dy), static_cast<const __half*>(x), nullptr, nullptr, nullptr,
                          nullptr, gamma, mean, rstd, dy_cls, static_cast<__half*>(dx), dgamma, dbeta, dbias, nullptr,
                          nullptr, n_seq, hidden, seq_len, 0, -1, in_scale, out_scale,
                          static_cast<cudaStream_t>(stream));
}

int cdr_ln_bwd_drop(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, void* dx,
                    void* dx_drop, float* dgamma, float* dbeta, float* dbias, int32_t rows, int32_t hidden,
                    float out_scale, const cdr_dropout* drop, void* stream) {
  if (int rc = check_hidden(hidden)) return rc;
  CDR_REQUIRE(dy && x && gamma && mean && rstd && dx && dx_drop && drop, "cdr_ln_bwd_drop: null pointer");
  CDR_REQUIRE(drop->state != nullptr && drop->threshold > 0 && drop->threshold < 65536,
              "cdr_ln_bwd_drop: needs drop.state and 0 < drop.threshold < 65536");
  CDR_REQUIRE(hidden <= 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && (reinterpret_cast<uintptr_t>(dx_drop) & 15) == 0,
              "cdr_ln_bwd_drop: needs hidden <= 1024 and 16-byte aligned tensors");
  CDR_REQUIRE(static_cast<long long>(rows) * (drop->row_mul > 0 ? drop->row_mul : 1) * (hidden / 8) < (1ll << 32),
              "cdr_ln_bwd_drop: dropout group index overflows 32 bits");
  if (rows <= 0) return CDR_OK;
  if (ln_bwd_rows_enabled())
    return launch_ln_bwd_rows<true>(static_cast<const __half*>(dy), static_cast<const __half*>(x), gamma, mean, rstd,
                                    static_cast<__half*>(dx), dgamma, dbeta, dbias, rows, hidden, out_scale,
                                    static_cast<__half*>(dx_drop), *drop, static_cast<cudaStream_t>(stream));
  const size_t smem = static_cast<size_t>(LNS_STAGES) * 2 * LNS_ROWS * hidden * 2;
  const int vpl = (hidden / 8 + 31) / 32;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = lns_grid(rows, smem);
#define LNS_BWD_DROP(V)                                                                                              \
  do {                                                                                                               \
    static bool cfg = false;                                                                                         \
    if (!cfg) {                                                                                                      \
      CDR_CUDA(cudaFuncSetAttribute(ln_bwd_staged_kernel<V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                    112 * 1024));                                                                    \
      cfg = true;                                                                                                    \
    }                                                                                                                \
    CDR_CUDA(launch_pdl(ln_bwd_staged_kernel<V, true>, dim3(grid), dim3(LNS_THREADS), smem, st,                       \
                        static_cast<const __half*>(dy), static_cast<const __half*>(x), gamma, mean, rstd,             \
                        static_cast<__half*>(dx), dgamma, dbeta, dbias, rows, hidden, out_scale,                     \
                        static_cast<__half*>(dx_drop), *drop));                                                      \
  } while (0)
  if (vpl <= 1) LNS_BWD_DROP(1);
  else if (vpl <= 2) LNS_BWD_DROP(2);
  else if (vpl <= 3) LNS_BWD_DROP(3);
  else LNS_BWD_DROP(4);
#undef LNS_BWD_DROP
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_dropout_f16(const void* x, void* out, int64_t rows, int32_t cols, const cdr_dropout* drop, void* stream) {
  CDR_REQUIRE(x && out && drop, "cdr_dropout_f16: null pointer");
  CDR_REQUIRE(cols > 0 && cols % 8 == 0, "cdr_dropout_f16: cols must be a positive multiple of 8");
  CDR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "cdr_dropout_f16: tensors must be 16-byte aligned");
  CDR_REQUIRE(drop->state != nullptr && drop->threshold < 65536, "cdr_dropout_f16: needs drop.state, threshold < 65536");
  CDR_REQUIRE(rows * (drop->row_mul > 0 ? drop->row_mul : 1) * (cols / 8) < (1ll << 32),
              "cdr_dropout_f16: dropout group index overflows 32 bits");
  if (rows <= 0) return CDR_OK;
  const long long groups = rows * (cols / 8);
  long long blocks = (groups + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  dropout_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<__half*>(out), groups, cols / 8, *drop);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_colsum_f16(const void* x, float* out, int64_t rows, int64_t cols, int64_t ld, float scale, void* stream) {
  CDR_REQUIRE(x && out, "cdr_colsum_f16: null pointer");
  CDR_REQUIRE(cols % 8 == 0 && ld % 8 == 0, "cdr_colsum_f16: cols and ld must be multiples of 8");
  if (rows <= 0 || cols <= 0) return CDR_OK;
  const int threads = 128;
  const int gx = static_cast<int>((cols / 8 + threads - 1) / threads);
  int rpb = static_cast<int>((rows + 4 * sm_count() - 1) / (4 * sm_count()));
  if (rpb < 32) rpb = 32;
  const int gy = static_cast<int>((rows + rpb - 1) / rpb);
  colsum_kernel<<<dim3(gx, gy), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), out, static_cast<int>(rows), static_cast<int>(cols), ld, scale, rpb);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_cast_multi(const cdr_cast_item* items_device, int32_t count, int64_t max_n, void* stream) {
  CDR_REQUIRE(items_device != nullptr && count > 0 && max_n > 0, "cdr_cast_multi: bad arguments");
  long long gx = (max_n + 256 * 8 * 4 - 1) / (256 * 8 * 4);  // ~4 vectors per thread for the largest entry
  if (gx < 1) gx = 1;
  if (gx > 1024) gx = 1024;
  cast_multi_kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(count)), 256, 0,
                      static_cast<cudaStream_t>(stream)>>>(items_device);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_cast_f32_f16(const float* src, void* dst, int64_t n, void* stream) {
  CDR_REQUIRE(src && dst, "cdr_cast_f32_f16: null pointer");
  CDR_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "cdr_cast_f32_f16: pointers must be 16-byte aligned");
  if (n <= 0) return CDR_OK;
  const long long threads = (n + 7) / 8;
  cast_f32_f16_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__half*>(dst), n);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

}  // extern "C"
