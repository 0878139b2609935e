// Fused multi-head attention (K3) for BERT, head_dim 64, seq_len <= 512, forward and backward, on the
// sm_100a tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// A work item is one (sequence, head) (seq_len <= 128: persistent CTAs walk the items; longer sequences: one CTA per
// (sequence, head, 128-row tile)).  Q/K/V come straight out of the packed QKV projection [T, 3H]
// (columns [0,H) = Q, [H,2H) = K, [2H,3H) = V; head h owns 64 contiguous columns), so a single 2-D
// TMA box {64, 128} per operand lands a 128B-swizzled [128 rows][64 halfs] tile that is at the same
// time
//   * a K-major  operand over the head dimension  (S = Q K^T,  dP = dO V^T)
//   * an MN-major operand over the row dimension   (O = P V,  dV = P^T dO,  dK = dS^T Q,  dQ = dS K)
// -- no transposes are ever materialised.  P / dS are written by the softmax threads into shared
// memory in the same swizzled layout as two [128][64] chunks, which again serves both as a K-major A
// operand (P V, dS K) and as an MN-major A operand (P^T dO, dS^T Q).
//
// Rows beyond seq_len inside the 128-row box belong to the next sequence (or are TMA zero fill);
// they are neutralised by the key bias (-inf) on columns and by explicit zeroing of P/dS rows.
//
// Replaces HF eager_attention_forward / SDPA reached through self.bert(...) in
// ANCE/model/models.py:226 and COCO/modeling.py:199-204.  DROP = true adds the dropout of the attention
// probabilities (HF BertSelfAttention: dropout(softmax(S)) V):
//   forward   P~ = mask . P is what multiplies V, the row sum stays that of the undropped P, O = P~ V * s / sum
//   backward  dP = mask . s . (dO V^T);  dV = (mask . s . P)^T dO;  delta = rowsum(P . dP) = rowsum(dO . O) as before
// The keep bits come from Philox counters (dropout.cuh), but NOT inside these kernels: the softmax warps sit on the
// critical path of every item, and 16 Philox calls per query row there cost +0.4 ms per training step (measured).  A
// bandwidth-trivial pre-pass (att_keep_bits_kernel, ~3 M Philox calls per layer at full issue rate) writes one bit per
// probability into a caller-provided buffer (16 B per query row at L <= 128) that the forward and the backward read
// with one vector load per row.
#include "cdr_common.cuh"
#include "dropout.cuh"
#include "tma_host.h"

namespace cdr {

constexpr int ATT_T = 128;  // query rows per tile == max keys
constexpr int ATT_D = 64;   // head dim
constexpr int ATT_TILE_BYTES = ATT_T * ATT_D * 2;  // 16 KB
constexpr float LOG2E = 1.4426950408889634f;

struct AttParams {
  int n_seq, seq_len, heads, hidden;
  const float* key_bias;  // [n_seq, seq_len] additive (0 / -large) or null
  float scale;            // 1/sqrt(64)
  // forward
  __half* out;            // ctx [T, hidden]
  float* lse;             // [n_seq, heads, seq_len]
  // backward
  const __half* o;        // ctx [T, hidden]
  const __half* d_o;      // dctx [T, hidden]
  __half* dqkv;           // [T, 3*hidden]
  float* dbias;           // optional fp32 [3*hidden]: += dbias_scale * column sums of dqkv
  float dbias_scale;
  cdr_dropout drop;       // attention-probability dropout (DROP kernels)
  const uint8_t* keep_bits;  // DROP: [n_seq * heads * seq_len rows][bits_stride bytes], bit c of a row = key c is kept
  int bits_stride;           // bytes per row: ceil(seq_len / 8) rounded up to 16
};

// Philox group of the 8 probabilities (item = seq * heads + head, query row r, keys [8 * kg, 8 * kg + 8))
__device__ __forceinline__ uint32_t att_drop_group(int item, int L, int r, int kg) {
  return (static_cast<uint32_t>(item) * static_cast<uint32_t>(L) + static_cast<uint32_t>(r)) * 64u +
         static_cast<uint32_t>(kg);
}

// byte offset of element (r, c), c in [0,128), inside two 128B-swizzled [128][64-half] chunks
__device__ __forceinline__ uint32_t swz_off(int r, int c) {
  return static_cast<uint32_t>(((c >> 6) * ATT_TILE_BYTES) + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) +
                               ((c & 7) << 1));
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 q;
  __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
  for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
  return q;
}

__device__ __forceinline__ float dot8(const uint4& a, const uint4& b) {
  const __half2* x = reinterpret_cast<const __half2*>(&a);
  const __half2* y = reinterpret_cast<const __half2*>(&b);
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = __half22float2(x[t]), g = __half22float2(y[t]);
    s += f.x * g.x + f.y * g.y;
  }
  return s;
}

// ------------------------------------------------------------------------------------------ forward
// seq_len <= 128.  Persistent and warp-specialised: one CTA per SM walks (sequence, head) items.
//   warp 0       : TMA producer -- a 3-stage ring of (Q, K) tile pairs and a 3-stage ring of V tiles.  (Q, K) of a
//                  stage are handed back as soon as S = Q K^T has been formed (long before the item ends), V after
//                  O = P V, so the loads run up to three items ahead of the softmax; with the earlier 2-stage ring of
//                  whole (Q K V) stages a stage was refilled only after the item's P V and the softmax warps spent most
//                  of their time waiting for the next S (ncu: the s_full wait was the top stall, DRAM at 31 %).
//   warp 1       : tcgen05.mma issuer (one thread) + TMEM owner
//   warps 2..5   : softmax group 0 (even items);  warps 6..9 : softmax group 1 (odd items)
// A softmax thread owns one query row (TMEM lane) end to end: exact row maximum (first pass over S in TMEM),
// P = exp2(S - max) -> fp16 smem + row sum (second pass), then -- once O = P V of ITS item has landed -- the
// epilogue O / sum -> ctx rows, coalesced through a warp-private transposition tile (the first 2 KB of the warp's own
// rows of the group's P buffer, which is dead between P V and the next softmax).  The two groups alternate
// items and everything they touch (S / O columns, P buffer) is per group, so group 1 runs its softmax
// while group 0 waits for its P V: no block-wide barrier anywhere.
// TMEM columns: S[g] at g * 128, O[g] at 256 + g * 64.   smem: 3 x (Q K) 96 KB | 3 x V 48 KB | 2 x P 64 KB | small buffers.
constexpr int ATT_FWD_THREADS = 320;
constexpr int ATT_FWD_NST = 3;  // ring depth (both rings)
constexpr int ATT_FWD_SMEM = (3 * ATT_FWD_NST + 4) * ATT_TILE_BYTES + 8 * 512 + 256 + 1024;

template <bool DROP>
__global__ void __launch_bounds__(ATT_FWD_THREADS, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tma_qkv, const AttParams p, const int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  constexpr int NST = ATT_FWD_NST;
  uint8_t* sQK = smem;                                    // [NST][Q K]
  uint8_t* sV = smem + 2 * NST * ATT_TILE_BYTES;          // [NST]
  uint8_t* sPall = smem + 3 * NST * ATT_TILE_BYTES;       // [2][two [128][64] chunks]
  float* sBiasW = reinterpret_cast<float*>(smem + (3 * NST + 4) * ATT_TILE_BYTES);  // [8 warps][128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (3 * NST + 4) * ATT_TILE_BYTES + 8 * 512);
  uint64_t* full_qk = bar;                // [NST] TMA -> MMA
  uint64_t* full_v = bar + NST;           // [NST]
  uint64_t* qk_empty = bar + 2 * NST;     // [NST] MMA -> TMA: S of the item has been formed
  uint64_t* v_empty = bar + 3 * NST;      // [NST] MMA -> TMA: O of the item has been formed
  uint64_t* s_full = bar + 4 * NST;       // [2] MMA -> softmax group
  uint64_t* s_empty = s_full + 2;         // [2] softmax group (4 warps) -> MMA
  uint64_t* p_full = s_full + 4;          // [2] softmax group (4 warps) -> MMA
  uint64_t* o_full = s_full + 6;          // [2] MMA -> softmax group
  uint64_t* o_empty = s_full + 8;         // [2] softmax group (4 warps) -> MMA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_full + 10);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.seq_len;
  const int n_local = static_cast<int>(blockIdx.x) < n_items
                          ? (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                          : 0;

  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
#pragma unroll
    for (int i = 0; i < 4 * NST; ++i) mbar_init(&bar[i], 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  pdl_wait();  // the prologue above overlapped the previous kernel's tail
  pdl_launch();

  if (warp == 0) {
    // ===================== TMA producer =====================
    // Issue order (Q, K)(i + 1) before V(i): the MMA thread forms S(i - 2) before O(i - 3), so neither wait below
    // ever holds back a load whose stage is already free.
    if (lane == 0) {
      auto load_qk = [&](int it) {
        const int s = it % NST;
        const int item = blockIdx.x + it * gridDim.x;
        const int seq = item / p.heads, h = item % p.heads;
        uint8_t* st = sQK + s * 2 * ATT_TILE_BYTES;
        mbar_wait(&qk_empty[s], ((it / NST) & 1) ^ 1);
        mbar_expect_tx(&full_qk[s], 2 * ATT_TILE_BYTES);
        tma_load_2d(st, &tma_qkv, &full_qk[s], h * ATT_D, seq * L);
        tma_load_2d(st + ATT_TILE_BYTES, &tma_qkv, &full_qk[s], p.hidden + h * ATT_D, seq * L);
      };
      auto load_v = [&](int it) {
        const int s = it % NST;
        const int item = blockIdx.x + it * gridDim.x;
        const int seq = item / p.heads, h = item % p.heads;
        mbar_wait(&v_empty[s], ((it / NST) & 1) ^ 1);
        mbar_expect_tx(&full_v[s], ATT_TILE_BYTES);
        tma_load_2d(sV + s * ATT_TILE_BYTES, &tma_qkv, &full_v[s], 2 * p.hidden + h * ATT_D, seq * L);
      };
      if (n_local > 0) load_qk(0);
      for (int it = 0; it < n_local; ++it) {
        if (it + 1 < n_local) load_qk(it + 1);
        load_v(it);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(ATT_T, ATT_T, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(ATT_T, ATT_D, 0, 1);
      auto issue_s = [&](int jt) {  // S = Q K^T of local item jt into the S columns of group jt & 1
        const int g = jt & 1;
        const int s = jt % NST;
        const uint32_t qa = smem_u32(sQK + s * 2 * ATT_TILE_BYTES), ka = qa + ATT_TILE_BYTES;
        mbar_wait(&full_qk[s], (jt / NST) & 1);
        mbar_wait(&s_empty[g], ((jt >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < ATT_D / 16; ++kk)
          tc_mma_f16(tmem + g * 128, make_smem_desc(qa + kk * 32, 16, 1024), make_smem_desc(ka + kk * 32, 16, 1024),
                     idesc_s, kk > 0);
        tc_commit(&s_full[g]);
        tc_commit(&qk_empty[s]);  // Q and K of this stage are consumed
      };
      if (n_local > 0) issue_s(0);
      for (int it = 0; it < n_local; ++it) {
        const int g = it & 1;
        const int s = it % NST;
        const uint32_t k = (it >> 1) & 1;
        if (it + 1 < n_local) issue_s(it + 1);  // the other group's S first
        const uint32_t pa = smem_u32(sPall + g * 2 * ATT_TILE_BYTES);
        const uint32_t va = smem_u32(sV + s * ATT_TILE_BYTES);
        mbar_wait(&p_full[g], k);
        mbar_wait(&full_v[s], (it / NST) & 1);
        mbar_wait(&o_empty[g], k ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < ATT_T / 16; ++kk)  // O = P V
          tc_mma_f16(tmem + 256 + g * 64, make_smem_desc(pa + (kk >> 2) * ATT_TILE_BYTES + (kk & 3) * 32, 16, 1024),
                     make_smem_desc(va + kk * 2048, 8192, 1024), idesc_o, kk > 0);
        tc_commit(&o_full[g]);
        tc_commit(&v_empty[s]);  // V of this stage is consumed
      }
    }
  } else {
    // ===================== softmax + epilogue threads =====================
    const int sw = warp - 2;         // 0..7
    const int g = sw >> 2;           // group: items with (local index & 1) == g
    const int quad = warp & 3;       // TMEM lane quadrant this warp may touch
    const int r = quad * 32 + lane;  // TMEM lane == query row
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem + lane_off + g * 128;
    const uint32_t t_o = tmem + lane_off + 256 + g * 64;
    uint8_t* sP = sPall + g * 2 * ATT_TILE_BYTES;
    float* wb = sBiasW + sw * 128;   // this warp's private copy of the 128 key-bias values
    uint8_t* tile = sP + quad * 4096;  // the first 16 of this warp's own P rows (chunk 0): dead once O has landed
    const float sl2 = p.scale * LOG2E;
    const float drop_scale = DROP ? p.drop.scale : 1.f;
    auto fetch_bias = [&](int item, int j) -> float {  // x LOG2E at use: nothing waits on the load here
      const int c = j * 32 + lane;
      if (c >= L) return -INFINITY;
      return p.key_bias ? p.key_bias[static_cast<long long>(item / p.heads) * L + c] : 0.f;
    };
    const int first = blockIdx.x + g * gridDim.x, step = 2 * gridDim.x;
    float nb[4] = {0.f, 0.f, 0.f, 0.f};
    if (first < n_items) {
#pragma unroll
      for (int j = 0; j < 4; ++j) nb[j] = fetch_bias(first, j);
    }
    uint32_t k = 0;
    for (int item = first; item < n_items; item += step, k ^= 1) {
      const int seq = item / p.heads, h = item % p.heads;
      const int row0 = seq * L;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) wb[j * 32 + lane] = nb[j] * LOG2E;
      __syncwarp();
      if (item + step < n_items) {
#pragma unroll
        for (int j = 0; j < 4; ++j) nb[j] = fetch_bias(item + step, j);
      }
      uint4 keepw = make_uint4(0u, 0u, 0u, 0u);  // DROP: bit c of word w = key 32 * w + c of this query row is kept
      if constexpr (DROP)  // (in flight while the tensor core still forms S)
        keepw = __ldg(reinterpret_cast<const uint4*>(p.keep_bits + (static_cast<long long>(item) * L + min(r, L - 1)) * 16));
      mbar_wait(&s_full[g], k);
      tc_fence_after();
      // ---- the whole S row (128 fp32) comes into registers with ONE wait and stays there for both passes;
      // its TMEM columns are handed back immediately
      uint32_t v[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + c * 32, reinterpret_cast<uint32_t(&)[32]>(v[c * 32]));
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[g]);
      // ---- pass 1: exact row maximum (scores kept as scaled log2 values)
      // (packed fp32 pairs -- FFMA2 / FADD2 -- wherever two columns go through the same arithmetic: these threads are
      // bound by fp32-pipe issue slots and dependent-issue latency, not by the MUFU)
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four chains: the maximum is latency-, not issue-bound
      f32x2 vp[64];
      {
        const f32x2 sl2p = pk2(sl2);
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          float a, b;
          vp[j] = fma2(pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), sl2p, pk2(wb[2 * j], wb[2 * j + 1]));
          upk2(vp[j], a, b);
          mx4[j & 3] = fmaxf(mx4[j & 3], fmaxf(a, b));
        }
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      if (mx == -INFINITY) mx = 0.f;  // fully masked row: P = 0, output 0
      // ---- pass 2: P = exp2(S - max) -> smem, row sum
      f32x2 sum01 = pk2(0.f), sum23 = pk2(0.f);
      const f32x2 nmx = pk2(-mx);
#pragma unroll
      for (int gq = 0; gq < 16; ++gq) {
        float e[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a, b;
          upk2(add2(vp[gq * 4 + j], nmx), a, b);
          e[2 * j] = fast_ex2(a);
          e[2 * j + 1] = fast_ex2(b);
        }
        sum01 = add2(sum01, add2(pk2(e[0], e[1]), pk2(e[2], e[3])));
        sum23 = add2(sum23, add2(pk2(e[4], e[5]), pk2(e[6], e[7])));
        if constexpr (DROP) {  // the row sum is that of the undropped probabilities; 1 / (1 - p) joins 1 / sum below
          const uint32_t kw = (gq >> 2) == 0 ? keepw.x : (gq >> 2) == 1 ? keepw.y : (gq >> 2) == 2 ? keepw.z : keepw.w;
          const uint32_t keep = kw >> (8 * (gq & 3));
#pragma unroll
          for (int j = 0; j < 8; ++j) e[j] = ((keep >> j) & 1u) ? e[j] : 0.f;
        }
        *reinterpret_cast<uint4*>(sP + swz_off(r, gq * 8)) = pack8(e);
      }
      float sum0, sum1;
      {
        float s0, s1, s2, s3;
        upk2(sum01, s0, s1);
        upk2(sum23, s2, s3);
        sum0 = s0 + s1;
        sum1 = s2 + s3;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      const float sum = sum0 + sum1;
      if (r < L && p.lse) p.lse[static_cast<long long>(item) * L + r] = (mx + log2f(sum)) / LOG2E;
      const float inv = (sum > 0.f ? 1.f / sum : 0.f) * drop_scale;
      // ---- epilogue of the same item: O / sum -> ctx rows (transposed through the warp's tile for 64-byte row
      // segments: a store instruction covers 8 whole rows instead of 32 different lines)
      mbar_wait(&o_full[g], k);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_o + c * 32, v);
        tc_wait_ld();
        if (c == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_empty[g]);
        }
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          float e[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) e[j] = __uint_as_float(v[gq * 8 + j]) * inv;
          *reinterpret_cast<uint4*>(tile + lane * 64 + ((gq ^ ((lane >> 1) & 3)) << 4)) = pack8(e);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = (lane >> 2) + 8 * i;
          const int pc = lane & 3;
          const uint4 q = *reinterpret_cast<const uint4*>(tile + rr * 64 + ((pc ^ ((rr >> 1) & 3)) << 4));
          const int orow = quad * 32 + rr;
          if (orow < L)
            *reinterpret_cast<uint4*>(p.out + static_cast<long long>(row0 + orow) * p.hidden + h * ATT_D + c * 32 +
                                      pc * 8) = q;
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ----------------------------------------------------------------------------------------- backward
constexpr int ATT_BWD_SM_WARPS = 16;   // softmax warps: 4 per TMEM lane quadrant, 32 key columns per thread
constexpr int ATT_BWD_THREADS = 64 + 32 * ATT_BWD_SM_WARPS + 128;  // producer, MMA issuer, softmax, 4 epilogue warps
// (setmaxnreg was tried to move registers from the single-thread roles to the softmax threads, whose loop state spills
// at 80 registers: ptxas fails to allocate any split in which a role shrinks below the launch-time count)
constexpr int ATT_BWD_PRE_BYTES = ATT_BWD_SM_WARPS * 2 * 3 * 128;  // per softmax warp: 2 buffers x (bias | lse | keep bits) x 32 words
constexpr int ATT_BWD_SMEM = 12 * ATT_TILE_BYTES + 2048 + 2048 + 256 + 4 * 2048 + ATT_BWD_PRE_BYTES + 1024;  // tiles | per-warp bias | delta quarters | barriers | epilogue transposition tiles | prefetch slots

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Column totals of a per-lane row vector: on return lane j holds sum over the 32 lanes of v[j].
__device__ __forceinline__ float warp_colsum32(const float (&v)[32], int lane) {
  float a16[16], a8[8], a4[4], a2[2];
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int t = 0; t < 16; ++t)
      a16[t] = (up ? v[t + 16] : v[t]) + __shfl_xor_sync(0xffffffffu, up ? v[t] : v[t + 16], 16);
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int t = 0; t < 8; ++t)
      a8[t] = (up ? a16[t + 8] : a16[t]) + __shfl_xor_sync(0xffffffffu, up ? a16[t] : a16[t + 8], 8);
  }
  {
    const bool up = (lane & 4) != 0;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      a4[t] = (up ? a8[t + 4] : a8[t]) + __shfl_xor_sync(0xffffffffu, up ? a8[t] : a8[t + 4], 4);
  }
  {
    const bool up = (lane & 2) != 0;
#pragma unroll
    for (int t = 0; t < 2; ++t)
      a2[t] = (up ? a4[t + 2] : a4[t]) + __shfl_xor_sync(0xffffffffu, up ? a4[t] : a4[t + 2], 2);
  }
  const bool up = (lane & 1) != 0;
  return (up ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, up ? a2[0] : a2[1], 1);
}

// Backward, seq_len <= 128.  Persistent and warp-specialised: one CTA per SM walks (sequence, head) items.
//   warp 0        : TMA producer -- Q K V dO of item i+1 land in the other smem stage while item i computes
//   warp 1        : tcgen05.mma issuer (one thread) + TMEM owner
//   warps 2..17   : softmax threads; thread = (TMEM lane r = query row, one quarter of the 128 key columns):
//                   P = exp2(S - lse) -> smem;  delta_r = sum_j P dP (exact: the whole row is in one tile; the
//                   four quarter-row partials meet through a 128-thread named barrier);  dS = P (dP - delta) scale
//   warps 18..21  : epilogue threads; thread = key/query row r: dV, dK, dQ rows -> packed fp16 dQKV and the
//                   QKV bias gradient (column sums in fp32 straight from the accumulators)
// so that softmax(i+1), the epilogue of item i, the MMAs and the loads all overlap; only mbarriers connect them.
// MMA order: S, dP (i) | dV (i) as soon as P is in smem | S, dP (i+1) | dK, dQ (i) once dS is.
// TMEM columns: S [0,128) dP [128,256) dV [256,320) dK [320,384) dQ [384,448).
// smem: 2 stages x (Q K V dO) = 128 KB | P 32 KB | dS 32 KB | per-warp key bias | delta halves | barriers.
template <bool DROP>
__global__ void __launch_bounds__(ATT_BWD_THREADS, 1)
fmha_bwd_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                const AttParams p, const int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sStage = smem;                         // [2][Q K V dO]
  uint8_t* sP = smem + 8 * ATT_TILE_BYTES;        // two [128][64] chunks
  uint8_t* sdS = smem + 10 * ATT_TILE_BYTES;
  float* sBiasW = reinterpret_cast<float*>(smem + 12 * ATT_TILE_BYTES);           // [16 warps][32]
  float* sDelta = reinterpret_cast<float*>(smem + 12 * ATT_TILE_BYTES + 2048);    // [4 quarters][128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 12 * ATT_TILE_BYTES + 4096);
  uint64_t* full_qk = bar;          // [2] TMA -> MMA
  uint64_t* full_vdo = bar + 2;     // [2]
  uint64_t* stage_empty = bar + 4;  // [2] MMA -> TMA
  uint64_t* sdp_full = bar + 6;     // MMA -> softmax
  uint64_t* sdp_empty = bar + 7;    // softmax -> MMA (8 warps): S / dP have been read
  uint64_t* p_full = bar + 8;       // softmax -> MMA (8 warps): P is in smem
  uint64_t* ds_full = bar + 9;      // softmax -> MMA (8 warps): dS is in smem
  uint64_t* out_full = bar + 10;    // MMA -> epilogue
  uint64_t* out_empty = bar + 11;   // epilogue -> MMA (4 warps)
  uint64_t* dv_done = bar + 12;     // MMA -> softmax: dV of the item has been formed (the P buffer is free)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 13);
  uint8_t* sEpi = smem + 12 * ATT_TILE_BYTES + 4096 + 256;  // [4 epilogue warps][32 rows][64 B]
  uint32_t* sPre = reinterpret_cast<uint32_t*>(sEpi + 4 * 2048);  // [16 softmax warps][2][bias | lse | keep][32]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.seq_len;

  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
#pragma unroll
    for (int i = 0; i < 6; ++i) mbar_init(&bar[i], 1);
    mbar_init(sdp_full, 1);
    mbar_init(sdp_empty, ATT_BWD_SM_WARPS);
    mbar_init(p_full, ATT_BWD_SM_WARPS);
    mbar_init(ds_full, ATT_BWD_SM_WARPS);
    mbar_init(out_full, 1);
    mbar_init(out_empty, 4);
    mbar_init(dv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  pdl_wait();  // the prologue above overlapped the previous kernel's tail
  pdl_launch();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int s = it & 1;
        const int seq = item / p.heads, h = item % p.heads;
        const int row0 = seq * L;
        uint8_t* st = sStage + s * 4 * ATT_TILE_BYTES;
        mbar_wait(&stage_empty[s], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&full_qk[s], 2 * ATT_TILE_BYTES);
        tma_load_2d(st, &tma_qkv, &full_qk[s], h * ATT_D, row0);
        tma_load_2d(st + ATT_TILE_BYTES, &tma_qkv, &full_qk[s], p.hidden + h * ATT_D, row0);
        mbar_expect_tx(&full_vdo[s], 2 * ATT_TILE_BYTES);
        tma_load_2d(st + 2 * ATT_TILE_BYTES, &tma_qkv, &full_vdo[s], 2 * p.hidden + h * ATT_D, row0);
        tma_load_2d(st + 3 * ATT_TILE_BYTES, &tma_do, &full_vdo[s], h * ATT_D, row0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(ATT_T, ATT_T, 0, 0);
      constexpr uint32_t idesc_tt = make_idesc_f16(ATT_T, ATT_D, 1, 1);
      constexpr uint32_t idesc_nt = make_idesc_f16(ATT_T, ATT_D, 0, 1);
      const uint32_t pa = smem_u32(sP), dsa = smem_u32(sdS);
      // S = Q K^T and dP = dO V^T of local item `jt` (stage jt & 1) -> sdp_full
      auto issue_sdp = [&](int jt) {
        const int s = jt & 1;
        const uint32_t phs = (jt >> 1) & 1;
        const uint32_t qa = smem_u32(sStage + s * 4 * ATT_TILE_BYTES), ka = qa + ATT_TILE_BYTES,
                       va = qa + 2 * ATT_TILE_BYTES, da = qa + 3 * ATT_TILE_BYTES;
        mbar_wait(&full_qk[s], phs);
        mbar_wait(sdp_empty, (jt & 1) ^ 1);  // the softmax threads hold item jt-1's S / dP in registers
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          tc_mma_f16(tmem, make_smem_desc(qa + k * 32, 16, 1024), make_smem_desc(ka + k * 32, 16, 1024), idesc_s, k > 0);
        mbar_wait(&full_vdo[s], phs);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          tc_mma_f16(tmem + 128, make_smem_desc(da + k * 32, 16, 1024), make_smem_desc(va + k * 32, 16, 1024), idesc_s,
                     k > 0);
        tc_commit(sdp_full);
      };
      if (static_cast<int>(blockIdx.x) < n_items) issue_sdp(0);
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = it & 1;
        const uint32_t qa = smem_u32(sStage + s * 4 * ATT_TILE_BYTES), ka = qa + ATT_TILE_BYTES,
                       da = qa + 3 * ATT_TILE_BYTES;
        // S / dP of the NEXT item are issued as soon as its operands have landed and the softmax threads have copied
        // this item's S / dP out of TMEM -- normally while they still form P, so the next softmax never waits for the
        // tensor core (with the fixed order dV(i), S / dP(i+1) the softmax threads spent a third of their time in the
        // sdp_full wait).  The probes are non-blocking: P of this item must not wait for a late load either.
        bool sdp_issued = item + static_cast<int>(gridDim.x) >= n_items;
        bool dv_issued = false;
        for (;;) {
          if (!sdp_issued) {
            const int s1 = (it + 1) & 1;
            const uint32_t ph1 = ((it + 1) >> 1) & 1;
            if (mbar_test_wait(&full_qk[s1], ph1) && mbar_test_wait(&full_vdo[s1], ph1) &&
                mbar_test_wait(sdp_empty, ((it + 1) & 1) ^ 1)) {
              issue_sdp(it + 1);
              sdp_issued = true;
            }
          }
          if (!dv_issued && mbar_test_wait(p_full, ph) && mbar_test_wait(out_empty, ph ^ 1)) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < ATT_T / 16; ++k)  // dV[kv,d] = sum_q P[q,kv] dO[q,d]
              tc_mma_f16(tmem + 256, make_smem_desc(pa + k * 2048, ATT_TILE_BYTES, 1024),
                         make_smem_desc(da + k * 2048, 8192, 1024), idesc_tt, k > 0);
            tc_commit(dv_done);
            dv_issued = true;
          }
          if (dv_issued && mbar_test_wait(ds_full, ph)) break;
        }
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < ATT_T / 16; ++k)  // dK[kv,d] = sum_q dS[q,kv] Q[q,d]
          tc_mma_f16(tmem + 320, make_smem_desc(dsa + k * 2048, ATT_TILE_BYTES, 1024),
                     make_smem_desc(qa + k * 2048, 8192, 1024), idesc_tt, k > 0);
#pragma unroll
        for (int k = 0; k < ATT_T / 16; ++k)  // dQ[q,d] = sum_kv dS[q,kv] K[kv,d]
          tc_mma_f16(tmem + 384, make_smem_desc(dsa + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32, 16, 1024),
                     make_smem_desc(ka + k * 2048, 8192, 1024), idesc_nt, k > 0);
        tc_commit(out_full);
        tc_commit(&stage_empty[s]);  // Q K V dO of this stage (and P / dS) are consumed
        if (!sdp_issued) issue_sdp(it + 1);  // still pending (late loads): now it is the only thing left to do
      }
    }
  } else if (warp < 2 + ATT_BWD_SM_WARPS) {
    // ===================== softmax / dS threads =====================
    const int sw = warp - 2;        // 0..15
    const int quad = warp & 3;      // TMEM lane quadrant this warp may touch
    const int qtr = sw >> 2;        // which 32 key columns
    const int r = quad * 32 + lane; // TMEM lane == query row
    const int c0 = qtr * 32;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const float sl2 = p.scale * LOG2E;
    float* wb = sBiasW + sw * 32;   // this warp's private copy of the 32 key-bias values it needs
    const float ds_scale = p.scale * (DROP ? p.drop.scale : 1.f);
    // Per-item scalars (key bias of this warp's 32 columns, lse of this thread's row, DROP: the row's keep bits of these
    // columns) are fetched ONE ITEM AHEAD with cp.async straight into a per-warp shared-memory slot: no register
    // carries them across the item (at 80 registers per thread the compiler spilled them, and the spill store waited
    // for the global load -- 13 % of the softmax threads' time).
    uint32_t* pre = sPre + sw * (2 * 3 * 32);
    auto prefetch = [&](int item, int buf) {
      uint32_t* dst = pre + buf * 96;
      const int c = c0 + lane;
      if (c < L && p.key_bias != nullptr) cp_async4(dst + lane, p.key_bias + static_cast<long long>(item / p.heads) * L + c);
      else dst[lane] = __float_as_uint(c < L ? 0.f : -INFINITY);
      if (r < L) cp_async4(dst + 32 + lane, p.lse + static_cast<long long>(item) * L + r);
      else dst[32 + lane] = 0u;
      if constexpr (DROP)
        cp_async4(dst + 64 + lane, p.keep_bits + (static_cast<long long>(item) * L + min(r, L - 1)) * 16 + qtr * 4);
      cp_async_commit();
    };
    if (static_cast<int>(blockIdx.x) < n_items) prefetch(blockIdx.x, 0);
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      uint32_t* cur = pre + (it & 1) * 96;
      cp_async_wait_all();
      const float lse2 = __uint_as_float(cur[32 + lane]) * LOG2E;
      uint32_t keep32 = 0xffffffffu;  // DROP: bit c = key column c0 + c of this row survived the forward dropout
      if constexpr (DROP) keep32 = cur[64 + lane];
      __syncwarp();
      wb[lane] = __uint_as_float(cur[lane]) * LOG2E;
      __syncwarp();
      if (item + static_cast<int>(gridDim.x) < n_items) prefetch(item + gridDim.x, (it + 1) & 1);
      mbar_wait(sdp_full, ph);
      tc_fence_after();
      // ---- S and dP of this thread's 32 columns: both loads in flight together, kept in registers to the end
      uint32_t sv[32], dv[32];
      {
        // (the TMEM address is re-derived from shared memory per item: kept in a register across the item it was
        // spilled to local memory, and with ~10 KB of L1 next to 217 KB of shared memory the reload went to L2)
        const uint32_t ta = *reinterpret_cast<volatile uint32_t*>(tmem_ptr) + (static_cast<uint32_t>(quad * 32) << 16) + c0;
        tmem_ld_32x32(ta, sv);
        tmem_ld_32x32(ta + 128, dv);
      }
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_empty);  // S / dP columns may take item i+1 as soon as every warp has its copy
      // ---- P (fp32, in registers as packed pairs) and this quarter's share of delta.  Nothing is written to shared
      // memory yet: S / dP of this item may have been formed BEFORE dV of the previous item was issued (the MMA
      // thread's early issue), so the P buffer may still be in use.
      f32x2 pp[16], dd[16];  // P (undropped) and dP (DROP: masked) of columns c0 + 2 i, c0 + 2 i + 1
      f32x2 dacc0 = pk2(0.f), dacc1 = pk2(0.f);
      {
        const f32x2 sl2p = pk2(sl2), nlse = pk2(-lse2);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a0, a1;
          upk2(add2(fma2(pk2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), sl2p, pk2(wb[2 * i], wb[2 * i + 1])), nlse),
               a0, a1);
          // masked keys carry bias = -inf -> P = 0; rows beyond seq_len are zeroed explicitly
          float p0 = fast_ex2(a0), p1 = fast_ex2(a1);
          p0 = (r < L) ? p0 : 0.f;
          p1 = (r < L) ? p1 : 0.f;
          uint32_t d0 = dv[2 * i], d1 = dv[2 * i + 1];
          if constexpr (DROP) {
            // forward used mask . P / (1 - p) and dP = mask / (1 - p) . (dO V^T): both masks are applied as selects,
            // the two 1 / (1 - p) factors are folded into the dV epilogue and the dS scale
            d0 = ((keep32 >> (2 * i)) & 1u) ? d0 : 0u;
            d1 = ((keep32 >> (2 * i + 1)) & 1u) ? d1 : 0u;
          }
          pp[i] = pk2(p0, p1);
          dd[i] = pk2(__uint_as_float(d0), __uint_as_float(d1));
          if (i & 1) dacc1 = fma2(pp[i], dd[i], dacc1);
          else dacc0 = fma2(pp[i], dd[i], dacc0);
        }
      }
      {
        float t0, t1;
        upk2(add2(dacc0, dacc1), t0, t1);
        sDelta[qtr * ATT_T + r] = t0 + t1;
      }
      if (it > 0) mbar_wait(dv_done, (it - 1) & 1);  // dV of the previous item has read the P buffer
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float pv[8];  // what multiplied V in the forward
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          upk2(pp[g * 4 + j], pv[2 * j], pv[2 * j + 1]);
          if constexpr (DROP) {
            pv[2 * j] = ((keep32 >> (g * 8 + 2 * j)) & 1u) ? pv[2 * j] : 0.f;
            pv[2 * j + 1] = ((keep32 >> (g * 8 + 2 * j + 1)) & 1u) ? pv[2 * j + 1] : 0.f;
          }
        }
        *reinterpret_cast<uint4*>(sP + swz_off(r, c0 + g * 8)) = pack8(pv);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);  // dV = P^T dO may start while dS is being formed
      named_bar_sync(1 + quad, 128);       // the four warps that share these 32 rows
      const float delta = (sDelta[r] + sDelta[ATT_T + r]) + (sDelta[2 * ATT_T + r] + sDelta[3 * ATT_T + r]);
      // ---- dS = P (dP - delta) scale
      if (it > 0) mbar_wait(out_full, (it - 1) & 1);  // dK / dQ of the previous item have read the dS buffer
      {
        const f32x2 scp = pk2(ds_scale), nds = pk2(-delta * ds_scale);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float ds[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) upk2(mul2(pp[g * 4 + j], fma2(dd[g * 4 + j], scp, nds)), ds[2 * j], ds[2 * j + 1]);
          *reinterpret_cast<uint4*>(sdS + swz_off(r, c0 + g * 8)) = pack8(ds);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);  // dS is in shared memory
      named_bar_sync(1 + quad, 128);  // sDelta of this item has been read by all four warps before it is rewritten
    }
  } else {
    // ===================== epilogue threads: dV, dK, dQ rows -> dQKV + bias gradient =====================
    // tcgen05.ld hands a thread one ROW; 32 lanes storing 16 B each would touch 32 different 128-byte lines per
    // instruction.  Each warp therefore transposes its 32 x 32 fp16 chunk through a private, XOR-swizzled 2 KB
    // tile: afterwards 4 neighbouring lanes own one row's 64 bytes and a store instruction covers 8 whole rows.
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // key row (dV, dK) == query row (dQ)
    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    uint8_t* tile = sEpi + quad * 2048;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int seq = item / p.heads, h = item % p.heads;
      const int row0 = seq * L;
      mbar_wait(out_full, ph);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 3; ++t) {  // t: 0 = dV, 1 = dK, 2 = dQ
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(trow + 256 + t * 64 + c * 32, v);
          tc_wait_ld();
          if (t == 2 && c == 1) {  // last read of this item's accumulators: hand them back before the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(out_empty);
          }
          float f[32];
          const float osc = (DROP && t == 0) ? p.drop.scale : 1.f;  // dV = (mask . P)^T dO / (1 - p)
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * osc;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) e[j] = f[g * 8 + j];
            *reinterpret_cast<uint4*>(tile + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = pack8(e);
          }
          __syncwarp();
          float cs[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) cs[j] = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = (lane >> 2) + 8 * i;  // row of the chunk
            const int pc = lane & 3;             // 16-byte piece of that row
            const uint4 q = *reinterpret_cast<const uint4*>(tile + rr * 64 + ((pc ^ ((rr >> 1) & 3)) << 4));
            const int grow_i = quad * 32 + rr;
            if (grow_i < L)
              *reinterpret_cast<uint4*>(p.dqkv + static_cast<long long>(row0 + grow_i) * (3 * p.hidden) +
                                        (2 - t) * p.hidden + h * ATT_D + c * 32 + pc * 8) = q;
            const __half2* qh = reinterpret_cast<const __half2*>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // rows beyond seq_len are exact zeros: no predicate needed
              const float2 x = __half22float2(qh[j]);
              cs[2 * j] += x.x;
              cs[2 * j + 1] += x.y;
            }
          }
          if (p.dbias != nullptr) {
            // bias gradient = column sums of the stored values: the 8 lanes that share a piece hold the same 8
            // columns for different rows; a halving butterfly over lane bits 4,3,2 leaves one column per lane
            float a4[4], a2[2];
            {
              const bool up = (lane & 16) != 0;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                a4[j] = (up ? cs[j + 4] : cs[j]) + __shfl_xor_sync(0xffffffffu, up ? cs[j] : cs[j + 4], 16);
            }
            {
              const bool up = (lane & 8) != 0;
#pragma unroll
              for (int j = 0; j < 2; ++j)
                a2[j] = (up ? a4[j + 2] : a4[j]) + __shfl_xor_sync(0xffffffffu, up ? a4[j] : a4[j + 2], 8);
            }
            const bool up = (lane & 4) != 0;
            const float tot = (up ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, up ? a2[0] : a2[1], 4);
            const int col = (lane & 3) * 8 + ((lane & 16) ? 4 : 0) + ((lane & 8) ? 2 : 0) + ((lane & 4) ? 1 : 0);
            atomicAdd(p.dbias + (2 - t) * p.hidden + h * ATT_D + c * 32 + col, tot * p.dbias_scale);
          }
          __syncwarp();  // the transposition tile is rewritten by the next chunk
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ===================================================================================== seq_len > 128
// Longer sequences (BERT-large at L = 256 for COCO pre-training, up to the 512-position table) are tiled
// 128 x 128.  Forward: one CTA per (sequence, head, query tile), two passes over the key tiles -- pass 1
// takes the exact row maxima (S = Q K_j^T only), pass 2 recomputes S_j, exponentiates against the final
// maximum and accumulates O += P_j V_j in TMEM, so no running rescale of O is needed.  Backward: one CTA
// per (sequence, head, key tile) loops over the query tiles; dV_j / dK_j accumulate in TMEM across the
// loop, dQ_i contributions of the different key tiles meet in an fp32 workspace (red.add) that
// dq_convert_kernel folds into the packed fp16 dQKV afterwards.
constexpr int ATT_FWDM_SMEM = 5 * ATT_TILE_BYTES + 512 + 128 + 1024;  // Q K V | P (2 tiles)

template <bool DROP>
__global__ void __launch_bounds__(128, 2)
fmha_fwd_multi_kernel(const __grid_constant__ CUtensorMap tma_qkv, const AttParams p, const int q_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;
  uint8_t* sV = smem + 2 * ATT_TILE_BYTES;
  uint8_t* sP = smem + 3 * ATT_TILE_BYTES;
  float* sBias = reinterpret_cast<float*>(smem + 5 * ATT_TILE_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 5 * ATT_TILE_BYTES + 512);  // 0: Q, 1: K(/V), 2: S, 3: O
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = blockIdx.x % q_tiles;
  const int sh = blockIdx.x / q_tiles;
  const int seq = sh / p.heads, h = sh % p.heads;
  const int L = p.seq_len;
  const int row0 = seq * L;
  const int q0 = qt * ATT_T;
  const int n_kv = (L + ATT_T - 1) / ATT_T;

  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t tmem_o = tmem + 128;
  const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK), va = smem_u32(sV), pa = smem_u32(sP);
  constexpr uint32_t idesc_s = make_idesc_f16(ATT_T, ATT_T, 0, 0);
  constexpr uint32_t idesc_o = make_idesc_f16(ATT_T, ATT_D, 0, 1);
  const int r = tid;
  const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const float sl2 = p.scale * LOG2E;
  uint32_t ph_k = 0, ph_s = 0, ph_o = 0;
  const float drop_scale = DROP ? p.drop.scale : 1.f;
  const uint8_t* bits_row = DROP ? p.keep_bits + (static_cast<long long>(sh) * L + min(q0 + r, L - 1)) * p.bits_stride : nullptr;

  if (tid == 0) {
    mbar_expect_tx(&bar[0], ATT_TILE_BYTES);
    tma_load_2d(sQ, &tma_qkv, &bar[0], h * ATT_D, row0 + q0);
  }
  float mx = -INFINITY, sum = 0.f;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
    for (int j = 0; j < n_kv; ++j) {
      const int k0 = j * ATT_T;
      // all threads are past their reads of the previous tile's bias / S / P here (loop-end barrier)
      {
        const int c = k0 + tid;
        float b = -INFINITY;
        if (c < L) b = p.key_bias ? p.key_bias[static_cast<long long>(seq) * L + c] * LOG2E : 0.f;
        sBias[tid] = b;
      }
      if (tid == 0) {
        mbar_expect_tx(&bar[1], (pass == 1 ? 2 : 1) * ATT_TILE_BYTES);
        tma_load_2d(sK, &tma_qkv, &bar[1], p.hidden + h * ATT_D, row0 + k0);
        if (pass == 1) tma_load_2d(sV, &tma_qkv, &bar[1], 2 * p.hidden + h * ATT_D, row0 + k0);
        if (pass == 0 && j == 0) mbar_wait(&bar[0], 0);
        mbar_wait(&bar[1], ph_k);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          tc_mma_f16(tmem, make_smem_desc(qa + k * 32, 16, 1024), make_smem_desc(ka + k * 32, 16, 1024), idesc_s,
                     k > 0);
        tc_commit(&bar[2]);
      }
      ph_k ^= 1;
      __syncthreads();  // sBias visible
      mbar_wait(&bar[2], ph_s);
      ph_s ^= 1;
      tc_fence_after();
      __syncwarp();
      if (pass == 0) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(trow + c * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int t = 0; t < 32; ++t) mx = fmaxf(mx, fmaf(__uint_as_float(v[t]), sl2, sBias[c * 32 + t]));
        }
      } else {
        const float m_use = (mx == -INFINITY) ? 0.f : mx;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(trow + c * 32, v);
          uint32_t kw = 0xffffffffu;
          if constexpr (DROP) kw = __ldg(reinterpret_cast<const uint32_t*>(bits_row + (k0 >> 3) + c * 4));
          tc_wait_ld();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float e[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              e[t] = exp2f(fmaf(__uint_as_float(v[g * 8 + t]), sl2, sBias[c * 32 + g * 8 + t]) - m_use);
              sum += e[t];
            }
            if constexpr (DROP) {
              const uint32_t keep = kw >> (8 * g);
#pragma unroll
              for (int t = 0; t < 8; ++t) e[t] = ((keep >> t) & 1u) ? e[t] : 0.f;
            }
            *reinterpret_cast<uint4*>(sP + swz_off(r, c * 32 + g * 8)) = pack8(e);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < ATT_T / 16; ++k)
            tc_mma_f16(tmem_o, make_smem_desc(pa + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32, 16, 1024),
                       make_smem_desc(va + k * 2048, 8192, 1024), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
          tc_commit(&bar[3]);
        }
        __syncwarp();
        mbar_wait(&bar[3], ph_o);  // P, K, V consumed: buffers reusable; O advanced
        ph_o ^= 1;
        tc_fence_after();
      }
      tc_fence_before();
      __syncthreads();  // everyone done with S / sBias of this tile before the next one is produced
      tc_fence_after();
    }
  }
  const int qrow = q0 + r;
  if (qrow < L && p.lse) {
    const float m_use = (mx == -INFINITY) ? 0.f : mx;
    p.lse[(static_cast<long long>(seq) * p.heads + h) * L + qrow] = (m_use + log2f(sum)) / LOG2E;
  }
  const float inv = (sum > 0.f ? 1.f / sum : 0.f) * drop_scale;
  __half* orow = p.out + static_cast<long long>(row0 + qrow) * p.hidden + h * ATT_D;
  const uint32_t trow_o = tmem_o + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(trow_o + c * 32, v);
    tc_wait_ld();
    if (qrow < L) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float e[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) e[t] = __uint_as_float(v[g * 8 + t]) * inv;
        *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = pack8(e);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// smem: K V (fixed) | Q dO (per query tile) | P then dS (32 KB)
constexpr int ATT_BWDM_SMEM = 6 * ATT_TILE_BYTES + 1024 + 128 + 1024;

template <bool DROP>
__global__ void __launch_bounds__(256, 1)
fmha_bwd_multi_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                      const AttParams p, float* __restrict__ dq_ws, const int kv_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sK = smem;
  uint8_t* sV = smem + ATT_TILE_BYTES;
  uint8_t* sQ = smem + 2 * ATT_TILE_BYTES;
  uint8_t* sdO = smem + 3 * ATT_TILE_BYTES;
  uint8_t* sP = smem + 4 * ATT_TILE_BYTES;
  float* sBias = reinterpret_cast<float*>(smem + 6 * ATT_TILE_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 6 * ATT_TILE_BYTES + 1024);  // 0: KV, 1: Q dO, 2: S dP, 3: dV, 4: dK dQ
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int jt = blockIdx.x % kv_tiles;
  const int sh = blockIdx.x / kv_tiles;
  const int seq = sh / p.heads, h = sh % p.heads;
  const int L = p.seq_len;
  const int row0 = seq * L;
  const int k0 = jt * ATT_T;
  const int q_tiles = kv_tiles;

  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
#pragma unroll
    for (int i = 0; i < 5; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (tid < ATT_T) {
    const int c = k0 + tid;
    float b = -INFINITY;
    if (c < L) b = p.key_bias ? p.key_bias[static_cast<long long>(seq) * L + c] * LOG2E : 0.f;
    sBias[tid] = b;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  // TMEM columns: S [0,128) dP [128,256) dV [256,320) dK [320,384) dQ [384,448)
  const uint32_t t_s = tmem, t_dp = tmem + 128, t_dv = tmem + 256, t_dk = tmem + 320, t_dq = tmem + 384;
  const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK), va = smem_u32(sV), da = smem_u32(sdO), pa = smem_u32(sP);
  constexpr uint32_t idesc_s = make_idesc_f16(ATT_T, ATT_T, 0, 0);
  constexpr uint32_t idesc_tt = make_idesc_f16(ATT_T, ATT_D, 1, 1);
  constexpr uint32_t idesc_nt = make_idesc_f16(ATT_T, ATT_D, 0, 1);
  const int rl = (warp & 3) * 32 + lane;  // TMEM lane
  const int half = warp >> 2;
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl2 = p.scale * LOG2E;
  const float drop_scale = DROP ? p.drop.scale : 1.f;

  if (tid == 0) {
    mbar_expect_tx(&bar[0], 2 * ATT_TILE_BYTES);
    tma_load_2d(sK, &tma_qkv, &bar[0], p.hidden + h * ATT_D, row0 + k0);
    tma_load_2d(sV, &tma_qkv, &bar[0], 2 * p.hidden + h * ATT_D, row0 + k0);
  }
  uint32_t ph = 0;
#pragma unroll 1
  for (int it = 0; it < q_tiles; ++it) {
    const int q0 = it * ATT_T;
    if (tid == 0) {
      mbar_expect_tx(&bar[1], 2 * ATT_TILE_BYTES);
      tma_load_2d(sQ, &tma_qkv, &bar[1], h * ATT_D, row0 + q0);
      tma_load_2d(sdO, &tma_do, &bar[1], h * ATT_D, row0 + q0);
      if (it == 0) mbar_wait(&bar[0], 0);
      mbar_wait(&bar[1], ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < ATT_D / 16; ++k)
        tc_mma_f16(t_s, make_smem_desc(qa + k * 32, 16, 1024), make_smem_desc(ka + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
      for (int k = 0; k < ATT_D / 16; ++k)
        tc_mma_f16(t_dp, make_smem_desc(da + k * 32, 16, 1024), make_smem_desc(va + k * 32, 16, 1024), idesc_s, k > 0);
      tc_commit(&bar[2]);
    }
    const int qrow = q0 + rl;
    float delta = 0.f, lse2 = 0.f;
    if (qrow < L) {
      const long long off = static_cast<long long>(row0 + qrow) * p.hidden + h * ATT_D;
      const uint4* po = reinterpret_cast<const uint4*>(p.o + off);
      const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + off);
#pragma unroll
      for (int i = 0; i < 8; ++i) delta += dot8(__ldg(po + i), __ldg(pd + i));
      lse2 = p.lse[(static_cast<long long>(seq) * p.heads + h) * L + qrow] * LOG2E;
    }
    __syncwarp();
    mbar_wait(&bar[2], ph);
    tc_fence_after();
    __syncwarp();

    uint4 ds_keep[8];
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c0 = half * 64 + cc * 32;
      uint32_t s[32], d[32];
      tmem_ld_32x32(t_s + lane_off + c0, s);
      tmem_ld_32x32(t_dp + lane_off + c0, d);
      uint32_t kw = 0xffffffffu;
      if constexpr (DROP)
        kw = __ldg(reinterpret_cast<const uint32_t*>(p.keep_bits + (static_cast<long long>(sh) * L + min(qrow, L - 1)) * p.bits_stride +
                                                     ((k0 + c0) >> 3)));
      tc_wait_ld();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float pv[8], ds[8];
        const uint32_t keep = kw >> (8 * g);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int c = c0 + g * 8 + t;
          const bool ok = (qrow < L) && (k0 + c < L);
          const float pe = ok ? exp2f(fmaf(__uint_as_float(s[g * 8 + t]), sl2, sBias[c]) - lse2) : 0.f;
          const float mj = DROP ? (((keep >> t) & 1u) ? drop_scale : 0.f) : 1.f;
          pv[t] = pe * mj;
          ds[t] = ok ? pe * (mj * __uint_as_float(d[g * 8 + t]) - delta) * p.scale : 0.f;
        }
        *reinterpret_cast<uint4*>(sP + swz_off(rl, c0 + g * 8)) = pack8(pv);
        ds_keep[cc * 4 + g] = pack8(ds);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < ATT_T / 16; ++k)  // dV_j += P^T dO_i
        tc_mma_f16(t_dv, make_smem_desc(pa + k * 2048, ATT_TILE_BYTES, 1024), make_smem_desc(da + k * 2048, 8192, 1024),
                   idesc_tt, (it > 0 || k > 0) ? 1u : 0u);
      tc_commit(&bar[3]);
    }
    __syncwarp();
    mbar_wait(&bar[3], ph);
    tc_fence_after();
    __syncwarp();
#pragma unroll
    for (int cc = 0; cc < 2; ++cc)
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(sP + swz_off(rl, half * 64 + cc * 32 + g * 8)) = ds_keep[cc * 4 + g];
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < ATT_T / 16; ++k)  // dK_j += dS^T Q_i
        tc_mma_f16(t_dk, make_smem_desc(pa + k * 2048, ATT_TILE_BYTES, 1024), make_smem_desc(qa + k * 2048, 8192, 1024),
                   idesc_tt, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < ATT_T / 16; ++k)  // dQ_i (this key tile's share) = dS K_j
        tc_mma_f16(t_dq, make_smem_desc(pa + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32, 16, 1024),
                   make_smem_desc(ka + k * 2048, 8192, 1024), idesc_nt, k > 0);
      tc_commit(&bar[4]);
    }
    __syncwarp();
    mbar_wait(&bar[4], ph);
    tc_fence_after();
    __syncwarp();
    {  // dQ share -> fp32 workspace
      uint32_t v[32];
      tmem_ld_32x32(t_dq + lane_off + half * 32, v);
      tc_wait_ld();
      if (qrow < L) {
        float* dst = dq_ws + static_cast<long long>(row0 + qrow) * p.hidden + h * ATT_D + half * 32;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          atomicAdd(reinterpret_cast<float4*>(dst + g * 4),
                    make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                                __uint_as_float(v[4 * g + 3])));
      }
    }
    ph ^= 1;
    tc_fence_before();
    __syncthreads();  // Q / dO / dS buffers and the S / dP / dQ columns are free for the next query tile
    tc_fence_after();
  }
  // dV_j, dK_j rows = key rows k0 + rl
  const int krow = k0 + rl;
  __half* grow = p.dqkv + static_cast<long long>(row0 + krow) * (3 * p.hidden) + h * ATT_D + half * 32;
#pragma unroll 1
  for (int t = 0; t < 2; ++t) {  // 0 = dV, 1 = dK
    uint32_t v[32];
    tmem_ld_32x32((t == 0 ? t_dv : t_dk) + lane_off + half * 32, v);
    tc_wait_ld();
    if (krow < L) {
      __half* dst = grow + (2 - t) * p.hidden;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) e[u] = __uint_as_float(v[g * 8 + u]);
        *reinterpret_cast<uint4*>(dst + g * 8) = pack8(e);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// dqkv[:, 0:hidden] = fp16(dq_ws)
__global__ void dq_convert_kernel(const float* __restrict__ ws, __half* __restrict__ dqkv, long long rows, int hidden) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i >= rows * hidden) return;
  const long long r = i / hidden;
  const int c = static_cast<int>(i - r * hidden);
  const float4 a = *reinterpret_cast<const float4*>(ws + i);
  const float4 b = *reinterpret_cast<const float4*>(ws + i + 4);
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  *reinterpret_cast<uint4*>(dqkv + r * 3 * hidden + c) = pack8(v);
}

// keep bits of the attention-probability dropout: byte kg of row (item * L + r) = Philox group (item * L + r) * 64 + kg
// (a thread produces one 32-bit word = 4 groups: four independent Philox chains in flight, one 4-byte store)
__global__ void __launch_bounds__(256)
att_keep_bits_kernel(uint8_t* __restrict__ bits, long long n_rows, int groups_per_row, int stride, const cdr_dropout drop) {
  const DropCtx dc = drop_load(drop);
  const int words_per_row = stride >> 2;
  const long long total = n_rows * words_per_row;
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += step) {
    const long long row = i / words_per_row;
    const int w = static_cast<int>(i - row * words_per_row);
    uint32_t word = 0u;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int kg = 4 * w + b;
      if (kg < groups_per_row)
        word |= drop_keep8(dc, static_cast<uint32_t>(row) * 64u + static_cast<uint32_t>(kg)) << (8 * b);
    }
    *reinterpret_cast<uint32_t*>(bits + row * stride + 4 * w) = word;
  }
}

static int att_bits_stride(int seq_len) { return ((seq_len + 7) / 8 + 15) / 16 * 16; }

static int att_fill_bits(const cdr_attn_args* a, void* stream) {
  const int stride = att_bits_stride(a->seq_len);
  const long long n_rows = static_cast<long long>(a->n_seq) * a->heads * a->seq_len;
  const int gpr = (a->seq_len + 7) / 8;
  long long blocks = (n_rows * (stride / 4) + 255) / 256;
  if (blocks > 32LL * sm_count()) blocks = 32LL * sm_count();
  att_keep_bits_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint8_t*>(a->drop_bits), n_rows, gpr, stride, a->drop);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

static int att_check(const cdr_attn_args* a) {
  CDR_REQUIRE(a != nullptr, "cdr_attn: null args");
  if (a->drop.state != nullptr && a->drop.threshold > 0) {
    CDR_REQUIRE(a->drop.threshold < 65536, "cdr_attn: drop.threshold out of range");
    CDR_REQUIRE(a->drop_bits != nullptr && (reinterpret_cast<uintptr_t>(a->drop_bits) & 15) == 0,
                "cdr_attn: dropout needs drop_bits (16-byte aligned, cdr_attn_dropout_bits_bytes() bytes)");
    CDR_REQUIRE(static_cast<long long>(a->n_seq) * a->heads * a->seq_len * 64 < (1ll << 32),
                "cdr_attn: dropout group index overflows 32 bits");
  }
  CDR_REQUIRE(a->qkv != nullptr, "cdr_attn: null qkv");
  CDR_REQUIRE(a->n_seq > 0 && a->seq_len > 0 && a->heads > 0, "cdr_attn: empty problem");
  CDR_REQUIRE(a->seq_len <= 512, "cdr_attn: seq_len %d > 512 not supported", a->seq_len);
  CDR_REQUIRE(a->head_dim == ATT_D, "cdr_attn: head_dim must be 64 (got %d)", a->head_dim);
  return CDR_OK;
}

}  // namespace cdr

using namespace cdr;

extern "C" {

size_t cdr_attn_dropout_bits_bytes(int32_t n_seq, int32_t heads, int32_t seq_len) {
  if (n_seq <= 0 || heads <= 0 || seq_len <= 0) return 0;
  return static_cast<size_t>(n_seq) * heads * seq_len * att_bits_stride(seq_len);
}

int cdr_attn_fwd(const cdr_attn_args* a, void* stream) {
  if (int rc = att_check(a)) return rc;
  CDR_REQUIRE(a->out != nullptr && a->lse != nullptr, "cdr_attn_fwd: null output");
  const int hidden = a->heads * ATT_D;
  const long long T = static_cast<long long>(a->n_seq) * a->seq_len;
  CUtensorMap tq;
  if (int rc = make_tma_2d_f16(&tq, a->qkv, 3 * hidden, T, 3 * hidden, ATT_D, ATT_T)) return rc;
  AttParams p{};
  p.n_seq = a->n_seq; p.seq_len = a->seq_len; p.heads = a->heads; p.hidden = hidden;
  p.key_bias = a->key_bias;
  p.scale = a->scale;
  p.out = static_cast<__half*>(a->out);
  p.lse = a->lse;
  p.drop = a->drop;
  const bool drop = drop_on(a->drop);
  if (drop) {  // pre-pass: one keep bit per attention probability (read again by cdr_attn_bwd)
    p.keep_bits = static_cast<const uint8_t*>(a->drop_bits);
    p.bits_stride = att_bits_stride(a->seq_len);
    if (!a->drop_bits_ready) {
      if (int rc = att_fill_bits(a, stream)) return rc;
    }
  }
  static bool configured = false;
  if (!configured) {
    CDR_CUDA(cudaFuncSetAttribute(fmha_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_FWD_SMEM));
    CDR_CUDA(cudaFuncSetAttribute(fmha_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_FWD_SMEM));
    CDR_CUDA(cudaFuncSetAttribute(fmha_fwd_multi_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_FWDM_SMEM));
    CDR_CUDA(cudaFuncSetAttribute(fmha_fwd_multi_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_FWDM_SMEM));
    configured = true;
  }
  if (a->seq_len > ATT_T) {
    const int q_tiles = (a->seq_len + ATT_T - 1) / ATT_T;
    auto kern = drop ? fmha_fwd_multi_kernel<true> : fmha_fwd_multi_kernel<false>;
    kern<<<a->n_seq * a->heads * q_tiles, 128, ATT_FWDM_SMEM, static_cast<cudaStream_t>(stream)>>>(tq, p, q_tiles);
    CDR_LAUNCH_CHECK();
    return CDR_OK;
  }
  const int n_items = a->n_seq * a->heads;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  CDR_CUDA(launch_pdl(drop ? fmha_fwd_kernel<true> : fmha_fwd_kernel<false>, dim3(grid), dim3(ATT_FWD_THREADS),
                      ATT_FWD_SMEM, static_cast<cudaStream_t>(stream), tq, p, n_items));
  return CDR_OK;
}

int cdr_attn_dropout_bits_fill(const cdr_attn_args* a, void* stream) {
  CDR_REQUIRE(a != nullptr, "cdr_attn_dropout_bits_fill: null args");
  CDR_REQUIRE(drop_on(a->drop) && a->drop.threshold < 65536, "cdr_attn_dropout_bits_fill: no dropout configured");
  CDR_REQUIRE(a->drop_bits != nullptr && (reinterpret_cast<uintptr_t>(a->drop_bits) & 15) == 0,
              "cdr_attn_dropout_bits_fill: drop_bits must be 16-byte aligned");
  CDR_REQUIRE(a->n_seq > 0 && a->seq_len > 0 && a->heads > 0, "cdr_attn_dropout_bits_fill: empty problem");
  CDR_REQUIRE(static_cast<long long>(a->n_seq) * a->heads * a->seq_len * 64 < (1ll << 32),
              "cdr_attn_dropout_bits_fill: dropout group index overflows 32 bits");
  return att_fill_bits(a, stream);
}

int cdr_attn_bwd(const cdr_attn_args* a, void* stream) {
  if (int rc = att_check(a)) return rc;
  CDR_REQUIRE(a->out != nullptr && a->lse != nullptr && a->d_out != nullptr && a->dqkv != nullptr,
              "cdr_attn_bwd: null pointer");
  const int hidden = a->heads * ATT_D;
  const long long T = static_cast<long long>(a->n_seq) * a->seq_len;
  CUtensorMap tq, td;
  if (int rc = make_tma_2d_f16(&tq, a->qkv, 3 * hidden, T, 3 * hidden, ATT_D, ATT_T)) return rc;
  if (int rc = make_tma_2d_f16(&td, a->d_out, hidden, T, hidden, ATT_D, ATT_T)) return rc;
  AttParams p{};
  p.n_seq = a->n_seq; p.seq_len = a->seq_len; p.heads = a->heads; p.hidden = hidden;
  p.key_bias = a->key_bias;
  p.scale = a->scale;
  p.lse = a->lse;
  p.o = static_cast<const __half*>(a->out);
  p.d_o = static_cast<const __half*>(a->d_out);
  p.dqkv = static_cast<__half*>(a->dqkv);
  p.drop = a->drop;
  const bool drop = drop_on(a->drop);
  p.keep_bits = static_cast<const uint8_t*>(a->drop_bits);
  p.bits_stride = att_bits_stride(a->seq_len);
  static bool configured = false;
  if (!configured) {
    CDR_CUDA(cudaFuncSetAttribute(fmha_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_BWD_SMEM));
    CDR_CUDA(cudaFuncSetAttribute(fmha_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_BWD_SMEM));
    CDR_CUDA(cudaFuncSetAttribute(fmha_bwd_multi_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_BWDM_SMEM));
    CDR_CUDA(cudaFuncSetAttribute(fmha_bwd_multi_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_BWDM_SMEM));
    configured = true;
  }
  if (a->seq_len > ATT_T) {
    CDR_REQUIRE(a->dq_workspace != nullptr, "cdr_attn_bwd: seq_len > %d needs dq_workspace (fp32 [T, hidden])", ATT_T);
    CDR_REQUIRE(a->dbias_qkv == nullptr, "cdr_attn_bwd: the fused QKV bias gradient needs seq_len <= %d", ATT_T);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int kv_tiles = (a->seq_len + ATT_T - 1) / ATT_T;
    CDR_CUDA(cudaMemsetAsync(a->dq_workspace, 0, sizeof(float) * static_cast<size_t>(T) * hidden, st));
    auto kern = drop ? fmha_bwd_multi_kernel<true> : fmha_bwd_multi_kernel<false>;
    kern<<<a->n_seq * a->heads * kv_tiles, 256, ATT_BWDM_SMEM, st>>>(tq, td, p, a->dq_workspace, kv_tiles);
    CDR_LAUNCH_CHECK();
    const long long n8 = (static_cast<long long>(T) * hidden + 7) / 8;
    dq_convert_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, st>>>(a->dq_workspace, p.dqkv, T, hidden);
    CDR_LAUNCH_CHECK();
    return CDR_OK;
  }
  p.dbias = a->dbias_qkv;
  p.dbias_scale = a->dbias_scale;
  const int n_items = a->n_seq * a->heads;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  CDR_CUDA(launch_pdl(drop ? fmha_bwd_kernel<true> : fmha_bwd_kernel<false>, dim3(grid), dim3(ATT_BWD_THREADS),
                      ATT_BWD_SMEM, static_cast<cudaStream_t>(stream), tq, td, p, n_items));
  return CDR_OK;
}

}  // extern "C"
