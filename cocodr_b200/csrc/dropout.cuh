// Counter-based dropout shared by every kernel that drops activations (HF BERT's nn.Dropout sites reached through
// self.bert(...) in ANCE/model/models.py:226 with the model in train(); COCO/modeling.py:216-220 for the c_head layers).
//
// A mask is never stored: forward and backward regenerate it from (seed, offset, site, element index) with
// Philox4x32-10.  One Philox call serves a GROUP of 8 consecutive elements of a row:
//     counter = (group index, site, offset lo, offset hi)      key = (seed lo, seed hi)
//     output  = 4 x u32 = 8 x u16;  element j of the group keeps its value iff u16_j >= threshold
// with threshold = round(p * 65536), so P(keep) = 1 - threshold / 65536; kept values are multiplied by 1 / (1 - p).
// Group index of element (row m, column n) of a [rows, cols] tensor: (m * row_mul) * (cols / 8) + n / 8 (row_mul lets
// the [CLS]-rows-only last layer address the same masks as the full layer); attention probabilities use
// ((seq * heads + head) * seq_len + query row) * 64 + key / 8.  oracle/dropout_ref.py restates exactly this.
#pragma once
#include <stdint.h>

#include "../../include/cocodr_b200.h"

namespace cdr {

struct DropCtx {
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
  uint32_t site, thr;
  float scale;
  int row_mul;
  uint8_t* keep_bits;  // optional [rows, cols / 8] stored masks (GEMM epilogue -> LayerNorm backward)
};

// state = device [seed, offset]; must be read after pdl_wait() (a preceding kernel of the stream may have written it)
__device__ __forceinline__ DropCtx drop_load(const cdr_dropout& d) {
  DropCtx c;
  const unsigned long long seed = d.state[0], off = d.state[1];
  c.seed_lo = static_cast<uint32_t>(seed);
  c.seed_hi = static_cast<uint32_t>(seed >> 32);
  c.off_lo = static_cast<uint32_t>(off);
  c.off_hi = static_cast<uint32_t>(off >> 32);
  c.site = d.site;
  c.thr = d.threshold;
  c.scale = d.scale;
  c.row_mul = d.row_mul > 0 ? d.row_mul : 1;
  c.keep_bits = static_cast<uint8_t*>(d.keep_bits);
  return c;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = static_cast<unsigned long long>(0xD2511F53u) * c0;
    const unsigned long long p1 = static_cast<unsigned long long>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    c1 = static_cast<uint32_t>(p1);
    c3 = static_cast<uint32_t>(p0);
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// bit j of the result: element j of group `gidx` is KEPT
__device__ __forceinline__ uint32_t drop_keep8(const DropCtx& c, uint32_t gidx) {
  const uint4 r = philox4x32_10(gidx, c.site, c.off_lo, c.off_hi, c.seed_lo, c.seed_hi);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t m = 0u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m |= ((w[j] & 0xffffu) >= c.thr ? 1u : 0u) << (2 * j);
    m |= ((w[j] >> 16) >= c.thr ? 1u : 0u) << (2 * j + 1);
  }
  return m;
}

// group index of 8 consecutive columns starting at n (n % 8 == 0) of row m in a [*, cols] tensor
__device__ __forceinline__ uint32_t drop_group(const DropCtx& c, long long m, int n, int cols) {
  return static_cast<uint32_t>(m * c.row_mul * (cols >> 3) + (n >> 3));
}

__device__ __forceinline__ void drop_apply8(const DropCtx& c, uint32_t keep, float (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * c.scale : 0.f;
}

// host-side validation shared by the entry points
inline bool drop_on(const cdr_dropout& d) { return d.state != nullptr && d.threshold > 0; }

}  // namespace cdr
