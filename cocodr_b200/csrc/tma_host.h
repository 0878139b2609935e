// Host helper: encode a 2-D fp16 tiled tensor map with 128B swizzle (driver API via runtime entry point,
// so the library never links libcuda at build time).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace cdr {
// Global view {inner (contiguous), outer}, row stride ld_elems; box {box_inner (<= 64), box_outer (<= 256)}.
int make_tma_2d_f16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                    uint32_t box_inner, uint32_t box_outer);
// Same for fp32 elements (box_inner <= 32: one 128-byte swizzle row).
int make_tma_2d_f32(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                    uint32_t box_inner, uint32_t box_outer);
}  // namespace cdr
