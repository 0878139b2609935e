// Library core: version, thread-local error text, device check, driver entry point for TMA maps.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>

#include "cdr_common.cuh"
#include "tma_host.h"

namespace cdr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// SMs the persistent kernels size their grids for.  cdr_set_sm_budget(n) (or CDR_SM_BUDGET in the environment)
// leaves the remaining SMs to concurrent communication kernels: a persistent GEMM CTA takes a whole SM's shared
// memory, so an NCCL CTA squatting on one SM would otherwise push that GEMM CTA into a second wave.
static int g_sm_budget = -1;

int sm_count() {
  if (g_sm_budget < 0) {
    const char* e = getenv("CDR_SM_BUDGET");
    g_sm_budget = e ? atoi(e) : 0;
  }
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  int n = cached[dev];
  if (g_sm_budget > 0 && g_sm_budget < n) n = g_sm_budget & ~1;  // even: CTA pairs
  return n;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CDR_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int make_tma_2d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, uint64_t esize, uint64_t inner,
                       uint64_t outer, uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return CDR_ECUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld_elems * esize) % 16 != 0) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row stride (base=%p ld=%llu)", base,
              (unsigned long long)ld_elems);
    return CDR_EINVAL;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * esize};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_elems, box_inner, box_outer);
    return CDR_ECUDA;
  }
  return CDR_OK;
}

int make_tma_2d_f16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                    uint32_t box_inner, uint32_t box_outer) {
  return make_tma_2d(map, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, inner, outer, ld_elems, box_inner, box_outer);
}

int make_tma_2d_f32(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                    uint32_t box_inner, uint32_t box_outer) {
  return make_tma_2d(map, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, inner, outer, ld_elems, box_inner, box_outer);
}

}  // namespace cdr

extern "C" {

int cdr_version(void) { return 100; }

int cdr_set_sm_budget(int32_t n) {
  cdr::g_sm_budget = n > 0 ? n : 0;
  return CDR_OK;
}

const char* cdr_last_error(void) { return cdr::g_err; }

int cdr_device_check(void) {
  int dev = 0;
  CDR_CUDA(cudaGetDevice(&dev));
  int major = 0;
  CDR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    cdr::set_error("cocodr_b200 needs an sm_100 device, found compute capability %d.x", major);
    return CDR_EARCH;
  }
  return CDR_OK;
}

}  // extern "C"
