// Common helpers for the cocodr_b200 sm_100a kernels: error plumbing for the C ABI and thin
// inline-PTX wrappers (mbarrier, TMA, tcgen05/TMEM).  No torch headers anywhere in csrc/.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cocodr_b200.h"

namespace cdr {

void set_error(const char* fmt, ...);
int sm_count();

#define CDR_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      cdr::set_error(__VA_ARGS__);        \
      return CDR_EINVAL;                  \
    }                                     \
  } while (0)

#define CDR_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      cdr::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CDR_ECUDA;                                                                  \
    }                                                                                    \
  } while (0)

#define CDR_LAUNCH_CHECK()                                                                 \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      cdr::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return CDR_ECUDA;                                                                    \
    }                                                                                      \
  } while (0)

// ------------------------------------------------------------------------------------------
// device-side PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Programmatic dependent launch: kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization, so
// grid N+1 may be scheduled (and run its prologue: barrier init, TMEM allocation, tensor-map prefetch) while grid N
// drains.  pdl_wait() blocks until the preceding grid has completed and its memory is visible -- NOTHING that a
// predecessor wrote (or still reads) may be touched before it; pdl_launch() lets the successor start its own prologue.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();  // core.cu: CDR_PDL=0 in the environment turns the launch attribute off

// <<<grid, block, smem, stream>>> with the programmatic-serialization attribute; the kernel must call pdl_wait()
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// First 1024-byte aligned address of the dynamic shared memory (SWIZZLE_128B tiles need it).  The offset is
// computed on the 32-bit shared address and ADDED to the __shared__ array, so the compiler keeps the shared
// address space (LDS / STS); rounding the generic pointer through uintptr_t degrades every access to generic LD / ST.
__device__ __forceinline__ uint8_t* smem_align1024(uint8_t* raw) {
  const uint32_t base = smem_u32(raw);
  return raw + ((1024u - (base & 1023u)) & 1023u);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Warpgroup register reallocation (sm_90+): all four warps of a warpgroup give registers back to / take registers from
// the CTA's pool.  Roles that run on one thread (TMA producer, MMA issuer) shrink, the register-hungry role grows.
template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
// non-blocking probe (try_wait may suspend the thread for a while; a thread that polls SEVERAL barriers must not)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// 4-byte asynchronous global -> shared copy (no register in between), grouped with commit / wait
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// TMA tiled loads (global -> shared, completion on an mbarrier)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// tcgen05.commit: arrive on an mbarrier once all previously issued MMAs of this thread retire
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], fp16/bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- thread-block-cluster / CTA-pair (cta_group::2) variants --------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Relaxed variant for signals that order nothing but tcgen05.ld (already fenced): no release fence, so the
// arriving warp does not wait for its outstanding global stores to drain.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-SM TMA load: data lands in THIS CTA's shared memory, completion bytes are posted on the barrier at
// shared::cluster address `bar_cluster_addr` (the pair leader's)
__device__ __forceinline__ void tma_load_2d_cg2(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of the pair's MMAs: arrive on the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void tc_commit_cg2(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// D[tmem of both CTAs, 256 x N] (+)= A (128 rows from each CTA) * B (N/2 rows from each CTA)
__device__ __forceinline__ void tc_mma_f16_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets TMEM lane (base_lane + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B.
// Fields (cute/arch/mma_sm100_desc.hpp SmemDescriptor): [0,14) addr>>4, [16,30) LBO>>4,
// [32,46) SBO>>4, [46,48) version=1, [61,64) layout (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16 with fp16 A/B and fp32 accumulate (InstrDescriptor):
// c_format [4,6)=1 (F32); a/b_format [7,10)/[10,13) = 0 (F16); a_major bit 15, b_major bit 16
// (0 = K-major, 1 = MN-major); n_dim [17,23) = N>>3; m_dim [24,29) = M>>4.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// Exact-erf GELU with one ex2 and one rcp per element, flush-to-zero MUFU forms (the tensor-core epilogues
// are instruction-issue bound on this math: every FSETP / FMUL of the denormal-safe library sequences
// counts).  erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, i.e. exact at fp16 output precision);
// erf(|x|/sqrt2) and the Gaussian pdf share exp(-x^2/2); the lower tail Phi(-|x|) = poly(t)*exp(-x^2/2)
// is used directly (no 1 - erf cancellation).
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// tail = Phi(-|x|), e = exp(-x^2/2)
__device__ __forceinline__ void gelu_tail(float x, float& tail, float& e) {
  const float ax = fabsf(x);
  const float a2 = ax * 0.84932180028801904f;             // |x| * sqrt(log2(e) / 2)
  e = fast_ex2(-a2 * a2);                                 // exp(-x^2 / 2)
  const float t = fast_rcp(fmaf(ax, 0.23164189f, 1.0f));  // 1 / (1 + 0.3275911 |x| / sqrt2)
  float poly = fmaf(t, 0.5307027145f, -0.7265760135f);    // 0.5 * A-S coefficients
  poly = fmaf(t, poly, 0.7107068705f);
  poly = fmaf(t, poly, -0.142248368f);
  poly = fmaf(t, poly, 0.127414796f);
  tail = poly * t * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float tail, e;
  gelu_tail(x, tail, e);
  return fmaf(-fabsf(x), tail, fmaxf(x, 0.f));  // x > 0: x - x*tail ; x < 0: x*tail
}
// d/dx gelu_erf(x) = Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float tail, e;
  gelu_tail(x, tail, e);
  const float cdf = x > 0.f ? 1.0f - tail : tail;
  return fmaf(x * 0.3989422804014327f, e, cdf);
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): one instruction works on TWO fp32 values held in a 64-bit
// register pair.  The epilogues and the softmax threads are bound by fp32-pipe issue slots, not by latency, so the
// arithmetic that can be paired is (IEEE fp32 per half, same results as the scalar forms).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ f32x2 pk2(float a) { return pk2(a, a); }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// gelu_erf_both for two values at once: same A-S 7.1.26 evaluation, arranged so that everything but the two MUFU
// calls per element is a packed operation.  With c = 1 / sqrt(2 pi):  e' = c exp(-x^2/2) comes out of ONE ex2 (the
// factor is an addend of the exponent), tail = Phi(-|x|) = (poly(t) / c) t e', Phi(x) = 1/2 + sign(x) (1/2 - tail),
// gelu = x Phi(x), gelu' = Phi(x) + x e'.
__device__ __forceinline__ void gelu_erf_both2(float x0, float x1, f32x2& y, f32x2& dy) {
  const f32x2 x = pk2(x0, x1);
  const f32x2 ax = pk2(fabsf(x0), fabsf(x1));
  float a0, a1, n0, n1;
  upk2(fma2(mul2(ax, ax), pk2(-0.72134752044448170f), pk2(-1.3257480647361593f)), a0, a1);  // -x^2 log2(e)/2 + log2(c)
  upk2(fma2(ax, pk2(0.23164189f), pk2(1.0f)), n0, n1);
  const f32x2 e = pk2(fast_ex2(a0), fast_ex2(a1));
  const f32x2 t = pk2(fast_rcp(n0), fast_rcp(n1));
  f32x2 poly = fma2(t, pk2(1.3302744929f), pk2(-1.8212559978f));  // 0.5 * A-S coefficients / c
  poly = fma2(t, poly, pk2(1.7814779315f));
  poly = fma2(t, poly, pk2(-0.3565637813f));
  poly = fma2(t, poly, pk2(0.3193815300f));
  const f32x2 tail = mul2(mul2(poly, t), e);
  const f32x2 half = pk2(0.5f);
  const f32x2 sgn = pk2(copysignf(1.0f, x0), copysignf(1.0f, x1));
  const f32x2 cdf = fma2(sgn, fma2(tail, pk2(-1.0f), half), half);
  y = mul2(x, cdf);
  dy = fma2(x, e, cdf);
}

// value and derivative from one shared tail / pdf evaluation (the forward GELU epilogue stores both)
__device__ __forceinline__ void gelu_erf_both(float x, float& y, float& dy) {
  float tail, e;
  gelu_tail(x, tail, e);
  const float ax = fabsf(x);
  y = fmaf(-ax, tail, fmaxf(x, 0.f));
  const float cdf = x > 0.f ? 1.0f - tail : tail;
  dy = fmaf(x * 0.3989422804014327f, e, cdf);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace cdr
