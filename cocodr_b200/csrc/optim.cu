// Fused multi-tensor optimizers for the training step of the hot path (SURVEY f-4): one launch updates every
// parameter of a group, and writes the fp16 operand shadows the tcgen05 GEMMs read in the same pass, so the
// separate per-step weight cast disappears.  HBM-bound: AdamW moves 28 B per parameter (+2 B shadow).
//
//   CDR_OPT_ADAMW_TORCH : torch.optim.AdamW        p *= 1 - lr wd ; m,v EMA ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
//   CDR_OPT_ADAMW_HF    : transformers.AdamW (the `AdamW` the reference imports, ANCE/drivers/run_ann.py:19,139-144)
//                         m,v EMA ; p -= lr sqrt(bc2)/bc1 * m / (sqrt(v) + eps) ; p -= lr wd p
//   Lamb (cdr_lamb_multi): ANCE/utils/lamb.py:71-121 -- no bias correction, adam_step = m / (sqrt(v) + eps) + wd p,
//                         trust = clamp(|p|, 0, 10) / |adam_step| (1 if either norm is 0), p -= lr trust adam_step
#include "cdr_common.cuh"
#include "peer.cuh"

namespace cdr {

__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st_shadow4(const cdr_opt_item& it, long long i, const float (&v)[4]) {
  if (it.shadow == nullptr) return;
  if (it.shadow_f32) {
    st4(static_cast<float*>(it.shadow) + i, v);
  } else {
    __half2 lo = __floats2half2_rn(v[0], v[1]), hi = __floats2half2_rn(v[2], v[3]);
    uint2 q;
    q.x = *reinterpret_cast<uint32_t*>(&lo);
    q.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(static_cast<__half*>(it.shadow) + i) = q;
  }
}

__global__ void opt_step_inc_kernel(float* step) { step[0] += 1.f; }

struct OptHyper {
  float beta1, beta2, eps, wd;
  const float* lr;
  const float* step;
  const float* grad_scale;
  int mode;
};

// Work decomposition: block b owns chunk b = elements [start, start + CDR_OPT_CHUNK) of tensor `item` (the chunk
// table is built once per parameter set on the host), so a 23 M-element embedding matrix and a 768-element bias
// cost proportionally -- no empty blocks, no tail imbalance.
__global__ void __launch_bounds__(256)
adam_multi_kernel(const cdr_opt_item* __restrict__ items, const cdr_opt_chunk* __restrict__ chunks, OptHyper h) {
  const cdr_opt_chunk ck = chunks[blockIdx.x];
  cdr_opt_item it = items[ck.item];
  const long long c_end = ck.start + CDR_OPT_CHUNK < it.n ? ck.start + CDR_OPT_CHUNK : it.n;
  const float lr = h.lr[0], t = h.step[0];
  const float gs = h.grad_scale ? h.grad_scale[0] : 1.f;
  const float bc1 = 1.f - powf(h.beta1, t), bc2 = 1.f - powf(h.beta2, t);
  const float rsq_bc2 = rsqrtf(bc2);
  const float step_torch = lr / bc1;                  // with denom = sqrt(v)/sqrt(bc2) + eps
  const float step_hf = lr * sqrtf(bc2) / bc1;        // with denom = sqrt(v) + eps
  const float decay = 1.f - lr * h.wd;
  it.n = c_end;
  for (long long i = ck.start + threadIdx.x * 4; i < c_end; i += 256 * 4) {
    if (i + 4 <= it.n && !it.reserved) {  // reserved != 0: some pointer of this entry is not 16-byte aligned
      float p[4], g[4], m[4], v[4];
      ld4(it.p + i, p); ld4(it.g + i, g); ld4(it.m + i, m); ld4(it.v + i, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = g[k] * gs;
        m[k] = h.beta1 * m[k] + (1.f - h.beta1) * gk;
        v[k] = h.beta2 * v[k] + (1.f - h.beta2) * gk * gk;
        if (h.mode == CDR_OPT_ADAMW_TORCH) {
          p[k] = p[k] * decay - step_torch * m[k] / (sqrtf(v[k]) * rsq_bc2 + h.eps);
        } else {
          p[k] = p[k] - step_hf * m[k] / (sqrtf(v[k]) + h.eps);
          p[k] = p[k] - lr * h.wd * p[k];
        }
      }
      st4(it.p + i, p); st4(it.m + i, m); st4(it.v + i, v);
      st_shadow4(it, i, p);
    } else {
      for (long long j = i; j < it.n && j < i + 4; ++j) {
        const float gk = it.g[j] * gs;
        const float mk = h.beta1 * it.m[j] + (1.f - h.beta1) * gk;
        const float vk = h.beta2 * it.v[j] + (1.f - h.beta2) * gk * gk;
        float pk = it.p[j];
        if (h.mode == CDR_OPT_ADAMW_TORCH) {
          pk = pk * decay - step_torch * mk / (sqrtf(vk) * rsq_bc2 + h.eps);
        } else {
          pk = pk - step_hf * mk / (sqrtf(vk) + h.eps);
          pk = pk - lr * h.wd * pk;
        }
        it.p[j] = pk; it.m[j] = mk; it.v[j] = vk;
        if (it.shadow != nullptr) {
          if (it.shadow_f32) static_cast<float*>(it.shadow)[j] = pk;
          else static_cast<__half*>(it.shadow)[j] = __float2half_rn(pk);
        }
      }
    }
  }
}

// ---- AdamW with the data-parallel gradient exchange fused in (cdr_adam_multi_peer, include/cocodr_b200.h)
__global__ void opt_epoch_inc_kernel(uint32_t* epoch) { epoch[0] += 1u; }
__global__ void opt_step_epoch_inc_kernel(float* step, uint32_t* epoch) {
  step[0] += 1.f;
  epoch[0] += 1u;
}

struct PeerOpt {
  int world, rank;
  long long delta[8];      // byte offset from a local arena address to rank r's copy
  uint32_t* peer_flag[8];  // flag words of rank r: [0, 8) set 0, [8, 16) set 1
  uint32_t* local_flag;
  const uint32_t* epoch;
  uint32_t* done;
  uint32_t* err;
  // mode 0: gradient exchange + update in one pass.  Gradient clipping needs the norm of the REDUCED gradient before
  // any update, so it splits the pass: mode 1 reduces the owned chunks (mean over ranks written back into this rank's
  // own gradient buffer), exchanges the partial sums of squares and leaves clip coefficient and norm in clip_out;
  // mode 2 updates from those local reduced chunks (no peer loads) and stores the result everywhere.
  int mode;
  float max_norm;
  float* sq_local;          // device scratch (zero between calls)
  float* peer_norm[8];      // [2][8] floats on every rank: partial sums of squares, double-buffered by epoch parity
  float* clip_out;          // local: [0] = coefficient min(1, max_norm / (norm + 1e-6)), [1] = norm
};

__device__ __forceinline__ float block_sum_f(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

// bounded spin (a peer that never launches must not hang the GPU): false after ~10 s
__device__ __forceinline__ bool peer_wait_bounded(const uint32_t* flags, int world, uint32_t epoch) {
  const long long t0 = clock64();
  for (int r = 0; r < world; ++r)
    while (static_cast<int>(ld_acquire_sys(flags + r) - epoch) < 0) {
      if (clock64() - t0 > 20000000000ll) return false;
    }
  return true;
}

template <typename T>
__device__ __forceinline__ T* peer_ptr(T* p, long long delta) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(p) + delta);
}
template <typename T>
__device__ __forceinline__ const T* peer_ptr(const T* p, long long delta) {
  return reinterpret_cast<const T*>(reinterpret_cast<const char*>(p) + delta);
}

__global__ void __launch_bounds__(256)
adam_multi_peer_kernel(const cdr_opt_item* __restrict__ items, const cdr_opt_chunk* __restrict__ chunks, int n_chunks,
                       OptHyper h, PeerOpt po) {
  __shared__ int ok;
  __shared__ float red[8];
  const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(po.epoch);
  if (threadIdx.x == 0) {
    ok = 1;
    if (po.mode != 2) {  // (mode 2 reads no peer gradients: nothing to wait for at the start)
      if (blockIdx.x == 0) {
        // every gradient of this rank was written by kernels that precede this one on the stream
        __threadfence_system();
        for (int r = 0; r < po.world; ++r) st_release_sys(po.peer_flag[r] + po.rank, epoch);
      }
      ok = peer_wait_bounded(po.local_flag, po.world, epoch) ? 1 : 0;
      if (!ok && po.err != nullptr) *po.err = 1u;
    }
  }
  __syncthreads();
  float sq = 0.f;
  const int c = blockIdx.x * po.world + po.rank;  // chunks are dealt round-robin to the ranks
  if (ok && c < n_chunks) {
    const cdr_opt_chunk ck = chunks[c];
    const cdr_opt_item it = items[ck.item];
    const long long c_end = ck.start + CDR_OPT_CHUNK < it.n ? ck.start + CDR_OPT_CHUNK : it.n;
    const float lr = h.lr[0], t = h.step[0];
    const float inv_world = 1.f / static_cast<float>(po.world);
    // mean over ranks (DDP); in mode 2 the local buffer already holds the mean
    const float gs = (h.grad_scale ? h.grad_scale[0] : 1.f) * (po.mode == 2 ? 1.f : inv_world);
    const float bc1 = 1.f - powf(h.beta1, t), bc2 = 1.f - powf(h.beta2, t);
    const float rsq_bc2 = rsqrtf(bc2);
    const float step_torch = lr / bc1, step_hf = lr * sqrtf(bc2) / bc1;
    const float decay = 1.f - lr * h.wd;
    auto update = [&](float& pk, float& mk, float& vk, float gsum) {
      const float gk = gsum * gs;
      mk = h.beta1 * mk + (1.f - h.beta1) * gk;
      vk = h.beta2 * vk + (1.f - h.beta2) * gk * gk;
      if (h.mode == CDR_OPT_ADAMW_TORCH) {
        pk = pk * decay - step_torch * mk / (sqrtf(vk) * rsq_bc2 + h.eps);
      } else {
        pk = pk - step_hf * mk / (sqrtf(vk) + h.eps);
        pk = pk - lr * h.wd * pk;
      }
    };
    for (long long i = ck.start + threadIdx.x * 4; i < c_end; i += 256 * 4) {
      if (i + 4 <= c_end && !it.reserved) {
        float p[4], m[4], v[4], g[4] = {0.f, 0.f, 0.f, 0.f};
        if (po.mode == 2) {
          ld4(it.g + i, g);
        } else {
          for (int r = 0; r < po.world; ++r) {  // fixed order: every rank would form the same sum
            float gr[4];
            ld4(peer_ptr(it.g, po.delta[r]) + i, gr);
#pragma unroll
            for (int k = 0; k < 4; ++k) g[k] += gr[k];
          }
        }
        if (po.mode == 1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            g[k] *= inv_world;
            sq = fmaf(g[k], g[k], sq);
          }
          st4(const_cast<float*>(it.g) + i, g);  // this rank's copy of the chunks it owns now holds the mean
          continue;
        }
        ld4(it.p + i, p); ld4(it.m + i, m); ld4(it.v + i, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) update(p[k], m[k], v[k], g[k]);
        st4(it.m + i, m); st4(it.v + i, v);
        for (int r = 0; r < po.world; ++r) {
          st4(peer_ptr(it.p, po.delta[r]) + i, p);
          if (it.shadow != nullptr) {
            cdr_opt_item pit = it;
            pit.shadow = peer_ptr(static_cast<char*>(it.shadow), po.delta[r]);
            st_shadow4(pit, i, p);
          }
        }
      } else {
        for (long long j = i; j < c_end && j < i + 4; ++j) {
          float gsum = 0.f;
          if (po.mode == 2) gsum = it.g[j];
          else for (int r = 0; r < po.world; ++r) gsum += peer_ptr(it.g, po.delta[r])[j];
          if (po.mode == 1) {
            gsum *= inv_world;
            sq = fmaf(gsum, gsum, sq);
            const_cast<float*>(it.g)[j] = gsum;
            continue;
          }
          float pk = it.p[j], mk = it.m[j], vk = it.v[j];
          update(pk, mk, vk, gsum);
          it.m[j] = mk; it.v[j] = vk;
          for (int r = 0; r < po.world; ++r) {
            peer_ptr(it.p, po.delta[r])[j] = pk;
            if (it.shadow != nullptr) {
              if (it.shadow_f32) peer_ptr(static_cast<float*>(it.shadow), po.delta[r])[j] = pk;
              else peer_ptr(static_cast<__half*>(it.shadow), po.delta[r])[j] = __float2half_rn(pk);
            }
          }
        }
      }
    }
  }
  // ---- the last block of the grid tells every peer that this rank's updates have landed (and that it has read all the
  // gradients it needs), then holds the kernel until every peer said the same: afterwards parameters and shadows
  // are complete on this rank and its gradient buffers may be overwritten
  if (po.mode == 1) {  // this block's share of the squared norm of the reduced gradient
    const float t = block_sum_f(sq, red);
    if (threadIdx.x == 0 && t != 0.f) atomicAdd(po.sq_local, t);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(po.done, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence_system();
      const int half = static_cast<int>(epoch & 1u) * 8;
      if (po.mode == 1) {  // partial sum of squares of the chunks this rank owns -> slot `rank` on every rank
        const float mine = *reinterpret_cast<volatile float*>(po.sq_local);
        *po.sq_local = 0.f;
        for (int r = 0; r < po.world; ++r) po.peer_norm[r][half + po.rank] = mine;
        __threadfence_system();
      }
      for (int r = 0; r < po.world; ++r) st_release_sys(po.peer_flag[r] + 8 + po.rank, epoch);
      *po.done = 0u;
      const bool arrived = peer_wait_bounded(po.local_flag + 8, po.world, epoch);
      if (!arrived && po.err != nullptr) *po.err = 1u;
      if (po.mode == 1) {
        float total = 0.f;
        for (int r = 0; r < po.world; ++r) total += reinterpret_cast<volatile float*>(po.peer_norm[po.rank])[half + r];
        const float n = sqrtf(total);
        po.clip_out[0] = fminf(1.f, po.max_norm / (n + 1e-6f));
        po.clip_out[1] = n;
      }
    }
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

// Lamb phase 1: m, v EMA (written); per-tensor sums of p^2 and adam_step^2 -> norms[2 * tensor + {0, 1}]
__global__ void __launch_bounds__(256)
lamb_phase1_kernel(const cdr_opt_item* __restrict__ items, const cdr_opt_chunk* __restrict__ chunks, OptHyper h,
                   float* __restrict__ norms) {
  __shared__ float red[8];
  const cdr_opt_chunk ck = chunks[blockIdx.x];
  const cdr_opt_item it = items[ck.item];
  const long long c_end = ck.start + CDR_OPT_CHUNK < it.n ? ck.start + CDR_OPT_CHUNK : it.n;
  const float gs = h.grad_scale ? h.grad_scale[0] : 1.f;
  float w2 = 0.f, a2 = 0.f;
  for (long long j = ck.start + threadIdx.x; j < c_end; j += 256) {
    const float gk = it.g[j] * gs, pk = it.p[j];
    const float mk = h.beta1 * it.m[j] + (1.f - h.beta1) * gk;
    const float vk = h.beta2 * it.v[j] + (1.f - h.beta2) * gk * gk;
    it.m[j] = mk;
    it.v[j] = vk;
    const float as = mk / (sqrtf(vk) + h.eps) + h.wd * pk;
    w2 += pk * pk;
    a2 += as * as;
  }
  const float tw = block_sum(w2, red), ta = block_sum(a2, red);
  if (threadIdx.x == 0) {
    atomicAdd(norms + 2 * ck.item, tw);
    atomicAdd(norms + 2 * ck.item + 1, ta);
  }
}

// Lamb phase 2: p -= lr * trust * adam_step (adam_step recomputed from the m, v phase 1 wrote)
__global__ void __launch_bounds__(256)
lamb_phase2_kernel(const cdr_opt_item* __restrict__ items, const cdr_opt_chunk* __restrict__ chunks, OptHyper h,
                   const float* __restrict__ norms, float* __restrict__ trust_out) {
  const cdr_opt_chunk ck = chunks[blockIdx.x];
  const cdr_opt_item it = items[ck.item];
  const long long c_end = ck.start + CDR_OPT_CHUNK < it.n ? ck.start + CDR_OPT_CHUNK : it.n;
  const float wn = fminf(sqrtf(norms[2 * ck.item]), 10.f), an = sqrtf(norms[2 * ck.item + 1]);
  const float trust = (wn == 0.f || an == 0.f) ? 1.f : wn / an;
  if (trust_out != nullptr && ck.start == 0 && threadIdx.x == 0) trust_out[ck.item] = trust;
  const float scale = h.lr[0] * trust;
  for (long long j = ck.start + threadIdx.x; j < c_end; j += 256) {
    float pk = it.p[j];
    const float as = it.m[j] / (sqrtf(it.v[j]) + h.eps) + h.wd * pk;
    pk -= scale * as;
    it.p[j] = pk;
    if (it.shadow != nullptr) {
      if (it.shadow_f32) static_cast<float*>(it.shadow)[j] = pk;
      else static_cast<__half*>(it.shadow)[j] = __float2half_rn(pk);
    }
  }
}

// out[0] += sum over every tensor of the table of g^2  (global gradient norm for clip_grad_norm_)
__global__ void __launch_bounds__(256)
grad_sqnorm_kernel(const cdr_opt_item* __restrict__ items, const cdr_opt_chunk* __restrict__ chunks,
                   float* __restrict__ out) {
  __shared__ float red[8];
  const cdr_opt_chunk ck = chunks[blockIdx.x];
  const cdr_opt_item it = items[ck.item];
  const long long c_end = ck.start + CDR_OPT_CHUNK < it.n ? ck.start + CDR_OPT_CHUNK : it.n;
  float s = 0.f;
  for (long long j = ck.start + threadIdx.x; j < c_end; j += 256) {
    const float g = it.g[j];
    s += g * g;
  }
  const float t = block_sum(s, red);
  if (threadIdx.x == 0 && t != 0.f) atomicAdd(out, t);
}

// coef[0] = min(1, max_norm / (sqrt(sq[0]) + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void clip_coef_kernel(const float* sq, float max_norm, float* coef, float* norm_out) {
  const float n = sqrtf(sq[0]);
  if (norm_out) norm_out[0] = n;
  coef[0] = fminf(1.f, max_norm / (n + 1e-6f));
}

static int opt_check(const cdr_opt_args* a, const char* who) {
  CDR_REQUIRE(a != nullptr && a->items != nullptr && a->count > 0, "%s: bad table", who);
  CDR_REQUIRE(a->chunks != nullptr && a->n_chunks > 0, "%s: bad chunk table", who);
  CDR_REQUIRE(a->lr != nullptr && a->step != nullptr, "%s: lr and step must be device pointers", who);
  return CDR_OK;
}

}  // namespace cdr

using namespace cdr;

extern "C" {

int cdr_adam_multi(const cdr_opt_args* a, void* stream) {
  if (int rc = opt_check(a, "cdr_adam_multi")) return rc;
  CDR_REQUIRE(a->mode == CDR_OPT_ADAMW_TORCH || a->mode == CDR_OPT_ADAMW_HF, "cdr_adam_multi: unknown mode %d", a->mode);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  opt_step_inc_kernel<<<1, 1, 0, st>>>(a->step);
  CDR_LAUNCH_CHECK();
  OptHyper h{a->beta1, a->beta2, a->eps, a->weight_decay, a->lr, a->step, a->grad_scale, a->mode};
  adam_multi_kernel<<<a->n_chunks, 256, 0, st>>>(a->items, a->chunks, h);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

static int adam_peer_launch(const cdr_opt_args* a, const cdr_peer_args* pa, uint32_t* epoch_rw, uint32_t* err, int mode,
                            float max_norm, float* sq_scratch, float* const* peer_norm, float* clip_out, void* stream,
                            const char* who) {
  if (int rc = opt_check(a, who)) return rc;
  if (int rc = peer_check(pa, who)) return rc;
  CDR_REQUIRE(a->mode == CDR_OPT_ADAMW_TORCH || a->mode == CDR_OPT_ADAMW_HF, "%s: unknown mode %d", who, a->mode);
  CDR_REQUIRE(epoch_rw != nullptr && epoch_rw == pa->epoch, "%s: epoch_rw must be peers->epoch (it is incremented)", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == 1) opt_epoch_inc_kernel<<<1, 1, 0, st>>>(epoch_rw);
  else opt_step_epoch_inc_kernel<<<1, 1, 0, st>>>(a->step, epoch_rw);
  CDR_LAUNCH_CHECK();
  OptHyper h{a->beta1, a->beta2, a->eps, a->weight_decay, a->lr, a->step, a->grad_scale, a->mode};
  PeerOpt po{};
  po.mode = mode;
  po.max_norm = max_norm;
  po.sq_local = sq_scratch;
  po.clip_out = clip_out;
  if (peer_norm != nullptr)
    for (int r = 0; r < pa->world; ++r) po.peer_norm[r] = peer_norm[r];
  po.world = pa->world;
  po.rank = pa->rank;
  for (int r = 0; r < pa->world; ++r) {
    po.delta[r] = static_cast<long long>(reinterpret_cast<intptr_t>(pa->peer_buf[r]) - reinterpret_cast<intptr_t>(pa->peer_buf[pa->rank]));
    po.peer_flag[r] = pa->peer_flag[r];
  }
  po.local_flag = pa->peer_flag[pa->rank];
  po.epoch = pa->epoch;
  po.done = pa->done_counter;
  po.err = err;
  const int blocks = (a->n_chunks + pa->world - 1) / pa->world;
  adam_multi_peer_kernel<<<blocks, 256, 0, st>>>(a->items, a->chunks, a->n_chunks, h, po);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_adam_multi_peer(const cdr_opt_args* a, const cdr_peer_args* pa, uint32_t* epoch_rw, uint32_t* err, void* stream) {
  return adam_peer_launch(a, pa, epoch_rw, err, 0, 0.f, nullptr, nullptr, nullptr, stream, "cdr_adam_multi_peer");
}

int cdr_grad_reduce_clip_peer(const cdr_opt_args* a, const cdr_peer_args* pa, uint32_t* epoch_rw, float max_norm,
                              float* sq_scratch, float* const* peer_norm, float* clip_out, uint32_t* err, void* stream) {
  CDR_REQUIRE(sq_scratch != nullptr && peer_norm != nullptr && clip_out != nullptr && max_norm > 0.f,
              "cdr_grad_reduce_clip_peer: scratch, norm slots, clip_out and a positive max_norm are required");
  if (pa != nullptr)
    for (int r = 0; r < pa->world && r < 8; ++r)
      CDR_REQUIRE(peer_norm[r] != nullptr, "cdr_grad_reduce_clip_peer: null norm slot pointer %d", r);
  return adam_peer_launch(a, pa, epoch_rw, err, 1, max_norm, sq_scratch, peer_norm, clip_out, stream,
                          "cdr_grad_reduce_clip_peer");
}

int cdr_adam_multi_peer_reduced(const cdr_opt_args* a, const cdr_peer_args* pa, uint32_t* epoch_rw, uint32_t* err,
                                void* stream) {
  return adam_peer_launch(a, pa, epoch_rw, err, 2, 0.f, nullptr, nullptr, nullptr, stream, "cdr_adam_multi_peer_reduced");
}

int cdr_lamb_multi(const cdr_opt_args* a, void* stream) {
  if (int rc = opt_check(a, "cdr_lamb_multi")) return rc;
  CDR_REQUIRE(a->norms != nullptr, "cdr_lamb_multi: norms scratch (2 floats per tensor) is required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  opt_step_inc_kernel<<<1, 1, 0, st>>>(a->step);
  CDR_LAUNCH_CHECK();
  CDR_CUDA(cudaMemsetAsync(a->norms, 0, sizeof(float) * 2 * a->count, st));
  OptHyper h{a->beta1, a->beta2, a->eps, a->weight_decay, a->lr, a->step, a->grad_scale, 0};
  lamb_phase1_kernel<<<a->n_chunks, 256, 0, st>>>(a->items, a->chunks, h, a->norms);
  CDR_LAUNCH_CHECK();
  lamb_phase2_kernel<<<a->n_chunks, 256, 0, st>>>(a->items, a->chunks, h, a->norms, a->trust);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_grad_sqnorm_multi(const cdr_opt_item* items, const cdr_opt_chunk* chunks, int32_t n_chunks, float* sq_accum,
                          void* stream) {
  CDR_REQUIRE(items && chunks && n_chunks > 0 && sq_accum, "cdr_grad_sqnorm_multi: bad arguments");
  grad_sqnorm_kernel<<<n_chunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(items, chunks, sq_accum);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

int cdr_grad_clip_coef(const float* sq, float max_norm, float* coef, float* norm_out, void* stream) {
  CDR_REQUIRE(sq && coef, "cdr_grad_clip_coef: bad arguments");
  clip_coef_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(sq, max_norm, coef, norm_out);
  CDR_LAUNCH_CHECK();
  return CDR_OK;
}

}  // extern "C"
