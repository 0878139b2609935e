/* cocodr_b200 -- C ABI of the B200 (sm_100a) hot path of OpenMatch/COCO-DR.
 *
 * Plain C: raw device pointers, explicit sizes/strides, fixed-width integers; no torch / C++ types.
 * Every entry point enqueues work on the caller's CUDA stream (`stream` is a cudaStream_t passed as
 * void*), never allocates device memory, never blocks the host (except where documented) and never
 * keeps a pointer past the call.  Return value: CDR_OK or a negative error code; the message is in
 * cdr_last_error() (thread-local).  There is no CPU fallback: a non-sm_100 device is CDR_EARCH.
 *
 * The reference (/root/reference, pure Python) has no FFI of its own for this path: every operator
 * below replaces a *library call* the reference makes through PyTorch / HuggingFace / faiss.  Each
 * declaration cites that call site (file:line relative to the reference root).
 */
#ifndef COCODR_B200_H
#define COCODR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDR_OK 0
#define CDR_EINVAL (-1)     /* bad shape / alignment / argument */
#define CDR_EARCH (-2)      /* device is not sm_100 */
#define CDR_EWORKSPACE (-3) /* workspace too small */
#define CDR_ECUDA (-4)      /* CUDA runtime/driver error, text in cdr_last_error() */
#define CDR_EOVERFLOW (-5)  /* scan candidate buffer overflow (caller retries in safe mode) */

int cdr_version(void);
const char* cdr_last_error(void);
/* CDR_OK iff the current device has compute capability 10.x */
int cdr_device_check(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM   D[M,N] = alpha * A[M,K] * B[N,K]^T  (+ fused epilogue), fp16 operands, fp32 accumulate.
 * Replaces the nn.Linear / torch.matmul calls inside HF BertLayer that the reference reaches through
 * self.bert(...) (ANCE/model/models.py:226, COCO/modeling.py:199-204) and their autograd backward.
 *   a_major/b_major: 0 = K-major  (operand stored [rows, K] row-major, ld = row stride)
 *                    1 = MN-major (operand stored [K, rows] row-major, ld = row stride)
 *   forward  y = x W^T      : A = x   (K-major), B = W  (K-major)
 *   dgrad    dx = dy W      : A = dy  (K-major), B = W  (MN-major, W stored [N_out, K_in])
 *   wgrad    dW = dy^T x    : A = dy  (MN-major), B = x (MN-major), split_k > 1, CDR_EPI_F32_ATOMIC
 * ---------------------------------------------------------------------------------------------- */
enum {
  CDR_EPI_STORE_F16 = 0,     /* out16 = alpha*acc + bias                                          */
  CDR_EPI_BIAS_GELU = 1,     /* out2 (optional) = z = alpha*acc + bias ; out16 = gelu_erf(z)      */
  CDR_EPI_BIAS_RESIDUAL = 2, /* out16 = alpha*acc + bias + aux[m,n]                               */
  CDR_EPI_DGELU = 3,         /* out16 = alpha*acc * gelu_erf'(aux[m,n])                           */
  CDR_EPI_F32_ATOMIC = 4,    /* out32 += alpha*acc  (red.add, for split-K wgrad)                  */
  CDR_EPI_F32_STORE = 5,     /* out32 = alpha*acc                                                 */
  CDR_EPI_SCAN_FILTER = 6    /* internal: threshold-filter scores into candidate buffers          */
};

typedef struct cdr_gemm_args {
  const void* a; /* fp16 */
  const void* b; /* fp16 */
  void* out;        /* fp16 or fp32 [M, ldo] */
  void* out2;       /* optional fp16 [M, ldo] (CDR_EPI_BIAS_GELU pre-activation) */
  const float* bias; /* [N] fp32 or NULL */
  const void* aux;   /* fp16 [M, ldaux] residual / pre-activation, or NULL */
  int64_t M, N, K;
  int64_t lda, ldb, ldo, ldaux; /* in elements */
  int32_t a_major, b_major;
  int32_t epilogue;
  int32_t split_k; /* 0 = auto (wgrad), 1 = none */
  float alpha;
  int32_t dbg_lbo, dbg_sbo; /* 0; test-only descriptor overrides */
} cdr_gemm_args;

int cdr_gemm(const cdr_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COCODR_B200_H */
