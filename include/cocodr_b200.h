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
/* Size the grids of the persistent kernels for n SMs (0 = all): leaves SMs to overlapped communication kernels. */
int cdr_set_sm_budget(int32_t n);

/* ------------------------------------------------------------------------------------------------
 * Counter-based dropout.  Replaces the nn.Dropout modules of HF BERT that the reference trains with (p = 0.1 on the
 * embeddings, on the attention probabilities and on both dense outputs of every layer: BertEmbeddings,
 * BertSelfAttention, BertSelfOutput, BertOutput reached through self.bert(...) in ANCE/model/models.py:226 under
 * model.train(), ANCE/drivers/run_ann.py:320; COCO/modeling.py:216-220 for the c_head layers).  No mask is stored:
 * forward and backward kernels regenerate it with Philox4x32-10 from
 *     counter = (group index, site, offset lo, offset hi), key = (seed lo, seed hi)
 * where one call covers a group of 8 consecutive elements of a row (8 x 16 random bits); element j is kept iff its
 * 16 bits are >= threshold (= round(p * 65536)) and kept values are multiplied by `scale` (= 1 / (1 - p)).
 * Group index of element (m, n) of a [rows, cols] activation: (m * row_mul) * (cols / 8) + n / 8; of attention
 * probability (seq, head, query row r, key c): ((seq * heads + head) * seq_len + r) * 64 + c / 8.
 * `state` is a DEVICE array {seed, offset} read by the kernels (the step counter advances on the device, so a CUDA
 * graph replays with fresh masks); state == NULL or threshold == 0 disables dropout.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cdr_dropout {
  const uint64_t* state; /* device {seed, offset}, or NULL */
  uint32_t site;         /* which dropout module (distinct per layer and position) */
  uint32_t threshold;    /* 0 .. 65535 */
  float scale;
  int32_t row_mul;       /* >= 1 (0 is read as 1) */
  void* keep_bits;       /* optional device [rows, cols / 8] bytes (cols % 32 == 0), bit j of byte (m, n / 8) = element
                            (m, 8 (n / 8) + j) is kept: written by the CDR_EPI_BIAS_DROP_RESIDUAL GEMM epilogue next to
                            its output and read by cdr_ln_bwd_drop instead of regenerating the masks (3 Philox calls per
                            row and lane in a latency-bound kernel); NULL = regenerate.  Other entry points ignore it */
} cdr_dropout;
/* out[m, :] = dropout(x[m, :]) on a contiguous fp16 [rows, cols] tensor (cols % 8 == 0); out may alias x. */
int cdr_dropout_f16(const void* x, void* out, int64_t rows, int32_t cols, const cdr_dropout* drop, void* stream);

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
  CDR_EPI_BIAS_GELU = 1,     /* z = alpha*acc + bias ; out16 = gelu_erf(z) ; out2 (optional) = gelu_erf'(z) */
  CDR_EPI_BIAS_RESIDUAL = 2, /* out16 = alpha*acc + bias + aux[m,n]                               */
  CDR_EPI_DGELU = 3,         /* out16 = alpha*acc * aux[m,n], aux = the gelu_erf'(z) saved by BIAS_GELU;
                                optional colsum[n] += colsum_scale * sum_m out[m,n] (the bias gradient) */
  CDR_EPI_F32_ATOMIC = 4,    /* out32 += alpha*acc  (red.add, for split-K wgrad)                  */
  CDR_EPI_F32_STORE = 5,     /* out32 = alpha*acc                                                 */
  CDR_EPI_SCAN_FILTER = 6,   /* internal: threshold-filter scores into candidate buffers (docs on M, queries on N) */
  CDR_EPI_SCAN_FILTER_Q = 7, /* internal: the same with queries on M (<= 128 queries: documents stream as the B operand) */
  CDR_EPI_F32_GROUPED = 8,   /* internal (cdr_gemm_grouped): F32_ATOMIC with per-group K ranges and output bases */
  CDR_EPI_BIAS_DROP_RESIDUAL = 9 /* out16 = dropout(alpha*acc + bias) + aux[m,n]   (args.drop; HF BertSelfOutput / BertOutput) */
};

typedef struct cdr_gemm_args {
  const void* a; /* fp16 */
  const void* b; /* fp16 */
  void* out;        /* fp16 or fp32 [M, ldo] */
  void* out2;       /* optional fp16 [M, ldo] (CDR_EPI_BIAS_GELU: derivative gelu_erf'(z) for the backward) */
  const float* bias; /* [N] fp32 or NULL */
  const void* aux;   /* fp16 [M, ldaux] residual / saved GELU derivative, or NULL */
  int64_t M, N, K;
  int64_t lda, ldb, ldo, ldaux; /* in elements */
  int32_t a_major, b_major;
  int32_t epilogue;
  int32_t split_k; /* 0 = auto (wgrad), 1 = none */
  float alpha;
  int32_t dbg_lbo, dbg_sbo; /* 0; descriptor overrides honoured only by -DCDR_GEMM_DEBUG builds (tools/build_variant.sh) */
  float* colsum;      /* optional fp32 [N], CDR_EPI_DGELU only: accumulates the column sums of the output */
  float colsum_scale;
  int32_t reserved;
  cdr_dropout drop;   /* CDR_EPI_BIAS_DROP_RESIDUAL only */
} cdr_gemm_args;

int cdr_gemm(const cdr_gemm_args* args, void* stream);

/* Row-segmented reduction (iDRO per-group wgrad, K11: replaces the G partial backwards of
 * ANCE/model/dro_loss.py:192-204 for the weights).  `base` describes a GEMM whose operands are both MN-major
 * (a_major = b_major = 1: A [K, M], B [K, N], i.e. K runs over rows) with an fp32 epilogue; for every segment i the
 * same GEMM is enqueued on rows [row_begin[i], row_begin[i] + row_count[i]) of A and B with its output at
 * (float*)base->out + out_offset[i] (elements).  The three arrays are HOST arrays of n_seg entries; segments with
 * row_count 0 are skipped. */
int cdr_gemm_segments(const cdr_gemm_args* base, int32_t n_seg, const int64_t* row_begin, const int64_t* row_count,
                      const int64_t* out_offset, void* stream);
/* The same reduction as ONE launch, with the segment table in DEVICE memory (no host transfer of the group sizes):
 * seg_kb[n_groups + 1] are ascending boundaries in units of 64 rows (k-blocks); group g reduces rows
 * [64 * seg_kb[g], 64 * seg_kb[g + 1]) of A and B (both MN-major, base->K = total rows) and ADDS alpha * A_g^T B_g to
 * (float*)base->out + g * out_group_stride (elements); groups with an empty range are left untouched.  Rows that pad a
 * group to a multiple of 64 must be zero.  base->epilogue must be CDR_EPI_F32_ATOMIC. */
int cdr_gemm_grouped(const cdr_gemm_args* base, int32_t n_groups, const int32_t* seg_kb, int64_t out_group_stride,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Memory-bound encoder kernels (K1, LN halves of K4/K6, bias gradients, casts).
 * Replace HF BertEmbeddings / nn.LayerNorm (+ their autograd) reached through self.bert(...)
 * (ANCE/model/models.py:226, COCO/modeling.py:199-204).  fp16 activations, fp32 parameters/statistics.
 * in_scale multiplies the incoming activation gradient, out_scale the emitted parameter gradients
 * (loss-scaling plumbing).  Parameter gradients are ACCUMULATED (+=) into the given buffers.
 * ---------------------------------------------------------------------------------------------- */
int cdr_embed_ln_fwd(const int64_t* ids, const float* word, const float* pos, const float* type0, const float* gamma,
                     const float* beta, void* out, float* mean, float* rstd, int32_t n_seq, int32_t seq_len,
                     int32_t hidden, int32_t vocab, float eps, void* stream);
int cdr_embed_ln_bwd(const void* dy, const int64_t* ids, const float* word, const float* pos, const float* type0,
                     const float* gamma, const float* mean, const float* rstd, float* dword, float* dpos, float* dtype0,
                     float* dgamma, float* dbeta, int32_t n_seq, int32_t seq_len, int32_t hidden, int32_t vocab,
                     int32_t pad_id, float in_scale, float out_scale, void* stream);
/* y = LN(x); optional cls_out[n_seq, hidden] fp32 = row 0 of every sequence (K7, CLS pooling:
 * ANCE/model/models.py:228, COCO/modeling.py:206) */
int cdr_ln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, float* cls_out,
               int32_t n_seq, int32_t seq_len, int32_t hidden, float eps, void* stream);
/* cdr_ln_bwd (below) for the LayerNorm that follows a dropped dense output y = x + dropout(d): besides dx (the
 * gradient of the residual branch) it writes dx_drop = dropout'(dx) -- the operand of the dense layer's dgrad / wgrad
 * GEMMs, mask regenerated from `drop` -- and dbias += out_scale * column sums of dx_drop.  Staged path only: fp16 dy,
 * hidden <= 1024, 16-byte aligned operands (CDR_EINVAL otherwise). */
int cdr_ln_bwd_drop(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, void* dx,
                    void* dx_drop, float* dgamma, float* dbeta, float* dbias, int32_t rows, int32_t hidden,
                    float out_scale, const cdr_dropout* drop, void* stream);
/* dx = LN'(dy [+ in_scale * dy_cls on row 0 of every sequence]); dbias (optional) += column sums of dx.
 * dy (fp16) is already in the scaled-gradient domain; dy_cls (fp32) enters it through in_scale. */
/* row_ws: optional scratch of 2*n_seq*seq_len floats; when given (and dy_cls is NULL) the backward runs as two
 * bandwidth-shaped passes (dx, then column sums) instead of one fused kernel. */
int cdr_ln_bwd(const void* dy, const float* dy_cls, const void* x, const float* gamma, const float* mean,
               const float* rstd, void* dx, float* dgamma, float* dbeta, float* dbias, float* row_ws, int32_t n_seq,
               int32_t seq_len, int32_t hidden, float in_scale, float out_scale, void* stream);
int cdr_colsum_f16(const void* x, float* out, int64_t rows, int64_t cols, int64_t ld, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Peer-memory exchange of the contrastive head over NVLink.  Replaces the NCCL all-gather of the
 * passage CLS embeddings / reduce-scatter of their gradients (the reference's dist.all_gather in
 * COCO/modeling.py:182-190; K9' uses the same exchange) with stores into symmetric buffers: every rank holds
 * the base pointer of the SAME allocation on every rank (peer_buf[r], peer_flag[r]; index `rank` is its own).
 * peer_flag[r] points at 16 uint32 (two sets of 8: forward, backward), zero-initialised once.  `epoch` is a LOCAL
 * device counter (so the whole exchange is CUDA-graph capturable): cdr_peer_next_epoch advances it once per
 * training step, identically on every rank.  done_counter: one zero-initialised local uint32 of scratch.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cdr_peer_args {
  int32_t world, rank;
  const uint32_t* epoch;
  void* peer_buf[8];
  uint32_t* peer_flag[8];
  uint32_t* done_counter;
} cdr_peer_args;
/* cdr_ln_fwd that ALSO pushes the fp32 CLS row of every sequence >= first_seq into row
 * rank * (n_seq - first_seq) + (seq - first_seq) of each peer's gather area and raises flag set 0 when done (fused
 * compute + all-gather).  The gather area is DOUBLE-BUFFERED: peer_buf[r] points at 2 x [world * (n_seq - first_seq),
 * hidden] floats and epoch parity selects the half, so a rank that is one step ahead never overwrites rows a slower
 * peer is still reading.  cdr_peer_wait_fetch waits for the world's pushes of the current epoch and copies that half
 * into a private [world * rows, hidden] tensor. */
int cdr_ln_fwd_push(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                    float* cls_out, int32_t n_seq, int32_t seq_len, int32_t hidden, float eps, int32_t first_seq,
                    const cdr_peer_args* peers, void* stream);
int cdr_peer_next_epoch(uint32_t* epoch, void* stream);
int cdr_peer_wait_fetch(const uint32_t* local_flags, int32_t world, const uint32_t* epoch, const float* gather,
                        int64_t half_elems, float* out, void* stream);
/* hold the stream until local_flags[0..world) all carry *epoch (or a later one) */
int cdr_peer_wait(const uint32_t* local_flags, int32_t world, const uint32_t* epoch, void* stream);
/* reduce-scatter, push side: row block r of src [world*rows, dim] -> slot `rank` of rank r's receive buffer
 * [world][rows, dim]; raises flag set 1.  Pull side: out[rows*dim] = sum over slots once all flags arrived. */
int cdr_peer_scatter_rows(const float* src, int32_t rows, int32_t dim, const cdr_peer_args* peers, void* stream);
int cdr_peer_reduce_slots(const float* recv, const uint32_t* local_flags, int32_t world, int64_t n,
                          const uint32_t* epoch, float* out, void* stream);
int cdr_cast_f32_f16(const float* src, void* dst, int64_t n, void* stream);
/* A whole table of casts in one launch (per-step refresh of all fp16 weight shadows).  The table lives in
 * DEVICE memory; every src/dst must be 16-byte aligned.  dst_f32 != 0 copies fp32 -> fp32 instead. */
typedef struct cdr_cast_item {
  const float* src;
  void* dst;
  int64_t n;
  int32_t dst_f32;
  int32_t reserved;
} cdr_cast_item;
int cdr_cast_multi(const cdr_cast_item* items_device, int32_t count, int64_t max_n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-tensor optimizers (SURVEY f-4).  Replace the per-tensor Python loops of the reference's optimizers
 * -- utils/lamb.py Lamb.step (ANCE/utils/lamb.py:61-121), transformers.AdamW / torch.optim.AdamW
 * (ANCE/drivers/run_ann.py:134-147) -- and torch.nn.utils.clip_grad_norm_ (run_ann.py:345-352) with one launch
 * per parameter group; the same pass writes the fp16 (or packed fp32) operand shadow of each parameter.
 * The table lives in DEVICE memory; p, g, m, v, shadow must be 16-byte aligned.  lr / step / grad_scale are
 * device scalars (CUDA-graph friendly); `step` (fp32 counter) is incremented by the call.
 * ---------------------------------------------------------------------------------------------- */
enum { CDR_OPT_ADAMW_TORCH = 0, CDR_OPT_ADAMW_HF = 1 };
typedef struct cdr_opt_item {
  float* p;        /* parameter, updated in place */
  const float* g;  /* gradient */
  float* m;        /* exp_avg */
  float* v;        /* exp_avg_sq */
  void* shadow;    /* optional copy of the updated parameter: fp16, or fp32 when shadow_f32 != 0 */
  int64_t n;
  int32_t shadow_f32;
  int32_t reserved; /* set to 1 when p / g / m / v are not all 16-byte aligned (or shadow not 8-byte): scalar path */
} cdr_opt_item;
/* Work list: block b of a launch processes elements [start, start + CDR_OPT_CHUNK) of tensor `item`; the caller
 * enumerates every chunk of every tensor once (ceil(n / CDR_OPT_CHUNK) entries per tensor). */
#define CDR_OPT_CHUNK 16384
typedef struct cdr_opt_chunk {
  int64_t start;
  int32_t item;
  int32_t reserved;
} cdr_opt_chunk;
typedef struct cdr_opt_args {
  const cdr_opt_item* items;   /* device */
  const cdr_opt_chunk* chunks; /* device */
  int32_t count;               /* tensors */
  int32_t n_chunks;
  int32_t mode;                /* cdr_adam_multi: CDR_OPT_ADAMW_* */
  int32_t reserved;
  float beta1, beta2, eps, weight_decay;
  const float* lr;           /* device scalar */
  float* step;               /* device scalar, += 1 */
  const float* grad_scale;   /* device scalar multiplied into every gradient (clip coefficient / unscale), or NULL */
  float* norms;              /* cdr_lamb_multi: device scratch, 2 floats per tensor */
  float* trust;              /* cdr_lamb_multi: optional device [count] trust ratios (lamb.py:114-116) */
} cdr_opt_args;
int cdr_adam_multi(const cdr_opt_args* args, void* stream);
int cdr_lamb_multi(const cdr_opt_args* args, void* stream);
/* Data-parallel AdamW with the gradient exchange INSIDE the optimizer kernel (replaces DistributedDataParallel's
 * bucketed all-reduce -- ANCE/drivers/run_ann.py:178-184 -- plus the replicated optimizer step, :345-353).  Every
 * p / g / shadow of the table lies in a symmetric arena with the same layout on every rank (peers->peer_buf[r] = arena
 * base of rank r, so rank r's copy of a local address a is a + (peer_buf[r] - peer_buf[rank])).  Rank r owns the chunks
 * c with c % world == r: it reads that chunk of the gradient from EVERY rank over NVLink, averages, updates its own
 * exp_avg / exp_avg_sq and writes the new parameter (and shadow) into every rank's arena -- reduce-scatter, optimizer
 * on 1 / world of the parameters, all-gather, in one pass over peer memory with no collective kernel competing with
 * the GEMMs of the backward.  Flag set 0 of peers->peer_flag publishes "my gradients are final", set 1 "my updates
 * have landed everywhere and I have read everything I need"; the call returns (stream-wise) only after all ranks
 * signalled set 1.  m / v / step of chunks a rank does not own are not touched (cdr_opt_item.m / .v may be stale
 * there).  err (optional device uint32) is set to 1 if a peer did not arrive within ~10 s. */
int cdr_adam_multi_peer(const cdr_opt_args* args, const cdr_peer_args* peers, uint32_t* epoch_rw, uint32_t* err,
                        void* stream);
/* The same with gradient clipping (torch.nn.utils.clip_grad_norm_ before optimizer.step(), run_ann.py:345-353): the
 * norm of the REDUCED gradient must exist before any update, so the pass splits in two.
 * cdr_grad_reduce_clip_peer: every rank reduces the chunks it owns (mean over ranks, written back into ITS OWN gradient
 * buffer), adds up their squares, publishes that partial sum in slot `rank` of every rank's peer_norm area ([2][8]
 * floats per rank, double-buffered by epoch parity; peer_norm[r] = rank r's area) and, once all partial sums arrived,
 * leaves clip_out[0] = min(1, max_norm / (norm + 1e-6)), clip_out[1] = norm.  sq_scratch: one zeroed device float.
 * cdr_adam_multi_peer_reduced: the update of the owned chunks from those local reduced gradients (args->grad_scale =
 * clip_out), stored into every rank's arena; same completion semantics as cdr_adam_multi_peer. */
int cdr_grad_reduce_clip_peer(const cdr_opt_args* args, const cdr_peer_args* peers, uint32_t* epoch_rw, float max_norm,
                              float* sq_scratch, float* const* peer_norm, float* clip_out, uint32_t* err, void* stream);
int cdr_adam_multi_peer_reduced(const cdr_opt_args* args, const cdr_peer_args* peers, uint32_t* epoch_rw, uint32_t* err,
                                void* stream);
/* torch.nn.utils.clip_grad_norm_ in two steps: sq_accum[0] += sum of g^2 over every gradient of a table (call once
 * per parameter group on a zeroed scalar), then coef[0] = min(1, max_norm / (sqrt(sq[0]) + 1e-6)) and, optionally,
 * norm_out[0] = sqrt(sq[0]).  The coefficient is consumed on the device as cdr_opt_args.grad_scale. */
int cdr_grad_sqnorm_multi(const cdr_opt_item* items, const cdr_opt_chunk* chunks, int32_t n_chunks, float* sq_accum,
                          void* stream);
int cdr_grad_clip_coef(const float* sq, float max_norm, float* coef, float* norm_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-head attention (K3): softmax(Q K^T * scale + key_bias) V on tcgen05, head_dim 64,
 * seq_len <= 512 (one 128 x 128 tile up to 128, tiled above), straight from / to the packed QKV projection.  Replaces HF
 * eager_attention_forward / SDPA inside BertSelfAttention (reached through ANCE/model/models.py:226,
 * COCO/modeling.py:199-204), including the dropout of the attention probabilities (args.drop).
 * ---------------------------------------------------------------------------------------------- */
typedef struct cdr_attn_args {
  const void* qkv;       /* fp16 [n_seq*seq_len, 3*heads*64]  (Q | K | V) */
  const float* key_bias; /* fp32 [n_seq, seq_len] additive key mask (0 / -large), or NULL */
  void* out;             /* fp16 ctx [n_seq*seq_len, heads*64]   fwd: written, bwd: read */
  float* lse;            /* fp32 [n_seq, heads, seq_len]         fwd: written, bwd: read */
  const void* d_out;     /* bwd: fp16 [n_seq*seq_len, heads*64] */
  void* dqkv;            /* bwd: fp16 [n_seq*seq_len, 3*heads*64] written */
  float* dq_workspace;   /* bwd, seq_len > 128 only: fp32 [n_seq*seq_len, heads*64] scratch */
  int32_t n_seq, seq_len, heads, head_dim;
  float scale;
  float dbias_scale;     /* bwd, seq_len <= 128: multiplies the column sums below */
  float* dbias_qkv;      /* bwd, optional fp32 [3*heads*64]: += dbias_scale * column sums of dqkv (QKV bias gradient) */
  cdr_dropout drop;      /* dropout of the attention probabilities (after the softmax, HF BertSelfAttention) */
  void* drop_bits;       /* with dropout: cdr_attn_dropout_bits_bytes() bytes, 16-byte aligned.  cdr_attn_fwd FILLS it (one
                            keep bit per probability, from the Philox counters of `drop`) and reads it; cdr_attn_bwd reads
                            the same buffer -- the softmax threads never run the generator themselves */
  int32_t drop_bits_ready; /* fwd: nonzero = drop_bits was already filled by cdr_attn_dropout_bits_fill (e.g. on another
                              stream, overlapped with the preceding GEMMs); cdr_attn_fwd then only reads it */
  int32_t reserved;
} cdr_attn_args;
size_t cdr_attn_dropout_bits_bytes(int32_t n_seq, int32_t heads, int32_t seq_len);
/* Fills args->drop_bits from args->drop for (n_seq, heads, seq_len): the generator pass of cdr_attn_fwd on its own.  It
 * depends on nothing but the dropout state, so a caller may run it ahead of time on a second stream. */
int cdr_attn_dropout_bits_fill(const cdr_attn_args* args, void* stream);
int cdr_attn_fwd(const cdr_attn_args* args, void* stream);
int cdr_attn_bwd(const cdr_attn_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Contrastive heads (fp32 throughout: trained CLS dot products are ~217 with gaps ~0.3).
 * ---------------------------------------------------------------------------------------------- */
/* K8  ANCE pairwise NLL (ANCE/model/models.py:101-108): logits = [<q,a>, <q,b>],
 * loss = -log_softmax(logits)[0], accs = argmax(logits) (int64). */
int cdr_pair_nll_fwd(const float* q, const float* a, const float* b, int32_t n, int32_t dim, float* loss,
                     int64_t* accs, float* logits, void* stream);
int cdr_pair_nll_bwd(const float* q, const float* a, const float* b, const float* logits, const float* dloss,
                     int32_t n, int32_t dim, float* dq, float* da, float* db, void* stream);

/* K9 / K9'  similarity matrix + softmax cross-entropy.
 *   CDR_SIM_COCO (COCO/modeling.py:244-248): S = q k^T, S[i, row_offset+i] = -inf,
 *                target_i = (row_offset+i) ^ 1, loss_i = loss_scale * CE(S[i,:], target_i)
 *   CDR_SIM_QP   (in-batch q x p InfoNCE over all-gathered passages): target_i = row_offset + i.
 * bwd: dq (optional) = dS k, dk (optional) = dS^T q with dS = loss_scale*dloss_i*(softmax - onehot), the score tiles
 * recomputed from q, k and the saved row log-sum-exps (fused gather -> q k^T -> softmax-CE, SURVEY K9: S is never
 * materialised; dim <= 2048). */
enum { CDR_SIM_QP = 0, CDR_SIM_COCO = 1 };
typedef struct cdr_simmat_args {
  const float* q;     /* [n_rows, dim] */
  const float* k;     /* [n_keys, dim] */
  float* scores;      /* fwd: workspace of cdr_simmat_workspace_bytes(n_rows, n_keys) bytes (per-split softmax partials;
                         the score matrix itself is NEVER written: tiles of q k^T live in registers / shared memory, in
                         the forward and again in the backward).  bwd: unused, may be NULL */
  float* gmat;        /* unused (kept for layout compatibility), may be NULL */
  float* loss;        /* [n_rows] fwd out */
  float* lse;         /* [n_rows] fwd out, bwd in */
  const float* dloss; /* [n_rows] bwd in */
  float* dq;          /* [n_rows, dim] bwd out or NULL */
  float* dk;          /* [n_keys, dim] bwd out or NULL */
  int32_t n_rows, n_keys, dim, mode, row_offset;
  float loss_scale;
} cdr_simmat_args;
size_t cdr_simmat_workspace_bytes(int32_t n_rows, int32_t n_keys);
int cdr_simmat_ce_fwd(const cdr_simmat_args* args, void* stream);
int cdr_simmat_ce_bwd(const cdr_simmat_args* args, void* stream);
/* CDR_SIM_QP, gradient of row i towards its OWN key only: dk_own[i, :] = dloss[i] * (softmax_i,own - 1) * q[i, :] with
 * softmax_i,own = exp(-loss[i]) (loss from cdr_simmat_ce_fwd with loss_scale 1).  With cdr_simmat_ce_bwd(dk = NULL)
 * this is the backward of the loss view "every key but the sample's own positive is a constant", from which iDRO
 * takes its group gradients for the in-batch head (ANCE/model/dro_loss.py:192-204 differentiates the group means;
 * see cocodr_b200/models.py BertDot_InBatch_NLL_LN). */
int cdr_simmat_own_key_grad(const float* q, const float* loss, const float* dloss, int32_t n, int32_t dim, float* dk_own,
                            void* stream);

/* K14 MLM head loss on gathered masked rows (HF BertForMaskedLM cross-entropy reached through
 * COCO/modeling.py:87-93, 199-204): logits fp32 [n_rows, ld] from the decoder GEMM, bias fp32 [n_cols] added
 * here (padding columns hold -inf), labels in [0, n_cols).  bwd writes fp16 dlogits = scale * dloss_i *
 * (softmax - onehot).  cdr_dgelu_f16: dz = dt * gprime for the MLM transform (GELU between GEMM and LN), gprime =
 * the gelu_erf'(z) tensor saved by the CDR_EPI_BIAS_GELU epilogue. */
int cdr_vocab_ce_fwd(const float* logits, const float* bias, const int64_t* labels, float* loss, float* lse,
                     int32_t n_rows, int32_t n_cols, int64_t ld, void* stream);
int cdr_vocab_ce_bwd(const float* logits, const float* bias, const int64_t* labels, const float* lse,
                     const float* dloss, void* dlogits, int32_t n_rows, int32_t n_cols, int64_t ld, float scale,
                     void* stream);
int cdr_dgelu_f16(const void* dt, const void* gprime, void* dz, int64_t n, void* stream);

/* K10 group statistics (ANCE/model/dro_loss.py:217-224): sums[g] = sum of loss_i with g_i == g,
 * counts[g] = #{i: g_i == g} (both overwritten); bwd: dloss_i = dsums[g_i]. */
int cdr_group_reduce_fwd(const float* loss, const int64_t* g, int32_t n, int32_t n_groups, float* sums, float* counts,
                         void* stream);
int cdr_group_reduce_bwd(const float* dsums, const int64_t* g, int32_t n, int32_t n_groups, float* dloss, void* stream);

/* K12 Gram matrix of the per-group gradient rows (ANCE/model/dro_loss.py:235-237 computes
 * normalise + G G^T; the row norms are the diagonal): gram[g,g] += x[g, 0:p] x[g, 0:p]^T, fp32. */
int cdr_gram_f32(const float* x, int32_t g, int64_t p, int64_t ldx, float* gram, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K15 corpus scan: exact inner products of fp16 queries against HBM-resident fp16 doc embeddings
 * (fp32 accumulate on tcgen05) with a fused threshold filter and a per-query sort; output ordered by
 * (score desc, doc index asc).  Replaces faiss IndexFlatIP.add/search
 * (evaluate/evaluation/evaluate_beir.py:220-224, ANCE/drivers/run_ann_data_gen.py:310-317,390).
 * Admission thresholds come from a strided sample of the corpus; status[0] (device int32) counts the
 * queries whose candidate set came out smaller than k or larger than the buffer.  Results are only
 * valid when it is 0; otherwise the caller re-scans in chunks of <= cdr_scan_exhaustive_docs(k)
 * documents (every document admitted: cannot fail) and folds them with cdr_topk_merge.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cdr_scan_args {
  const void* docs;    /* fp16 [n_docs, dim], row stride ld_docs */
  const void* queries; /* fp16 [n_q, dim] contiguous */
  float* out_scores;   /* [n_q, k] */
  int64_t* out_ids;    /* [n_q, k] doc row index + doc_base */
  void* workspace;
  size_t workspace_bytes;
  int32_t* status;     /* device int32[1] */
  int64_t n_docs, ld_docs, doc_base;
  int32_t n_q, dim, k;
  int32_t reserved;
} cdr_scan_args;
size_t cdr_scan_workspace_bytes(int64_t n_docs, int32_t n_q, int32_t k, int32_t dim);
/* largest n_docs for which the scan admits every document (no sampling; cannot under/overflow) */
int64_t cdr_scan_exhaustive_docs(int32_t k);
int cdr_scan_topk(const cdr_scan_args* args, void* stream);
/* merge n_in candidates per query (any order; e.g. all-gathered per-shard top-k lists) into the top k */
int cdr_topk_merge(const float* scores, const int64_t* ids, int32_t n_q, int32_t n_in, int32_t k, float* out_scores,
                   int64_t* out_ids, void* stream);
/* Sharded search (documents split over ranks; replaces the reference's per-rank pickles + numpy merge,
 * ANCE/utils/util.py:87-155).  cdr_topk_pack turns a shard's [n_q, k_in] result into order-preserving u64 keys
 * [n_q * ks + 2] (k_in <= ks, missing columns = empty; the two trailing words carry *status and `all_returned` = "this
 * shard holds no document beyond its list") -- ONE all-gather moves every shard's block.  cdr_topk_merge_keys sorts
 * the world * ks gathered keys of each query straight from that layout and writes the global top k; with ks < k the
 * merge is exact iff every truncated shard's last key is <= the k-th merged key, otherwise (or when a shard scan
 * reported a failure) *flag (device int32, caller-zeroed) is raised and the caller repeats the search with ks = k. */
int cdr_topk_pack(const float* scores, const int64_t* ids, int32_t n_q, int32_t k_in, int32_t ks, const int32_t* status,
                  int32_t all_returned, uint64_t* keys, void* stream);
int cdr_topk_merge_keys(const uint64_t* gathered, int32_t world, int32_t n_q, int32_t ks, int32_t k, float* out_scores,
                        int64_t* out_ids, int32_t* flag, void* stream);

/* ------------------------------------------------------------------------------------------------
 * ANN episode on the device (SURVEY f-2): the steps around the scan in the reference's ANN data generation.
 * cdr_mine_negatives replaces GenerateNegativePassaageID (ANCE/drivers/run_ann_data_gen.py:497-570): per query
 * row of I [n_q, k] (document rows from the scan, -1 = empty): rr = 1 / rank of the positive passage among all k
 * results (0 if absent); neg [n_q, n_neg] = the first n_neg distinct passage ids != positive met while walking the
 * candidates order[q, 0:n_sel] (NULL = 0..n_sel-1, the SelectTopK branch), padded with -1; neg_count [n_q].
 * cdr_kmeans_assign / _accumulate are the two halves of a Lloyd iteration for the query clustering that yields the
 * iDRO group ids (faiss.Kmeans + IndexFlatL2.search(q, 1), :340-351): assign[i] = argmax_c scores[i, c] - half_sq[c]
 * (scores = X C^T from cdr_gemm, ties to the lowest c); sums[g, :] += x[i, :], counts[g] += 1.
 * ---------------------------------------------------------------------------------------------- */
int cdr_mine_negatives(const int64_t* I, int32_t n_q, int32_t k, const int64_t* doc_pid, int64_t n_docs,
                       const int64_t* pos_pid, const int32_t* order, int32_t n_sel, int32_t n_neg, float* rr,
                       int64_t* neg, int32_t* neg_count, void* stream);
int cdr_kmeans_assign(const float* scores, int64_t ld, const float* half_sq, int64_t n, int32_t k, int32_t* assign,
                      void* stream);
int cdr_kmeans_accumulate(const void* x, const int32_t* assign, int64_t n, int32_t dim, int32_t k, float* sums,
                          float* counts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Token-record reader (host only; SURVEY f-3): the reference's EmbeddingCache files (ANCE/utils/util.py:316-370;
 * group variant evaluate/utils/util.py:338-369): fixed-size records [len u32 BE][ids int32 x embedding_size]
 * (or [group u32 BE][len u32 BE][ids ...]).  The file is mapped once; cdr_records_gather copies a batch of
 * records (any order, repeats allowed) into caller-owned -- typically pinned -- buffers as padded int32 ids
 * [n, max_len], byte mask [n, max_len] = [1]*len + [0]*pad (optional), lengths [n] (optional, clipped to max_len)
 * and group ids [n] (optional; -1 without the group header), using n_threads host threads.
 * cdr_records_open returns NULL on failure (text in cdr_last_error()).
 * ---------------------------------------------------------------------------------------------- */
typedef struct cdr_records cdr_records;
cdr_records* cdr_records_open(const char* path, int64_t record_bytes, int64_t total, int32_t group);
void cdr_records_close(cdr_records* handle);
int cdr_records_gather(const cdr_records* handle, const int64_t* keys, int64_t n, int32_t max_len, int32_t* ids,
                       uint8_t* mask, int32_t* lens, int32_t* groups, int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* COCODR_B200_H */
