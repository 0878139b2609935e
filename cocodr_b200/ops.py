"""torch.autograd.Function wrappers that chain the C-ABI kernels into the operators of the hot path.

One Function per BERT *layer* (not per encoder) so that ``output_hidden_states``, the Condenser
``skip_from`` tap and iDRO's re-entrant partial backward (ANCE/model/dro_loss.py:192-204 calls
``torch.autograd.grad(..., retain_graph=True)`` once per group) keep working unmodified.

Numerics: fp16 activations / fp16 weight shadows, fp32 accumulation, fp32 LayerNorm statistics, fp32
parameters, biases and parameter gradients; the CLS embedding leaves the last LayerNorm in fp32
(SURVEY.md H4).  Activation gradients travel between layers as fp16 tensors multiplied by a static
power-of-two ``grad_scale`` (they would underflow otherwise); every parameter gradient is divided by it
again inside the producing kernel (GEMM alpha / LN out_scale), so callers only ever see true gradients.
"""
from collections import namedtuple

import torch

from . import kernels as K

_GRAD_SCALE = 1024.0
FORCE_SHADOW_REFRESH = False  # set while a CUDA graph is being captured (graph.py)
GRAD_SYNC = None              # an active gradsync.GradSync collects the flat gradient buffers of each backward
CLS_PUSH = None               # (peer.PeerExchange, first_seq): the last LayerNorm pushes CLS rows to every rank (peer.py)
GROUP_CAPTURE = None          # a GroupCapture: layer backwards also hand over their wgrad operands (dro_loss.py, K11)
# first parameter of a Function -> forwards since the last GradSync exit (gradsync.py).  Keyed by the parameter OBJECT
# (weakly): keyed by id(), a stale entry of a deleted model could be inherited by a new parameter that happens to get
# the same id -- on some ranks only, which changed the order of GradSync's collectives there (an N = 8 hang).
from torch.utils.weak import WeakIdKeyDictionary as _WeakIdKeyDictionary  # noqa: E402
FWD_CALLS = _WeakIdKeyDictionary()


# Parameter-gradient GEMMs (wgrad) of a layer are independent of the activation-gradient chain.  With WGRAD_OVERLAP they
# are enqueued on a second stream (forked / joined with events, also inside a CUDA graph capture), so the persistent
# one-CTA-per-SM GEMMs of the two streams fill each other's last partial wave (a [16384, 768] output is 192 tiles on 74
# CTA pairs = 2.6 waves).  CDR_WGRAD_OVERLAP=0 keeps everything on one stream.
import os as _os
WGRAD_OVERLAP = _os.environ.get("CDR_WGRAD_OVERLAP", "0") != "0"
_WGRAD_STREAMS = {}


def _wgrad_stream(dev):
    st = _WGRAD_STREAMS.get(dev)
    if st is None:
        st = _WGRAD_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return st


def _wgrad(a, b, out, **kw):
    """dW += a^T b on the wgrad stream, after everything enqueued so far on the current stream."""
    if not WGRAD_OVERLAP:
        return K.gemm(a, b, out, **kw)
    side = _wgrad_stream(a.device)
    side.wait_event(torch.cuda.current_stream().record_event())
    with torch.cuda.stream(side):
        K.gemm(a, b, out, **kw)


def _wgrad_join(dev):
    """The current stream waits for the wgrad stream (end of a layer backward: operands die, gradients are consumed)."""
    if WGRAD_OVERLAP:
        torch.cuda.current_stream().wait_event(_wgrad_stream(dev).record_event())


def _note_forward(ctx, params):
    """Count the forwards of a parameter set that autograd records (grad mode is off INSIDE Function.forward, so the
    Function's needs_input_grad is the signal): GradSync only reduces a layer's flat gradient buffer in place when the
    layer ran exactly once (otherwise autograd accumulates several buffers into one .grad)."""
    if any(ctx.needs_input_grad):
        k = params[0]
        FWD_CALLS[k] = FWD_CALLS.get(k, 0) + 1
    return params


def _grad_flat(ctx, n, dev):
    """Zeroed fp32 [n] buffer the parameter gradients of one Function backward are views of.  Inside a GradSync that
    owns a peer arena (peeropt.PeerArena) it is the Function's persistent slice of the symmetric arena -- provided
    autograd will ADOPT the views as .grad (single forward call, no gradient yet); the optimizer then reads every
    rank's copy over NVLink instead of a collective."""
    gs = GRAD_SYNC
    arena = getattr(gs, "arena", None) if gs is not None else None
    if arena is not None and FWD_CALLS.get(ctx.params[0], 1) == 1 and all(p.grad is None for p in ctx.params):
        return arena.flat(ctx.params[0], n, dev)
    return torch.zeros(n, dtype=torch.float32, device=dev)


def _submit(ctx, flat):
    if GRAD_SYNC is not None:
        GRAD_SYNC.submit(flat, ctx.params, FWD_CALLS.get(ctx.params[0], 1))


# Dropout of one encoder pass: ``state`` = int64 device tensor {seed, offset} (a snapshot taken by the pass, so the
# backward regenerates the forward's masks whatever ran in between), p_hidden / p_attn = HF's hidden_dropout_prob /
# attention_probs_dropout_prob.  Sites: 0 embeddings; layer l: 4l + 1 attention probabilities, 4l + 2 attention-output
# dense, 4l + 3 FFN-output dense (include/cocodr_b200.h cdr_dropout; oracle/dropout_ref.py regenerates the same masks).
DropSpec = namedtuple("DropSpec", "state p_hidden p_attn attn_bits", defaults=(None,))
# attn_bits: optional {layer_index: (keep-bit buffer, event)} -- the attention keep bits of those layers are being
# generated on a second stream (prefill_attn_bits); the layer waits for the event instead of running the generator.

# Generating the keep bits of the attention dropout is 3 M Philox calls per layer (19 us) that depend on nothing but the
# dropout state: the encoder enqueues them for ALL its layers on a second stream at the start of the pass, where they
# share the SMs with the tensor-core GEMMs (whose fp32 / integer pipes are mostly idle) instead of standing between the
# QKV projection and the attention kernel of every layer.  CDR_ATTN_BITS_PREFILL=0 keeps the generator in cdr_attn_fwd.
ATTN_BITS_PREFILL = _os.environ.get("CDR_ATTN_BITS_PREFILL", "1") != "0"
STORE_DENSE_MASKS = _os.environ.get("CDR_STORE_DENSE_MASKS", "1") != "0"  # cdr_dropout.keep_bits for the dense outputs
_BITS_STREAMS = {}


def prefill_attn_bits(drop, n_layers, n_seq, heads, L, dev):
    """-> drop with attn_bits for layers 0 .. n_layers-1 (or drop unchanged when there is nothing to generate)."""
    if drop is None or drop.p_attn <= 0.0 or not ATTN_BITS_PREFILL or L > 128:
        return drop
    st = _BITS_STREAMS.get(dev)
    if st is None:
        st = _BITS_STREAMS[dev] = torch.cuda.Stream(device=dev)
    one = K.attn_dropout_bits(n_seq, heads, L, dev).numel()
    buf = torch.empty(n_layers, one, dtype=torch.uint8, device=dev)  # (allocated on the caller's stream)
    st.wait_event(torch.cuda.current_stream().record_event())       # the dropout-state snapshot exists
    table = {}
    with torch.cuda.stream(st):
        for i in range(n_layers):
            K.attn_fill_bits(buf[i], n_seq=n_seq, seq_len=L, heads=heads, drop=_site(drop, i, 1))
            table[i] = (buf[i], st.record_event())
    return drop._replace(attn_bits=table)


def _attn_bits(drop, layer_index, n_seq, heads, L, dev):
    """(bits, ready) for a layer's attention dropout: the prefilled buffer (after waiting for its event) or a new one."""
    if drop is not None and drop.attn_bits is not None and layer_index in drop.attn_bits:
        bits, ev = drop.attn_bits[layer_index]
        torch.cuda.current_stream().wait_event(ev)
        return bits, True
    return K.attn_dropout_bits(n_seq, heads, L, dev), False


def _site(drop, layer_index, which, row_mul=1, keep_bits=None):
    """cdr_dropout descriptor of site ``which`` (1 attention probs, 2 attention-output dense, 3 FFN-output dense, 0 with
    layer_index 0 = embeddings), or None when that dropout is off.  ``keep_bits``: buffer the dropout GEMM epilogue
    stores its masks in for the LayerNorm backward of the same site."""
    if drop is None:
        return None
    p = drop.p_attn if which == 1 else drop.p_hidden
    if p <= 0.0:
        return None
    return K.drop_args(drop.state, 4 * layer_index + which, p, row_mul, keep_bits)


def _ln_bwd_after_dropout(dy, dcls, y, gamma, mean, rstd, dgamma, dbeta, dbias, *, n_seq, seq_len, S, site, row_ws=None):
    """LayerNorm backward where the LN input was x + dropout(d): -> (dx, dx_drop); dbias += colsum(dx_drop) / S.
    One staged kernel when the fast path applies (fp16 dy only), else LN backward + mask + column sum."""
    rows, H = y.shape
    dev = y.device
    inv = 1.0 / S
    dx = _f16(rows, H, dev=dev)
    if site is None:
        K.ln_bwd(dy, dcls, y, gamma, mean, rstd, dx, dgamma, dbeta, dbias, n_seq=n_seq, seq_len=seq_len, hidden=H,
                 in_scale=S, out_scale=inv, row_ws=row_ws)
        return dx, dx
    dxm = _f16(rows, H, dev=dev)
    if dy is not None and dcls is None and H <= 1024:
        K.ln_bwd_drop(dy, y, gamma, mean, rstd, dx, dxm, dgamma, dbeta, dbias, rows=rows, hidden=H, out_scale=inv,
                      drop=site)
    else:
        K.ln_bwd(dy, dcls, y, gamma, mean, rstd, dx, dgamma, dbeta, None, n_seq=n_seq, seq_len=seq_len, hidden=H,
                 in_scale=S, out_scale=inv, row_ws=row_ws)
        K.dropout_f16(dx, dxm, drop=site)
        K.colsum(dxm, dbias, rows=rows, cols=H, scale=inv)
    return dx, dxm


class GroupCapture:
    """Collects, per encoder-layer backward, the operands of its parameter-gradient reductions (the activation
    gradients and the saved activations), so that iDRO can reduce them per GROUP of samples instead of over the
    whole batch (dro_loss.iDROLoss._get_grad_grouped) -- one shared dgrad pass instead of one backward per group."""

    def __init__(self):
        self.records = []


def set_grad_scale(s: float):
    """Static scale of the internal fp16 activation-gradient domain (power of two)."""
    global _GRAD_SCALE
    _GRAD_SCALE = float(s)


def get_grad_scale() -> float:
    return _GRAD_SCALE


SHADOW_ALLOC = None  # peeropt.PeerArena: operand shadows are allocated from the symmetric arena while this is set


def _f16(*shape, dev):
    return torch.empty(*shape, dtype=torch.float16, device=dev)


def _f32(*shape, dev):
    return torch.empty(*shape, dtype=torch.float32, device=dev)


def _z32(*shape, dev):
    return torch.zeros(*shape, dtype=torch.float32, device=dev)


class LayerShadow:
    """fp16 operand copies of one BertLayer's matrices (QKV packed [3H, H]) + packed fp32 QKV bias.

    The HF-named fp32 ``nn.Parameter``s stay the source of truth (state-dict contract);
    the shadow is rebuilt whenever a parameter's storage or version counter changes -- by its owner
    (``ShadowSet.refresh``: one launch for all layers) or, standalone, by ``refresh``."""

    __slots__ = ("key", "wqkv", "bqkv", "wo", "wi", "wo2", "managed", "table", "table_key", "max_n")

    def __init__(self):
        self.key = None
        self.managed = False
        self.table = self.table_key = None
        self.max_n = 0

    def _alloc(self, wq, wi):
        dev = wq.device
        H, I = wq.shape[0], wi.shape[0]
        if self.key is None or self.wqkv.device != dev or self.wqkv.shape != (3 * H, H):
            if SHADOW_ALLOC is not None:
                f16 = lambda *sh: SHADOW_ALLOC(sh, torch.float16)  # noqa: E731
                f32 = lambda *sh: SHADOW_ALLOC(sh, torch.float32)  # noqa: E731
            else:
                f16 = lambda *sh: _f16(*sh, dev=dev)  # noqa: E731
                f32 = lambda *sh: _f32(*sh, dev=dev)  # noqa: E731
            self.wqkv, self.bqkv = f16(3 * H, H), f32(3 * H)
            self.wo, self.wi, self.wo2 = f16(H, H), f16(I, H), f16(H, I)
            return True
        return False

    def pairs(self, wq, bq, wk, bk, wv, bv, wo, wi, wo2):
        H = wq.shape[0]
        out = []
        for i, (w, b) in enumerate(((wq, bq), (wk, bk), (wv, bv))):
            out.append((w.detach(), self.wqkv[i * H:(i + 1) * H]))
            out.append((b.detach(), self.bqkv[i * H:(i + 1) * H]))
        out += [(wo.detach(), self.wo), (wi.detach(), self.wi), (wo2.detach(), self.wo2)]
        return out

    def refresh(self, wq, bq, wk, bk, wv, bv, wo, wi, wo2):
        if self.managed:
            return self
        srcs = (wq, bq, wk, bk, wv, bv, wo, wi, wo2)
        key = tuple((t.data_ptr(), t._version) for t in srcs)
        if key == self.key and not FORCE_SHADOW_REFRESH:
            return self
        realloc = self._alloc(wq, wi)
        ptr_key = tuple(t.data_ptr() for t in srcs)
        with torch.no_grad():
            if realloc or self.table is None or self.table_key != ptr_key:
                # (host -> device copy of the pointer table: only when a parameter moved, never inside a graph capture
                # that follows warm-up steps)
                self.table, self.max_n = K.cast_table(self.pairs(*srcs), wq.device)
                self.table_key = ptr_key
            K.cast_multi(self.table, self.max_n)
        self.key = key
        return self


class ShadowSet:
    """All LayerShadows of an encoder, refreshed by ONE cast launch per optimizer step."""

    def __init__(self, n_layers):
        self.layers = [LayerShadow() for _ in range(n_layers)]
        for sh in self.layers:
            sh.managed = True
        self.ptr_key = self.ver_key = self.table = None
        self._srcs = None
        self.max_n = 0

    def mark_fresh(self):
        """The shadows were just rewritten by a fused optimizer step (optim.py): re-snapshot the parameter versions
        so the next forward does not cast them again."""
        if self._srcs is not None:
            self.ptr_key = tuple(t.data_ptr() for srcs in self._srcs for t in srcs)
            self.ver_key = tuple(t._version for srcs in self._srcs for t in srcs)

    def pairs(self, per_layer_srcs):
        """[(parameter, shadow view, is_f32)] for every shadowed parameter (allocates the shadows if needed)."""
        self.refresh(per_layer_srcs)
        out = []
        for sh, srcs in zip(self.layers, per_layer_srcs):
            for (src, dst), param in zip(sh.pairs(*srcs), srcs):
                out.append((param, dst, dst.dtype == torch.float32))
        return out

    def refresh(self, per_layer_srcs):
        """per_layer_srcs[i] = (wq, bq, wk, bk, wv, bv, wo, wi, wo2) of layer i."""
        self._srcs = per_layer_srcs
        ptr_key = tuple(t.data_ptr() for srcs in per_layer_srcs for t in srcs)
        ver_key = tuple(t._version for srcs in per_layer_srcs for t in srcs)
        if ptr_key == self.ptr_key and ver_key == self.ver_key and not FORCE_SHADOW_REFRESH:
            return
        with torch.no_grad():
            realloc = False
            for sh, srcs in zip(self.layers, per_layer_srcs):
                realloc |= sh._alloc(srcs[0], srcs[7])
                sh.key = True
            if realloc or ptr_key != self.ptr_key or self.table is None:
                entries = [e for sh, srcs in zip(self.layers, per_layer_srcs) for e in sh.pairs(*srcs)]
                self.table, self.max_n = K.cast_table(entries, per_layer_srcs[0][0].device)
            K.cast_multi(self.table, self.max_n)
        self.ptr_key, self.ver_key = ptr_key, ver_key


class EmbedLN(torch.autograd.Function):
    """K1: LayerNorm(word[ids] + pos[0:L] + type[0]) -> fp16 [n_seq*L, H]  (HF BertEmbeddings)."""

    @staticmethod
    def forward(ctx, ids, word, pos, typ, gamma, beta, eps, drop=None):
        n_seq, L = ids.shape
        H = word.shape[1]
        dev = word.device
        ids = ids.contiguous()
        out = _f16(n_seq * L, H, dev=dev)
        mean, rstd = _f32(n_seq * L, dev=dev), _f32(n_seq * L, dev=dev)
        K.embed_ln_fwd(ids, word, pos, typ, gamma, beta, out, mean, rstd, n_seq=n_seq, seq_len=L, hidden=H,
                       vocab=word.shape[0], eps=eps)
        ctx.drop = _site(drop, 0, 0)
        ctx.drop_state = drop.state if ctx.drop is not None else None  # keeps the device state alive
        if ctx.drop is not None:  # HF BertEmbeddings: dropout(LayerNorm(...))
            K.dropout_f16(out, out, drop=ctx.drop)
        ctx.save_for_backward(ids, word, pos, typ, gamma, mean, rstd)
        ctx.params = _note_forward(ctx, (word, pos, typ, gamma, beta))
        ctx.eps = eps
        ctx.scale = _GRAD_SCALE
        return out

    @staticmethod
    def backward(ctx, dy):
        ids, word, pos, typ, gamma, mean, rstd = ctx.saved_tensors
        n_seq, L = ids.shape
        H = word.shape[1]
        dev = word.device
        sizes = (word.numel(), pos.numel(), typ.numel(), H, H)
        flat = _grad_flat(ctx, sum(sizes), dev)
        views, off = [], 0
        for n in sizes:
            views.append(flat[off:off + n])
            off += n
        dword, dpos, dtyp = views[0].view_as(word), views[1].view_as(pos), views[2].view_as(typ)
        dgamma, dbeta = views[3], views[4]
        dy = dy.contiguous()
        if ctx.drop is not None:
            dy = K.dropout_f16(dy, torch.empty_like(dy), drop=ctx.drop)
        K.embed_ln_bwd(dy, ids, word, pos, typ, gamma, mean, rstd, dword, dpos, dtyp, dgamma, dbeta,
                       n_seq=n_seq, seq_len=L, hidden=H, vocab=word.shape[0], pad_id=0, in_scale=1.0,
                       out_scale=1.0 / ctx.scale)
        _submit(ctx, flat)
        return None, dword, dpos, dtyp, dgamma, dbeta, None, None


class BertLayerFn(torch.autograd.Function):
    """K2-K7: one post-LN transformer layer on fp16 [T, H] activations.

    forward(x, key_bias, <16 HF parameters>, shadow, n_seq, seq_len, heads, eps, emit_cls)
      -> y fp16 [T, H]  (and cls fp32 [n_seq, H] = row 0 of every sequence when emit_cls)
    """

    @staticmethod
    def forward(ctx, x, key_bias, wq, bq, wk, bk, wv, bv, wo, bo, g1, be1, wi, bi, wo2, bo2, g2, be2, shadow, n_seq,
                L, heads, eps, emit_cls, drop=None, layer_index=0):
        T, H = x.shape
        I = wi.shape[0]
        dev = x.device
        sh = shadow.refresh(wq, bq, wk, bk, wv, bv, wo, wi, wo2)
        x = x.contiguous()
        da = _site(drop, layer_index, 1)
        # the masks of the two dense-output dropouts are stored by their GEMM epilogues (one bit per element, 1.5 MB per
        # site at 16 384 x 768) and read back by the LayerNorm backward instead of 3 Philox calls per row and lane
        keep_b = keep_c = None
        if drop is not None and drop.p_hidden > 0.0 and H % 32 == 0 and STORE_DENSE_MASKS:
            keep_b = torch.empty(T * (H // 8), dtype=torch.uint8, device=dev)
            keep_c = torch.empty(T * (H // 8), dtype=torch.uint8, device=dev)
        db, dc = _site(drop, layer_index, 2, keep_bits=keep_b), _site(drop, layer_index, 3, keep_bits=keep_c)
        qkv = _f16(T, 3 * H, dev=dev)
        K.gemm(x, sh.wqkv, qkv, M=T, N=3 * H, K=H, bias=sh.bqkv)
        att = _f16(T, H, dev=dev)
        lse = _f32(n_seq, heads, L, dev=dev)
        bits, bits_ready = _attn_bits(drop, layer_index, n_seq, heads, L, dev) if da is not None else (None, False)
        K.attn_fwd(qkv, key_bias, att, lse, n_seq=n_seq, seq_len=L, heads=heads, drop=da, drop_bits=bits,
                   bits_ready=bits_ready)
        y1 = _f16(T, H, dev=dev)  # x + dropout(attn_out), pre-LayerNorm
        K.gemm(att, sh.wo, y1, M=T, N=H, K=H, bias=bo, aux=x, drop=db,
               epilogue=K.EPI_BIAS_RESIDUAL if db is None else K.EPI_BIAS_DROP_RESIDUAL)
        x1 = _f16(T, H, dev=dev)
        mean1, rstd1 = _f32(T, dev=dev), _f32(T, dev=dev)
        K.ln_fwd(y1, g1, be1, x1, mean1, rstd1, None, n_seq=n_seq, seq_len=L, hidden=H, eps=eps)
        gp, gl = _f16(T, I, dev=dev), _f16(T, I, dev=dev)  # gelu'(z) (saved for the backward), gelu(z)
        K.gemm(x1, sh.wi, gl, M=T, N=I, K=H, bias=bi, epilogue=K.EPI_BIAS_GELU, out2=gp)
        y2 = _f16(T, H, dev=dev)
        K.gemm(gl, sh.wo2, y2, M=T, N=H, K=I, bias=bo2, aux=x1, drop=dc,
               epilogue=K.EPI_BIAS_RESIDUAL if dc is None else K.EPI_BIAS_DROP_RESIDUAL)
        y = _f16(T, H, dev=dev)
        mean2, rstd2 = _f32(T, dev=dev), _f32(T, dev=dev)
        cls = _f32(n_seq, H, dev=dev) if emit_cls else None
        push = CLS_PUSH if emit_cls else None
        K.ln_fwd(y2, g2, be2, y, mean2, rstd2, cls, n_seq=n_seq, seq_len=L, hidden=H, eps=eps, push=push)
        ctx.save_for_backward(x, key_bias, qkv, att, lse, y1, x1, mean1, rstd1, gp, gl, y2, mean2, rstd2, g1, g2)
        ctx.shadow_w = (sh.wqkv, sh.wo, sh.wi, sh.wo2)
        ctx.drop = (da, db, dc, (drop.state if drop is not None else None, bits))
        ctx.meta = (n_seq, L, heads, I, emit_cls, _GRAD_SCALE)
        ctx.params = _note_forward(ctx, (wq, bq, wk, bk, wv, bv, wo, bo, g1, be1, wi, bi, wo2, bo2, g2, be2))
        ctx.param_keys = tuple(id(t) for t in ctx.params)
        ctx.set_materialize_grads(False)
        if emit_cls:
            return y, cls
        return y

    @staticmethod
    def backward(ctx, dy, dcls=None):
        x, key_bias, qkv, att, lse, y1, x1, mean1, rstd1, gp, gl, y2, mean2, rstd2, g1, g2 = ctx.saved_tensors
        wqkv, wo, wi, wo2 = ctx.shadow_w
        n_seq, L, heads, I, emit_cls, S = ctx.meta
        T, H = x.shape
        dev = x.device
        inv = 1.0 / S
        if dy is not None:
            dy = dy.contiguous()
        if dcls is not None:
            dcls = dcls.contiguous().float()
        # ---- one zero-fill for every parameter-gradient accumulator of the layer (views below)
        sizes = (H, H, H, H * I, I, I * H, H, H, H, H * H, 3 * H * H, 3 * H)
        flat = _grad_flat(ctx, sum(sizes), dev)
        views, off = [], 0
        for n in sizes:
            views.append(flat[off:off + n])
            off += n
        dg2, dbe2, dbo2, dwo2, dbi, dwi, dg1, dbe1, dbo, dwo, dwqkv, dbqkv = views
        dwo2, dwi, dwo, dwqkv = dwo2.view(H, I), dwi.view(I, H), dwo.view(H, H), dwqkv.view(3 * H, H)
        row_ws = _f32(2 * T, dev=dev)
        da, db, dc, (_, bits) = ctx.drop
        # ---- output LayerNorm; column sums of its (dropped) dx are the FFN-down bias gradient.  dy2 = gradient of the
        # residual branch, dy2m = dropout'(dy2) = gradient of the dense output (same tensor without dropout)
        dy2, dy2m = _ln_bwd_after_dropout(dy, dcls, y2, g2, mean2, rstd2, dg2, dbe2, dbo2, n_seq=n_seq, seq_len=L, S=S,
                                          site=dc, row_ws=row_ws)
        # ---- FFN down: dZ = (dy2m W2) * gelu'(z) with db1 = colsum(dZ) fused into the epilogue, dW2 = dy2m^T G
        # (every wgrad is enqueued as soon as its operands exist, ahead of the dgrad that shares them)
        _wgrad(dy2m, gl, dwo2, M=H, N=I, K=T, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, split_k=0, alpha=inv)
        dz = _f16(T, I, dev=dev)
        K.gemm(dy2m, wo2, dz, M=T, N=I, K=H, b_major=1, epilogue=K.EPI_DGELU, aux=gp, colsum=dbi, colsum_scale=inv)
        # ---- FFN up: dx1 = dZ W1 + dy2 (residual), dW1 = dZ^T x1
        _wgrad(dz, x1, dwi, M=I, N=H, K=T, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, split_k=0, alpha=inv)
        dx1 = _f16(T, H, dev=dev)
        K.gemm(dz, wi, dx1, M=T, N=H, K=I, b_major=1, epilogue=K.EPI_BIAS_RESIDUAL, aux=dy2)
        # ---- attention-output LayerNorm
        dy1, dy1m = _ln_bwd_after_dropout(dx1, None, y1, g1, mean1, rstd1, dg1, dbe1, dbo, n_seq=n_seq, seq_len=L, S=S,
                                          site=db, row_ws=row_ws)
        # ---- attention output projection
        _wgrad(dy1m, att, dwo, M=H, N=H, K=T, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, split_k=0, alpha=inv)
        datt = _f16(T, H, dev=dev)
        K.gemm(dy1m, wo, datt, M=T, N=H, K=H, b_major=1)
        # ---- attention core
        dqkv = _f16(T, 3 * H, dev=dev)
        fused_db = L <= 128  # the one-tile backward also emits the QKV bias gradient (column sums of dQKV)
        K.attn_bwd(qkv, key_bias, att, lse, datt, dqkv, n_seq=n_seq, seq_len=L, heads=heads,
                   dbias=dbqkv if fused_db else None, dbias_scale=inv, drop=da, drop_bits=bits)
        # ---- QKV projection: dx = dQKV Wqkv + dy1 (residual)
        _wgrad(dqkv, x, dwqkv, M=3 * H, N=H, K=T, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, split_k=0,
               alpha=inv)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _f16(T, H, dev=dev)
            K.gemm(dqkv, wqkv, dx, M=T, N=H, K=3 * H, b_major=1, epilogue=K.EPI_BIAS_RESIDUAL, aux=dy1)
        if not fused_db:
            K.colsum(dqkv, dbqkv, rows=T, cols=3 * H, scale=inv)
        _wgrad_join(dev)
        _submit(ctx, flat)
        if GROUP_CAPTURE is not None:
            GROUP_CAPTURE.records.append(dict(
                keys=ctx.param_keys, rps=L, L=L, n_seq=n_seq, S=S, dy2=dy2m, gl=gl, dz=dz, x1=x1, dy1=dy1m, att=att,
                dx1=dx1, y1=y1, mean1=mean1, rstd1=rstd1, y2=y2, mean2=mean2, rstd2=rstd2, din2=(dy, dcls), dqkv=dqkv,
                x=x))
        return (dx, None, dwqkv[0:H], dbqkv[0:H], dwqkv[H:2 * H], dbqkv[H:2 * H], dwqkv[2 * H:], dbqkv[2 * H:], dwo,
                dbo, dg1, dbe1, dwi, dbi, dwo2, dbo2, dg2, dbe2, None, None, None, None, None, None, None, None)


class BertLastLayerCLSFn(torch.autograd.Function):
    """The LAST encoder layer when only the [CLS] embedding is consumed (``BertDot_NLL_LN.query_emb / body_emb``:
    ``self.bert(...)[0][:, 0]``, ANCE/model/models.py:225-232): every token still feeds K and V, but the attention
    output projection, both LayerNorms and the whole FFN are only needed on the n_seq [CLS] rows -- and so is their
    backward, because d(hidden) is zero everywhere else.  Same kernels, 128x fewer rows for 6 of the layer's 8
    backward GEMMs and 3 of its 4 forward GEMMs; the [CLS] rows of ``x`` / ``att`` are read in place through a row
    stride of L * H (no gather copy).  Results equal ``BertLayerFn(..., emit_cls=True)[1]``.

    forward(x, key_bias, <16 HF parameters>, shadow, n_seq, seq_len, heads, eps) -> cls fp32 [n_seq, H]
    """

    @staticmethod
    def forward(ctx, x, key_bias, wq, bq, wk, bk, wv, bv, wo, bo, g1, be1, wi, bi, wo2, bo2, g2, be2, shadow, n_seq,
                L, heads, eps, drop=None, layer_index=0):
        T, H = x.shape
        I = wi.shape[0]
        dev = x.device
        sh = shadow.refresh(wq, bq, wk, bk, wv, bv, wo, wi, wo2)
        x = x.contiguous()
        # the dense-output masks are those of the full layer: [CLS] row of sequence s is row s * L of the [T, H] tensor
        da = _site(drop, layer_index, 1)
        db, dc = _site(drop, layer_index, 2, row_mul=L), _site(drop, layer_index, 3, row_mul=L)
        qkv = _f16(T, 3 * H, dev=dev)
        K.gemm(x, sh.wqkv, qkv, M=T, N=3 * H, K=H, bias=sh.bqkv)
        att = _f16(T, H, dev=dev)
        lse = _f32(n_seq, heads, L, dev=dev)
        bits, bits_ready = _attn_bits(drop, layer_index, n_seq, heads, L, dev) if da is not None else (None, False)
        K.attn_fwd(qkv, key_bias, att, lse, n_seq=n_seq, seq_len=L, heads=heads, drop=da, drop_bits=bits,
                   bits_ready=bits_ready)
        xc, attc = x.view(n_seq, L, H)[:, 0], att.view(n_seq, L, H)[:, 0]  # [n_seq, H] views, row stride L * H
        y1 = _f16(n_seq, H, dev=dev)
        K.gemm(attc, sh.wo, y1, M=n_seq, N=H, K=H, bias=bo, aux=xc, drop=db,
               epilogue=K.EPI_BIAS_RESIDUAL if db is None else K.EPI_BIAS_DROP_RESIDUAL)
        x1 = _f16(n_seq, H, dev=dev)
        mean1, rstd1 = _f32(n_seq, dev=dev), _f32(n_seq, dev=dev)
        K.ln_fwd(y1, g1, be1, x1, mean1, rstd1, None, n_seq=n_seq, seq_len=1, hidden=H, eps=eps)
        gp, gl = _f16(n_seq, I, dev=dev), _f16(n_seq, I, dev=dev)
        K.gemm(x1, sh.wi, gl, M=n_seq, N=I, K=H, bias=bi, epilogue=K.EPI_BIAS_GELU, out2=gp)
        y2 = _f16(n_seq, H, dev=dev)
        K.gemm(gl, sh.wo2, y2, M=n_seq, N=H, K=I, bias=bo2, aux=x1, drop=dc,
               epilogue=K.EPI_BIAS_RESIDUAL if dc is None else K.EPI_BIAS_DROP_RESIDUAL)
        y = _f16(n_seq, H, dev=dev)
        mean2, rstd2 = _f32(n_seq, dev=dev), _f32(n_seq, dev=dev)
        cls = _f32(n_seq, H, dev=dev)
        push = None
        if CLS_PUSH is not None:  # (exchange, first_seq): rows are sequences here (seq_len = 1)
            push = CLS_PUSH
        K.ln_fwd(y2, g2, be2, y, mean2, rstd2, cls, n_seq=n_seq, seq_len=1, hidden=H, eps=eps, push=push)
        ctx.save_for_backward(x, key_bias, qkv, att, lse, y1, x1, mean1, rstd1, gp, gl, y2, mean2, rstd2, g1, g2)
        ctx.shadow_w = (sh.wqkv, sh.wo, sh.wi, sh.wo2)
        ctx.drop = (da, db, dc, (drop.state if drop is not None else None, bits))
        ctx.meta = (n_seq, L, heads, I, _GRAD_SCALE)
        ctx.params = _note_forward(ctx, (wq, bq, wk, bk, wv, bv, wo, bo, g1, be1, wi, bi, wo2, bo2, g2, be2))
        ctx.param_keys = tuple(id(t) for t in ctx.params)
        return cls

    @staticmethod
    def backward(ctx, dcls):
        x, key_bias, qkv, att, lse, y1, x1, mean1, rstd1, gp, gl, y2, mean2, rstd2, g1, g2 = ctx.saved_tensors
        wqkv, wo, wi, wo2 = ctx.shadow_w
        n_seq, L, heads, I, S = ctx.meta
        T, H = x.shape
        dev = x.device
        inv = 1.0 / S
        dcls = dcls.contiguous().float()
        sizes = (H, H, H, H * I, I, I * H, H, H, H, H * H, 3 * H * H, 3 * H)
        flat = _grad_flat(ctx, sum(sizes), dev)
        views, off = [], 0
        for n in sizes:
            views.append(flat[off:off + n])
            off += n
        dg2, dbe2, dbo2, dwo2, dbi, dwi, dg1, dbe1, dbo, dwo, dwqkv, dbqkv = views
        dwo2, dwi, dwo, dwqkv = dwo2.view(H, I), dwi.view(I, H), dwo.view(H, H), dwqkv.view(3 * H, H)
        da, db, dc, (_, bits) = ctx.drop
        # ---- [CLS] rows only: output LayerNorm, FFN, attention-output LayerNorm and projection
        dy2, dy2m = _ln_bwd_after_dropout(None, dcls, y2, g2, mean2, rstd2, dg2, dbe2, dbo2, n_seq=n_seq, seq_len=1, S=S,
                                          site=dc)
        dz = _f16(n_seq, I, dev=dev)
        K.gemm(dy2m, wo2, dz, M=n_seq, N=I, K=H, b_major=1, epilogue=K.EPI_DGELU, aux=gp, colsum=dbi, colsum_scale=inv)
        K.gemm(dy2m, gl, dwo2, M=H, N=I, K=n_seq, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, split_k=0, alpha=inv)
        dx1 = _f16(n_seq, H, dev=dev)
        K.gemm(dz, wi, dx1, M=n_seq, N=H, K=I, b_major=1, epilogue=K.EPI_BIAS_RESIDUAL, aux=dy2)
        K.gemm(dz, x1, dwi, M=I, N=H, K=n_seq, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, split_k=0, alpha=inv)
        dy1, dy1m = _ln_bwd_after_dropout(dx1, None, y1, g1, mean1, rstd1, dg1, dbe1, dbo, n_seq=n_seq, seq_len=1, S=S,
                                          site=db, row_ws=_f32(2 * n_seq, dev=dev))
        attc = att.view(n_seq, L, H)[:, 0]
        datt = torch.zeros(T, H, dtype=torch.float16, device=dev)  # d(ctx) is zero off the [CLS] rows
        dattc = _f16(n_seq, H, dev=dev)
        K.gemm(dy1m, wo, dattc, M=n_seq, N=H, K=H, b_major=1)
        K.gemm(dy1m, attc, dwo, M=H, N=H, K=n_seq, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, split_k=0, alpha=inv)
        datt.view(n_seq, L, H)[:, 0].copy_(dattc)
        # ---- dense again from here: the [CLS] query attends to every key
        dqkv = _f16(T, 3 * H, dev=dev)
        fused_db = L <= 128
        K.attn_bwd(qkv, key_bias, att, lse, datt, dqkv, n_seq=n_seq, seq_len=L, heads=heads,
                   dbias=dbqkv if fused_db else None, dbias_scale=inv, drop=da, drop_bits=bits)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _f16(T, H, dev=dev)
            # residual branch: only the [CLS] rows carry it.  datt (zero off the [CLS] rows) has been consumed by the
            # attention backward: its [CLS] rows now take dy1 and it serves as the residual operand of the epilogue,
            # so the add happens in fp32 before the single rounding, exactly as in the full layer
            datt.view(n_seq, L, H)[:, 0].copy_(dy1)
            K.gemm(dqkv, wqkv, dx, M=T, N=H, K=3 * H, b_major=1, epilogue=K.EPI_BIAS_RESIDUAL, aux=datt)
        K.gemm(dqkv, x, dwqkv, M=3 * H, N=H, K=T, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, split_k=0,
               alpha=inv)
        if not fused_db:
            K.colsum(dqkv, dbqkv, rows=T, cols=3 * H, scale=inv)
        _submit(ctx, flat)
        if GROUP_CAPTURE is not None:  # rows of the post-attention operands are sequences here (one [CLS] row each)
            GROUP_CAPTURE.records.append(dict(
                keys=ctx.param_keys, rps=1, L=L, n_seq=n_seq, S=S, dy2=dy2m, gl=gl, dz=dz, x1=x1, dy1=dy1m,
                att=attc.contiguous(), dx1=dx1, y1=y1, mean1=mean1, rstd1=rstd1, y2=y2, mean2=mean2, rstd2=rstd2,
                din2=(None, dcls), dqkv=dqkv, x=x))
        return (dx, None, dwqkv[0:H], dbqkv[0:H], dwqkv[H:2 * H], dbqkv[H:2 * H], dwqkv[2 * H:], dbqkv[2 * H:], dwo,
                dbo, dg1, dbe1, dwi, dbi, dwo2, dbo2, dg2, dbe2, None, None, None, None, None, None, None)


class HiddenToFloat(torch.autograd.Function):
    """Boundary between the internal fp16 hidden states (scaled-gradient domain) and caller-visible fp32
    tensors: forward casts, backward multiplies the caller's true gradient by grad_scale."""

    @staticmethod
    def forward(ctx, h, n_seq, L):
        ctx.scale = _GRAD_SCALE
        return h.float().view(n_seq, L, -1)

    @staticmethod
    def backward(ctx, g):
        return (g.reshape(-1, g.shape[-1]) * ctx.scale).to(torch.float16), None, None


class FloatToHidden(torch.autograd.Function):
    """Inverse boundary: caller fp32 [n_seq, L, H] -> internal fp16 [T, H]."""

    @staticmethod
    def forward(ctx, h):
        ctx.scale = _GRAD_SCALE
        ctx.shape = h.shape
        return h.reshape(-1, h.shape[-1]).to(torch.float16)

    @staticmethod
    def backward(ctx, g):
        return (g.float() / ctx.scale).view(ctx.shape)


class PairNLL(torch.autograd.Function):
    """K8 (ANCE/model/models.py:101-108): (loss[B], accs[B] int64, logits[B,2])."""

    @staticmethod
    def forward(ctx, q, a, b):
        q, a, b = q.contiguous().float(), a.contiguous().float(), b.contiguous().float()
        n = q.shape[0]
        dev = q.device
        loss, logits = _f32(n, dev=dev), _f32(n, 2, dev=dev)
        accs = torch.empty(n, dtype=torch.int64, device=dev)
        K.pair_nll_fwd(q, a, b, loss, accs, logits)
        ctx.save_for_backward(q, a, b, logits)
        ctx.mark_non_differentiable(accs, logits)
        return loss, accs, logits

    @staticmethod
    def backward(ctx, dloss, _dacc, _dlogits):
        q, a, b, logits = ctx.saved_tensors
        dq, da, db = torch.empty_like(q), torch.empty_like(a), torch.empty_like(b)
        K.pair_nll_bwd(q, a, b, logits, dloss.contiguous().float(), dq, da, db)
        return dq, da, db


class SimmatCE(torch.autograd.Function):
    """K9 / K9': per-row softmax cross-entropy over the similarity matrix q k^T (fp32), fused: the scores are never
    written to memory (forward: online softmax over key tiles; backward: tiles recomputed from q, k and the row lse)."""

    @staticmethod
    def forward(ctx, q, k, mode, row_offset, loss_scale):
        q, k = q.contiguous().float(), k.contiguous().float()
        n = q.shape[0]
        dev = q.device
        loss, lse = _f32(n, dev=dev), _f32(n, dev=dev)
        K.simmat_ce_fwd(q, k, loss, lse, mode=mode, row_offset=row_offset, loss_scale=loss_scale)
        ctx.save_for_backward(q, k, lse)
        ctx.cfg = (mode, row_offset, loss_scale)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        q, k, lse = ctx.saved_tensors
        mode, row_offset, loss_scale = ctx.cfg
        dq = torch.empty_like(q) if ctx.needs_input_grad[0] else None
        dk = torch.empty_like(k) if ctx.needs_input_grad[1] else None
        K.simmat_ce_bwd(q, k, lse, dloss.contiguous().float(), dq, dk, mode=mode, row_offset=row_offset,
                        loss_scale=loss_scale)
        return dq, dk, None, None, None


class GroupStats(torch.autograd.Function):
    """K10 (ANCE/model/dro_loss.py:217-224): scatter-add of per-sample losses / ones by group id."""

    @staticmethod
    def forward(ctx, loss, g, n_groups):
        loss, g = loss.contiguous().float(), g.contiguous().long()
        dev = loss.device
        sums, cnts = _f32(n_groups, dev=dev), _f32(n_groups, dev=dev)
        K.group_reduce_fwd(loss, g, sums, cnts, n_groups=n_groups)
        ctx.save_for_backward(g)
        ctx.n_groups = n_groups
        ctx.mark_non_differentiable(cnts)
        return sums, cnts

    @staticmethod
    def backward(ctx, dsums, _dc):
        (g,) = ctx.saved_tensors
        dloss = _f32(g.numel(), dev=g.device)
        K.group_reduce_bwd(dsums.contiguous().float(), g, dloss, n_groups=ctx.n_groups)
        return dloss, None, None


class OwnPairCE(torch.autograd.Function):
    """Per-sample in-batch InfoNCE values ``CE(q_i K^T, row_offset + i)`` whose autograd graph reaches only the sample's
    OWN pair: dq_i = sum_j dS_ij k_j over all keys, dp_i = dS_i,own q_i, every other key a constant.  This is the
    loss view iDRO takes its group gradients from when the keys are in-batch passages (each sample's loss then depends
    on encoder sequences {i, i + B} only, so the grouped wgrad K11 applies); ``p_own`` must equal
    ``keys[row_offset : row_offset + B]``."""

    @staticmethod
    def forward(ctx, q, p_own, keys, row_offset):
        q, keys = q.contiguous().float(), keys.detach().contiguous().float()
        n = q.shape[0]
        dev = q.device
        loss, lse = _f32(n, dev=dev), _f32(n, dev=dev)
        K.simmat_ce_fwd(q, keys, loss, lse, mode=K.SIM_QP, row_offset=row_offset, loss_scale=1.0)
        ctx.save_for_backward(q, keys, lse, loss)
        ctx.row_offset = row_offset
        return loss

    @staticmethod
    def backward(ctx, dloss):
        q, keys, lse, loss = ctx.saved_tensors
        dloss = dloss.contiguous().float()
        dq = torch.empty_like(q)
        K.simmat_ce_bwd(q, keys, lse, dloss, dq, None, mode=K.SIM_QP, row_offset=ctx.row_offset, loss_scale=1.0)
        dp = torch.empty_like(q)
        K.own_key_grad(q, loss, dloss, dp)  # dS_i,own = dloss_i * (softmax_i,own - 1) = dloss_i * (exp(-loss_i) - 1)
        return dq, dp, None, None


def pair_nll(q, a, b):
    return PairNLL.apply(q, a, b)


def coco_contrastive(E, row_offset=0, n_rows=None, loss_scale=1.0):
    """loss_i = loss_scale * CE(S[i,:], i^1), S = E_rows E^T with the diagonal masked (COCO/modeling.py:244-248)."""
    q = E if n_rows is None else E[row_offset:row_offset + n_rows]
    return SimmatCE.apply(q, E, K.SIM_COCO, row_offset, float(loss_scale))


def qp_infonce(Q, P_all, row_offset=0):
    """In-batch q x p InfoNCE over (all-gathered) passages: loss_i = CE(Q_i P_all^T, row_offset + i)."""
    return SimmatCE.apply(Q, P_all, K.SIM_QP, row_offset, 1.0)


def group_stats(loss, g, n_groups):
    return GroupStats.apply(loss, g, n_groups)


class MLMShadow:
    """fp16 operands of the MLM head: transform weight [H,H]; decoder (= word embedding) matrix zero-padded to a
    multiple of 64 rows [Vp,H]; fp32 decoder bias padded with -inf so padding columns vanish in the softmax."""

    def __init__(self):
        self.key = None
        self.table = self.table_key = None
        self.max_n = 0

    def refresh(self, wt, emb, bv):
        key = tuple((t.data_ptr(), t._version) for t in (wt, emb, bv))
        if key == self.key and not FORCE_SHADOW_REFRESH:
            return self
        dev = wt.device
        V, H = emb.shape
        vp = (V + 63) // 64 * 64
        realloc = self.key is None or self.emb.shape != (vp, H) or self.emb.device != dev
        if realloc:
            self.vp = vp
            self.wt = _f16(H, H, dev=dev)
            self.emb = torch.zeros(vp, H, dtype=torch.float16, device=dev)
            self.bias = torch.full((vp,), float("-inf"), dtype=torch.float32, device=dev)
        ptr_key = tuple(t.data_ptr() for t in (wt, emb, bv))
        with torch.no_grad():
            if realloc or self.table is None or self.table_key != ptr_key:  # pointer table: rebuilt only when a parameter moved
                self.table, self.max_n = K.cast_table([(wt.detach(), self.wt), (emb.detach(), self.emb[:V]),
                                                       (bv.detach(), self.bias[:V])], dev)
                self.table_key = ptr_key
            K.cast_multi(self.table, self.max_n)
        self.key = key
        return self


class MLMHead(torch.autograd.Function):
    """K14: per-row MLM cross-entropy of HF BertOnlyMLMHead on GATHERED masked rows.

    loss_i = CE(LN(gelu(x_i Wt^T + bt)) E^T + b_v, label_i);  x fp16 [M, H] (internal hidden rows), labels [M]."""

    @staticmethod
    def forward(ctx, x, labels, wt, bt, gamma, beta, emb, bv, shadow, eps):
        M, H = x.shape
        V = emb.shape[0]
        dev = x.device
        sh = shadow.refresh(wt, emb, bv)
        x = x.contiguous()
        labels = labels.contiguous()
        gp, t = _f16(M, H, dev=dev), _f16(M, H, dev=dev)  # gelu'(z), gelu(z)
        K.gemm(x, sh.wt, t, M=M, N=H, K=H, bias=bt, epilogue=K.EPI_BIAS_GELU, out2=gp)
        u = _f16(M, H, dev=dev)
        mean, rstd = _f32(M, dev=dev), _f32(M, dev=dev)
        K.ln_fwd(t, gamma, beta, u, mean, rstd, None, n_seq=M, seq_len=1, hidden=H, eps=eps)
        logits = _f32(M, sh.vp, dev=dev)
        K.gemm(u, sh.emb, logits, M=M, N=sh.vp, K=H, epilogue=K.EPI_F32_STORE)
        loss, lse = _f32(M, dev=dev), _f32(M, dev=dev)
        K.vocab_ce_fwd(logits, sh.bias, labels, loss, lse, n_cols=sh.vp)
        ctx.save_for_backward(x, labels, gp, t, u, mean, rstd, logits, lse, gamma)
        ctx.sh = (sh.wt, sh.emb, sh.bias, sh.vp, V)
        ctx.scale = _GRAD_SCALE
        return loss

    @staticmethod
    def backward(ctx, dloss):
        x, labels, gp, t, u, mean, rstd, logits, lse, gamma = ctx.saved_tensors
        wt16, emb16, bias, vp, V = ctx.sh
        M, H = x.shape
        dev = x.device
        S = ctx.scale
        inv = 1.0 / S
        dlog = _f16(M, vp, dev=dev)
        K.vocab_ce_bwd(logits, bias, labels, lse, dloss.contiguous().float(), dlog, n_cols=vp, scale=S)
        sizes = (vp * H, vp, H, H, H * H, H)
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        views, off = [], 0
        for n in sizes:
            views.append(flat[off:off + n])
            off += n
        dE, dbv, dgamma, dbeta, dwt, dbt = views
        dE, dwt = dE.view(vp, H), dwt.view(H, H)
        du = _f16(M, H, dev=dev)
        K.gemm(dlog, emb16, du, M=M, N=H, K=vp, b_major=1)
        K.gemm(dlog, u, dE, M=vp, N=H, K=M, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, split_k=0, alpha=inv)
        K.colsum(dlog, dbv, rows=M, cols=vp, scale=inv)
        dt = _f16(M, H, dev=dev)
        K.ln_bwd(du, None, t, gamma, mean, rstd, dt, dgamma, dbeta, None, n_seq=M, seq_len=1, hidden=H, out_scale=inv,
                 row_ws=_f32(2 * M, dev=dev))
        dz = _f16(M, H, dev=dev)
        K.dgelu(dt, gp, dz)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _f16(M, H, dev=dev)
            K.gemm(dz, wt16, dx, M=M, N=H, K=H, b_major=1)
        K.gemm(dz, x, dwt, M=H, N=H, K=M, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, split_k=0, alpha=inv)
        K.colsum(dz, dbt, rows=M, cols=H, scale=inv)
        return dx, None, dwt, dbt, dgamma, dbeta, dE[:V], dbv[:V], None, None


class GatherRows(torch.autograd.Function):
    """rows idx of an internal fp16 [T, H] tensor (the ~15 % masked positions the MLM loss looks at)."""

    @staticmethod
    def forward(ctx, h, idx):
        ctx.save_for_backward(idx)
        ctx.rows = h.shape[0]
        return h.index_select(0, idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        out = torch.zeros(ctx.rows, g.shape[1], dtype=g.dtype, device=g.device)
        out.index_add_(0, idx, g)  # (add, not copy: fixed-capacity gathers repeat row 0 in their unused slots)
        return out, None
