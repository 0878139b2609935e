"""ctypes binding of libcocodr_b200.so (the C ABI in include/cocodr_b200.h).

Fails loudly: a missing library or a missing symbol raises -- there is no fallback path.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# COCODR_B200_LIB: development override used to A/B differently-compiled builds of the same library
LIB_PATH = os.environ.get("COCODR_B200_LIB") or os.path.join(_HERE, "libcocodr_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "cocodr_b200.h")

_lib = None

# epilogue enum (keep in sync with include/cocodr_b200.h)
EPI_STORE_F16, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_DGELU, EPI_F32_ATOMIC, EPI_F32_STORE, EPI_SCAN_FILTER = range(7)
EPI_BIAS_DROP_RESIDUAL = 9
CDR_EOVERFLOW = -5


class Dropout(C.Structure):
    """cdr_dropout: counter-based dropout descriptor (state = device {seed, offset})."""
    _fields_ = [("state", C.c_void_p), ("site", C.c_uint32), ("threshold", C.c_uint32), ("scale", C.c_float),
                ("row_mul", C.c_int32), ("keep_bits", C.c_void_p)]


class AttnArgs(C.Structure):
    _fields_ = [("qkv", C.c_void_p), ("key_bias", C.c_void_p), ("out", C.c_void_p), ("lse", C.c_void_p),
                ("d_out", C.c_void_p), ("dqkv", C.c_void_p), ("dq_workspace", C.c_void_p),
                ("n_seq", C.c_int32), ("seq_len", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
                ("scale", C.c_float), ("dbias_scale", C.c_float), ("dbias_qkv", C.c_void_p), ("drop", Dropout),
                ("drop_bits", C.c_void_p), ("drop_bits_ready", C.c_int32), ("reserved", C.c_int32)]


class SimmatArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("scores", C.c_void_p), ("gmat", C.c_void_p),
                ("loss", C.c_void_p), ("lse", C.c_void_p), ("dloss", C.c_void_p), ("dq", C.c_void_p),
                ("dk", C.c_void_p),
                ("n_rows", C.c_int32), ("n_keys", C.c_int32), ("dim", C.c_int32), ("mode", C.c_int32),
                ("row_offset", C.c_int32), ("loss_scale", C.c_float)]


class ScanArgs(C.Structure):
    _fields_ = [("docs", C.c_void_p), ("queries", C.c_void_p), ("out_scores", C.c_void_p), ("out_ids", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("status", C.c_void_p),
                ("n_docs", C.c_int64), ("ld_docs", C.c_int64), ("doc_base", C.c_int64),
                ("n_q", C.c_int32), ("dim", C.c_int32), ("k", C.c_int32), ("reserved", C.c_int32)]


class GemmArgs(C.Structure):
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("out", C.c_void_p), ("out2", C.c_void_p),
                ("bias", C.c_void_p), ("aux", C.c_void_p),
                ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
                ("lda", C.c_int64), ("ldb", C.c_int64), ("ldo", C.c_int64), ("ldaux", C.c_int64),
                ("a_major", C.c_int32), ("b_major", C.c_int32), ("epilogue", C.c_int32), ("split_k", C.c_int32),
                ("alpha", C.c_float), ("dbg_lbo", C.c_int32), ("dbg_sbo", C.c_int32),
                ("colsum", C.c_void_p), ("colsum_scale", C.c_float), ("reserved", C.c_int32), ("drop", Dropout)]


def declared_symbols():
    """Every function the public header declares (used by the CPU export test)."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cdr_[a-z0-9_]+)\s*\(", txt)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "cocodr_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    if missing:
        raise RuntimeError(f"libcocodr_b200.so does not export: {missing}")
    lib.cdr_last_error.restype = C.c_char_p
    lib.cdr_version.restype = C.c_int
    for name in declared_symbols():
        fn = getattr(lib, name)
        if name.endswith("_bytes"):
            fn.restype = C.c_size_t
        elif name == "cdr_scan_exhaustive_docs":
            fn.restype = C.c_int64
        elif name != "cdr_last_error":
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().cdr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"cocodr_b200 {what} failed (code {rc}): {msg}")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())
