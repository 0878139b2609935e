"""Token-record cache: host-side mirror of the reference's ``EmbeddingCache`` (ANCE/utils/util.py:316-370, group
variant evaluate/utils/util.py:329-391) over the native reader in csrc/records.cu.

Same constructor, context-manager protocol, ``__getitem__`` / ``__iter__`` / ``__len__`` / ``read_single_record``
results (``(passage_len, ids)`` or ``(group_id, passage_len, ids)``), so ``GetProcessingFn`` and
``GetTrainingDataProcessingFn`` (ANCE/data/msmarco_data.py:297-420) keep working unchanged -- plus ``gather``: a whole
batch of records copied by a few native threads straight into pinned host tensors shaped for the encoder
(int32 ids ``[n, L]``, bool mask, lengths, group ids), which replaces 3 Python seek+read calls per triplet.
"""
import ctypes as C
import json
import os

import numpy as np
import torch

from . import _lib
from ._lib import check


class EmbeddingCache:
    def __init__(self, base_path, group=False, seed=-1):
        self.base_path = base_path
        self.group = bool(group)
        with open(base_path + "_meta", "r") as f:
            meta = json.load(f)
        self.dtype = np.dtype(meta["type"])
        if self.dtype != np.int32:
            raise RuntimeError(f"EmbeddingCache: only int32 token records are supported (meta says {meta['type']})")
        self.total_number = int(meta["total_number"])
        self.embedding_size = int(meta["embedding_size"])
        self.record_size = self.embedding_size * self.dtype.itemsize + (8 if self.group else 4)
        if seed >= 0:
            self.ix_array = np.random.RandomState(seed).permutation(self.total_number)
        else:
            self.ix_array = np.arange(self.total_number)
        self._h = None
        self._pos = 0  # record cursor of read_single_record (the reference's file position)

    # ---- lifecycle (reference: open / close / __enter__ / __exit__) -------------------------------
    def open(self):
        lib = _lib.load()
        lib.cdr_records_open.restype = C.c_void_p
        h = lib.cdr_records_open(os.fsencode(self.base_path), C.c_int64(self.record_size),
                                 C.c_int64(self.total_number), C.c_int32(int(self.group)))
        if not h:
            raise RuntimeError("cocodr_b200 cdr_records_open failed: " + lib.cdr_last_error().decode("utf-8", "replace"))
        self._h = C.c_void_p(h)
        self._pos = 0

    def close(self):
        if self._h is not None:
            lib = _lib.load()
            lib.cdr_records_close.restype = None
            lib.cdr_records_close(self._h)
            self._h = None

    def __enter__(self):
        self.open()
        return self

    def __exit__(self, type, value, traceback):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- batched native path ------------------------------------------------------------------------
    def gather(self, keys, max_len=None, pin=True, threads=8):
        """Records ``keys`` (any order, repeats allowed) -> dict(ids int32 [n, L], mask bool [n, L], lens int32 [n],
        groups int32 [n] (-1 without the group header)); host tensors, pinned when a GPU is present."""
        if self._h is None:
            raise RuntimeError("EmbeddingCache is not open")
        L = int(max_len) if max_len is not None else self.embedding_size
        keys = np.ascontiguousarray(np.asarray(keys, dtype=np.int64).reshape(-1))
        n = len(keys)
        pin = bool(pin) and torch.cuda.is_available()
        ids = torch.empty((n, L), dtype=torch.int32, pin_memory=pin)
        mask = torch.empty((n, L), dtype=torch.uint8, pin_memory=pin)
        lens = torch.empty((n,), dtype=torch.int32, pin_memory=pin)
        groups = torch.empty((n,), dtype=torch.int32, pin_memory=pin)
        if n > 0:
            check(_lib.load().cdr_records_gather(self._h, C.c_void_p(keys.ctypes.data), C.c_int64(n), C.c_int32(L),
                                                 C.c_void_p(ids.data_ptr()), C.c_void_p(mask.data_ptr()),
                                                 C.c_void_p(lens.data_ptr()), C.c_void_p(groups.data_ptr()),
                                                 C.c_int32(threads)), "cdr_records_gather")
        return {"ids": ids, "mask": mask.view(torch.bool), "lens": lens, "groups": groups}

    # ---- reference-shaped single-record access ------------------------------------------------------
    def _record(self, key):
        b = self.gather([key], pin=False, threads=1)
        ids = b["ids"][0].numpy()
        plen = self._raw_len(key)
        if self.group:
            return int(b["groups"][0]), plen, ids
        return plen, ids

    def _raw_len(self, key):
        # passage_len exactly as stored (the native gather clips it to max_len for the mask)
        with open(self.base_path, "rb") as f:
            f.seek(key * self.record_size + (4 if self.group else 0))
            return int.from_bytes(f.read(4), "big")

    def read_single_record(self):
        rec = self._record(self._pos)
        self._pos += 1
        return rec[-2:] if self.group else rec

    def read_single_record_with_group(self):
        rec = self._record(self._pos)
        self._pos += 1
        return rec

    def __getitem__(self, key):
        if key < 0 or key > self.total_number:
            raise IndexError("Index {} is out of bound for cached embeddings of size {}".format(key, self.total_number))
        if key == self.total_number:  # the reference's off-by-one bound lets this through to an empty read
            raise IndexError("Index {} is out of bound for cached embeddings of size {}".format(key, self.total_number))
        self._pos = key + 1
        return self._record(key)

    def __iter__(self):
        for i in range(self.total_number):
            yield self.__getitem__(int(self.ix_array[i]))

    def __len__(self):
        return self.total_number
