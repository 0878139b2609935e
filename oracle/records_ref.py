"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): pure-Python restatement of the reference's token-record cache.

``write_cache`` follows the writers (ANCE/data/msmarco_data.py:66-95: each record is ``len.to_bytes(4, 'big') +
np.array(padded_ids, np.int32).tobytes()`` -- the 8-byte id prefix of the split files is dropped on merge; group
variant prefixes ``group.to_bytes(4, 'big')``) and ``pad_input_ids`` (ANCE/utils/util.py:158-172);
``read_record`` follows ``EmbeddingCache.read_single_record[_with_group]`` (ANCE/utils/util.py:338-343,
evaluate/utils/util.py:359-369); ``processing_fn`` follows ``GetProcessingFn`` (ANCE/data/msmarco_data.py:297-305).
Pinned against the unmodified reference class in tests/test_records_cpu.py (it imports in this container).
"""
import json

import numpy as np


def pad_input_ids(ids, max_length, pad_token=0):
    ids = list(ids)
    pad = max_length - len(ids)
    return ids[:max_length] if pad <= 0 else ids + [pad_token] * pad


def write_cache(path, token_lists, max_length, groups=None):
    with open(path, "wb") as f:
        for i, toks in enumerate(token_lists):
            plen = min(len(toks), max_length)
            rec = plen.to_bytes(4, "big") + np.array(pad_input_ids(toks, max_length), np.int32).tobytes()
            if groups is not None:
                rec = int(groups[i]).to_bytes(4, "big") + rec
            f.write(rec)
    with open(path + "_meta", "w") as f:
        json.dump({"type": "int32", "total_number": len(token_lists), "embedding_size": max_length}, f)


def read_record(path, key, embedding_size, group=False):
    record_size = embedding_size * 4 + (8 if group else 4)
    with open(path, "rb") as f:
        f.seek(key * record_size)
        b = f.read(record_size)
    if group:
        return int.from_bytes(b[:4], "big"), int.from_bytes(b[4:8], "big"), np.frombuffer(b[8:], dtype=np.int32)
    return int.from_bytes(b[:4], "big"), np.frombuffer(b[4:], dtype=np.int32)


def processing_fn(passage_len, passage, max_len):
    """ids / attention mask of GetProcessingFn for one record (passage is already padded to the record width)."""
    plen = min(passage_len, max_len)
    ids = np.zeros(max_len, dtype=np.int32)
    n = min(len(passage), max_len)
    ids[:n] = passage[:n]
    mask = np.array([1] * plen + [0] * (max_len - plen), dtype=bool)
    return ids, mask
