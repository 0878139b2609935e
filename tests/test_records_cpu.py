"""CPU: token-record reader (cdr_records_*, cocodr_b200.records.EmbeddingCache) against the oracle restatement AND
the unmodified reference class (ANCE/utils/util.py EmbeddingCache, imported when /root/reference is present)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest


def _make(tmp_path, group, n=257, L=32, seed=0):
    from oracle import records_ref
    rng = np.random.RandomState(seed)
    toks = [list(rng.randint(1, 30000, size=rng.randint(0, L + 9))) for _ in range(n)]  # empty, ragged, over-long
    groups = list(rng.randint(0, 50, size=n)) if group else None
    path = os.path.join(tmp_path, "cache_g" if group else "cache")
    records_ref.write_cache(path, toks, L, groups)
    return path, toks, groups


@pytest.mark.parametrize("group", [False, True])
def test_single_records_match_oracle(tmp_path, group):
    from cocodr_b200 import records
    from oracle import records_ref
    path, toks, groups = _make(str(tmp_path), group)
    with records.EmbeddingCache(path, group=group) as c:
        assert len(c) == len(toks) and c.record_size == 32 * 4 + (8 if group else 4)
        for key in (0, 1, 100, len(toks) - 1):
            got, ref = c[key], records_ref.read_record(path, key, 32, group)
            assert got[:-1] == ref[:-1] and np.array_equal(got[-1], ref[-1])
        with pytest.raises(IndexError):
            c[len(toks) + 1]
        with pytest.raises(IndexError):
            c[-1]
        c._pos = 0
        first = c.read_single_record()  # sequential reads advance like the reference's file cursor
        second = c.read_single_record()
        assert first[0] == records_ref.read_record(path, 0, 32, group)[-2]
        assert np.array_equal(second[1], records_ref.read_record(path, 1, 32, group)[-1])


@pytest.mark.parametrize("group,max_len,threads", [(False, 32, 1), (True, 32, 4), (False, 16, 3), (True, 48, 8)])
def test_gather_batches(tmp_path, group, max_len, threads):
    from cocodr_b200 import records
    from oracle import records_ref
    path, toks, groups = _make(str(tmp_path), group, n=1000)
    keys = np.random.RandomState(1).randint(0, 1000, size=777)  # repeats, any order
    with records.EmbeddingCache(path, group=group) as c:
        b = c.gather(keys, max_len=max_len, pin=False, threads=threads)
        assert b["ids"].shape == (777, max_len) and b["mask"].dtype.is_floating_point is False
        for i, k in enumerate(keys):
            rec = records_ref.read_record(path, int(k), 32, group)
            ids, mask = records_ref.processing_fn(rec[-2], rec[-1], max_len)
            assert np.array_equal(b["ids"][i].numpy(), ids)
            assert np.array_equal(b["mask"][i].numpy(), mask)
            assert int(b["lens"][i]) == min(rec[-2], max_len)
            assert int(b["groups"][i]) == (rec[0] if group else -1)
        with pytest.raises(RuntimeError):
            c.gather([1000])
        assert c.gather([], pin=False)["ids"].shape == (0, 32)


def test_errors_are_loud(tmp_path):
    from cocodr_b200 import records
    path, _, _ = _make(str(tmp_path), False, n=10)
    with open(path, "r+b") as f:
        f.truncate(100)  # shorter than total_number records
    with pytest.raises(RuntimeError):
        records.EmbeddingCache(path).open()
    with pytest.raises(RuntimeError):
        records.EmbeddingCache(path).gather([0])  # not open


@pytest.mark.skipif(not os.path.exists("/root/reference/ANCE/utils/util.py"), reason="reference tree not present")
def test_matches_the_unmodified_reference_class(tmp_path):
    """Pins the oracle and the native reader: the reference's own EmbeddingCache reads the same files identically."""
    for name in ("pytrec_eval", "faiss", "tensorboardX", "transformers.optimization"):  # absent imports of util.py
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference/ANCE")  # util.py imports the sibling `model` package
    spec = importlib.util.spec_from_file_location("ref_util", "/root/reference/ANCE/utils/util.py")
    ref_util = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(ref_util)
    except Exception as e:  # the module has further unrelated imports
        pytest.skip(f"reference util.py does not import here: {e}")
    from cocodr_b200 import records
    from oracle import records_ref
    path, toks, _ = _make(str(tmp_path), False, n=64)
    with ref_util.EmbeddingCache(path) as ref, records.EmbeddingCache(path) as ours:
        assert len(ref) == len(ours) and ref.record_size == ours.record_size
        for key in (0, 5, 63):
            a, b, o = ref[key], ours[key], records_ref.read_record(path, key, 32)
            assert a[0] == b[0] == o[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[1], o[1])
    seeded_ref, seeded = ref_util.EmbeddingCache(path, seed=3), records.EmbeddingCache(path, seed=3)
    assert np.array_equal(seeded_ref.ix_array, seeded.ix_array)


@pytest.mark.skipif(not os.path.exists("/root/reference/ANCE/data/msmarco_data.py"), reason="reference tree not present")
def test_gather_equals_the_reference_processing_fn(tmp_path):
    """ids / attention mask of a gathered batch == what the UNMODIFIED GetProcessingFn (ANCE/data/msmarco_data.py:297-325,
    its AST node compiled on its own: the module's other imports are absent here) builds record by record."""
    import ast
    import types as _types
    import torch
    from torch.utils.data import TensorDataset
    from cocodr_b200 import records
    path_ref = "/root/reference/ANCE/data/msmarco_data.py"
    tree = ast.parse(open(path_ref).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "GetProcessingFn"][0]
    ns = {"torch": torch, "TensorDataset": TensorDataset}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path_ref, "exec"), ns)
    path, toks, _ = _make(str(tmp_path), False, n=50, L=32)
    for query, max_len in ((True, 32), (False, 32)):
        args = _types.SimpleNamespace(max_query_length=max_len, max_seq_length=max_len)
        ref_fn = ns["GetProcessingFn"](args, query=query)
        with records.EmbeddingCache(path) as c:
            b = c.gather(list(range(50)), max_len=max_len, pin=False)
            for i in range(50):
                ids_ref, mask_ref, _tt, idx_ref = ref_fn(c[i], i)[0]
                assert torch.equal(b["ids"][i], ids_ref) and torch.equal(b["mask"][i], mask_ref) and int(idx_ref) == i
