"""CPU: the C-ABI shared library loads, exports every symbol include/cocodr_b200.h declares, its argument
structs have the layout the ctypes binding assumes, and host-only entry points answer without a GPU."""
import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    from cocodr_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from cocodr_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 20 and "cdr_gemm" in names and "cdr_attn_bwd" in names and "cdr_scan_topk" in names
    for n in names:
        assert hasattr(lib, n), n
    assert lib.cdr_version() >= 100


def test_struct_layouts_match_the_header(lib, tmp_path):
    from cocodr_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "cocodr_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(cdr_gemm_args),sizeof(cdr_attn_args),sizeof(cdr_simmat_args),sizeof(cdr_scan_args));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.GemmArgs), C.sizeof(_lib.AttnArgs), C.sizeof(_lib.SimmatArgs), C.sizeof(_lib.ScanArgs)]


def test_host_only_entry_points(lib):
    from cocodr_b200 import kernels as K
    assert K.scan_exhaustive_docs(100) == 8192 and K.scan_exhaustive_docs(1000) == 8192
    assert K.scan_exhaustive_docs(2000) == 16384
    small = K.scan_workspace_bytes(6000, 37, 100)
    big = K.scan_workspace_bytes(1_000_000, 1000, 1000)
    assert 0 < small < big
    assert big >= 1000 * 8192 * 8 + 1000 * 8192 * 4
    assert K.scan_workspace_bytes(0, 10, 10) == 0


def test_bad_arguments_fail_loudly_without_a_gpu(lib):
    from cocodr_b200 import _lib
    rc = lib.cdr_gemm(None, None)
    assert rc == -1 and b"null" in lib.cdr_last_error()
    a = _lib.AttnArgs()
    a.qkv, a.n_seq, a.seq_len, a.heads, a.head_dim = 16, 1, 600, 1, 64
    assert lib.cdr_attn_fwd(C.byref(a), None) == -1 and b"seq_len" in lib.cdr_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "cdr_gemm")
    # the row-segmented / grouped wgrad entry points validate before they touch the device
    g = _lib.GemmArgs()
    g.a, g.b, g.out, g.M, g.N, g.K, g.lda, g.ldb, g.ldo = 16, 16, 16, 128, 128, 64, 128, 128, 128
    g.a_major, g.b_major, g.epilogue = 1, 0, _lib.EPI_F32_ATOMIC
    one = (C.c_int64 * 1)(0)
    assert lib.cdr_gemm_segments(C.byref(g), C.c_int32(1), one, one, one, None) == -1
    assert b"MN-major" in lib.cdr_last_error()
    assert lib.cdr_gemm_grouped(C.byref(g), C.c_int32(4), C.c_void_p(16), C.c_int64(128 * 128), None) == -1
    assert b"MN-major" in lib.cdr_last_error()
    g.b_major, g.epilogue = 1, _lib.EPI_STORE_F16
    assert lib.cdr_gemm_segments(C.byref(g), C.c_int32(1), one, one, one, None) == -1 and b"fp32" in lib.cdr_last_error()
    assert lib.cdr_gemm_grouped(C.byref(g), C.c_int32(4), C.c_void_p(16), C.c_int64(128 * 128), None) == -1
    g.epilogue = 8  # CDR_EPI_F32_GROUPED is internal to cdr_gemm_grouped
    assert lib.cdr_gemm(C.byref(g), None) == -1 and b"internal" in lib.cdr_last_error()
    g.epilogue = _lib.EPI_F32_ATOMIC
    assert lib.cdr_gemm_grouped(C.byref(g), C.c_int32(0), C.c_void_p(16), C.c_int64(128 * 128), None) == -1
    assert lib.cdr_gemm_grouped(C.byref(g), C.c_int32(4), C.c_void_p(16), C.c_int64(130), None) == -1  # unaligned stride
    assert lib.cdr_gemm_segments(C.byref(g), C.c_int32(0), None, None, None, None) == 0  # nothing to do
    # the dropout epilogue refuses to run without a counter state / with a degenerate threshold
    g.a_major, g.b_major, g.epilogue, g.aux, g.ldaux = 0, 0, _lib.EPI_BIAS_DROP_RESIDUAL, 16, 128
    assert lib.cdr_gemm(C.byref(g), None) == -1 and b"drop" in lib.cdr_last_error()
    d = _lib.Dropout()
    d.state, d.threshold = 16, 70000
    assert lib.cdr_dropout_f16(C.c_void_p(16), C.c_void_p(16), C.c_int64(4), C.c_int32(64), C.byref(d), None) == -1
    assert lib.cdr_dropout_f16(C.c_void_p(16), C.c_void_p(16), C.c_int64(4), C.c_int32(60), C.byref(d), None) == -1
    d.threshold = 6554
    assert lib.cdr_ln_bwd_drop(C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), C.c_void_p(16),
                               C.c_void_p(16), C.c_void_p(16), None, None, None, C.c_int32(8), C.c_int32(2048),
                               C.c_float(1.0), C.byref(d), None) == -1


def test_round2_entry_points_validate_without_a_gpu(lib):
    """cdr_attn_dropout_bits_fill and the peer-memory optimizer entry points reject bad arguments before touching the
    device (no dropout configured / null tables / an epoch pointer that is not the peers' one)."""
    from cocodr_b200 import _lib, optim, peer
    a = _lib.AttnArgs()
    a.n_seq, a.seq_len, a.heads, a.head_dim, a.drop_bits = 2, 32, 2, 64, 32
    assert lib.cdr_attn_dropout_bits_fill(C.byref(a), None) == -1 and b"dropout" in lib.cdr_last_error()
    assert lib.cdr_attn_dropout_bits_fill(None, None) == -1
    a.drop.state, a.drop.threshold, a.drop_bits = 16, 6554, 24  # misaligned bit buffer
    assert lib.cdr_attn_dropout_bits_fill(C.byref(a), None) == -1 and b"aligned" in lib.cdr_last_error()
    o = optim.OptArgs()
    pa = peer.PeerArgs()
    assert lib.cdr_adam_multi_peer(C.byref(o), C.byref(pa), None, None, None) == -1  # empty table
    o.items, o.chunks, o.count, o.n_chunks, o.lr, o.step, o.mode = 16, 16, 1, 1, 16, 16, 0
    assert lib.cdr_adam_multi_peer(C.byref(o), C.byref(pa), None, None, None) == -1 and b"world" in lib.cdr_last_error()
    pa.world, pa.rank, pa.epoch, pa.done_counter = 2, 0, 64, 16
    for r in range(2):
        pa.peer_buf[r], pa.peer_flag[r] = 4096 * (r + 1), 256 * (r + 1)
    assert lib.cdr_adam_multi_peer(C.byref(o), C.byref(pa), C.c_void_p(128), None, None) == -1
    assert b"epoch" in lib.cdr_last_error()
    assert lib.cdr_grad_reduce_clip_peer(C.byref(o), C.byref(pa), C.c_void_p(64), C.c_float(1.0), None, None, None, None,
                                         None) == -1 and b"scratch" in lib.cdr_last_error()
    assert lib.cdr_adam_multi_peer_reduced(C.byref(o), None, C.c_void_p(64), None, None) == -1


def test_ctypes_structs_match_the_c_header(tmp_path):
    """Every args struct crossing the C ABI: sizeof and every field offset of the ctypes mirror == what gcc lays out
    from include/cocodr_b200.h (a silent mismatch would shift pointers, not fail)."""
    import ctypes as C
    import os
    import shutil
    import subprocess
    from cocodr_b200 import _lib, optim, peer
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = {"cdr_gemm_args": _lib.GemmArgs, "cdr_attn_args": _lib.AttnArgs, "cdr_simmat_args": _lib.SimmatArgs,
             "cdr_scan_args": _lib.ScanArgs, "cdr_opt_item": optim.OptItem, "cdr_opt_chunk": optim.OptChunk,
             "cdr_opt_args": optim.OptArgs, "cdr_peer_args": peer.PeerArgs, "cdr_dropout": _lib.Dropout}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{_lib.HEADER}"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src, exe = os.path.join(tmp_path, "abi.c"), os.path.join(tmp_path, "abi")
    with open(src, "w") as f:
        f.write("\n".join(lines))
    subprocess.run(["gcc", "-std=c11", "-o", exe, src], check=True, capture_output=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.strip().splitlines():
        cname, field, val = line.split()
        cls = pairs[cname]
        want = C.sizeof(cls) if field == "size" else getattr(cls, field).offset
        assert int(val) == want, f"{cname}.{field}: C {val} vs ctypes {want}"
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in pairs.values())
