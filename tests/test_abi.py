"""CPU tests of the C ABI: the shared library loads, exports every symbol the header declares, its
host-only entry points agree with the oracle, and it refuses to run without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from scasa_b200 import lib
    lib.build()
    return lib.load()


def test_header_and_library_agree(L):
    from scasa_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "scasa_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)              # declarations only, not the prose
    declared = set(re.findall(r"\b(scasa_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    for sym in declared:
        assert hasattr(L, sym), f"libscasa_cuda.so does not export {sym}"
    assert b"sm_100a" in L.scasa_version()


def test_no_cpu_fallback(L):
    """Without a CUDA device the library must fail loudly, not compute on the host."""
    if L.scasa_device_count() > 0:
        pytest.skip("a GPU is present")
    from scasa_b200 import lib
    from scasa_b200.xmatrix import XMatrix
    from scasa_b200.api import Scasa
    xm = XMatrix.load_npz(os.path.join(ROOT, "tests", "golden", "xmatrix_hg38_alevin.npz"))
    with pytest.raises(lib.ScasaError) as e:
        Scasa(xm)
    assert e.value.code == -2 and "no CUDA device" in str(e.value)


def test_opts_default_are_the_reference_defaults(L):
    from scasa_b200 import lib
    o = lib.default_opts()
    # estim_chunk(Y, DM0, maxiter.X=1000, maxerr=0.01, lim=0.01, maxiter.bt=1000, modify=TRUE)  Rsource:944
    assert (o.maxiter_x, o.maxiter_bt, o.maxerr, o.lim, o.modify) == (1000, 1000, 0.01, 0.01, 1)
    assert (o.adjust_esp, o.hamming_dist, o.fixed_iter, o.fp32) == (2.0, 1, 0, 0)   # Rsource:973, :534
    assert (o.maxerr_bt, o.lim_bt) == (0.01, 0.01)      # estim.BETA's own defaults (Rsource:726-727): estim_chunk does not forward its maxerr / lim (:1016)


def test_chunk_split_equals_oracle(L, oracle):
    from scasa_b200.api import chunk_split
    for n in list(range(1, 420)) + [999, 1000, 1049, 1050, 1051, 3955, 10000, 100000, 123457]:
        for core in (1, 4, 8):
            if n <= 100 and core > n:
                continue                                           # reference misbehaves (repeated breaks); out of scope
            assert np.array_equal(chunk_split(n, core), oracle.chunk_split(n, core)), (n, core)
    rc = L.scasa_chunk_split(C.c_int64(0), 4, None, 0, None)
    assert rc == -1                                                # SCASA_ERR_ARG, no crash


def test_shard_chunks_partitions_contiguously():
    from scasa_b200.shard import shard_chunks, local_problem
    from scasa_b200.api import chunk_split
    chunks = chunk_split(3955, 4)
    for world in (1, 2, 3, 8, 40, 64):
        parts = shard_chunks(chunks, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == len(chunks) - 1
        assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
        sizes = [chunks[b] - chunks[a] for a, b in parts]
        if world <= 40:
            assert max(sizes) - min(sizes) <= 100 and min(sizes) > 0  # balanced to within about one chunk
    cell_ptr = np.arange(3956) * 7
    c0, c1, loc, e0, e1 = local_problem(chunks, cell_ptr, shard_chunks(chunks, 4)[2])
    assert loc[0] == 0 and loc[-1] == c1 - c0 and (e0, e1) == (c0 * 7, c1 * 7)


def _expand_case(seed, n_rows, n_cols, offset):
    import ctypes as C
    from scasa_b200 import lib as _l
    L = _l.load()
    rng = np.random.default_rng(seed)
    dense = rng.normal(size=(n_rows, n_cols)) * (rng.random((n_rows, n_cols)) < 0.08)
    dense[0] = 0                                   # an empty row
    dense[1] = rng.normal(size=n_cols)             # a full row
    dense[2, 0] = -0.0
    dense[3, -1] = np.nan
    mask = dense.view(np.int64) != 0               # bit pattern, like the device kernel
    nnz = mask.sum(1).astype(np.int32)
    start = np.concatenate([[0], np.cumsum(nnz)[:-1]]).astype(np.int64)
    rows, cols = np.nonzero(mask)
    vals = dense[rows, cols]
    buf = np.full(n_rows * n_cols + offset + 8, 7.0)     # offset: 8-byte aligned but not 64-byte aligned rows
    out = buf[offset:offset + n_rows * n_cols]
    rc = L.scasa_expand_pairs(_l.ptr(out, C.c_double), n_rows, n_cols, _l.ptr(start, C.c_int64), _l.ptr(nnz, C.c_int32),
                              _l.ptr(cols.astype(np.int32), C.c_int32), _l.ptr(np.ascontiguousarray(vals), C.c_double), 3)
    assert rc == 0
    assert np.array_equal(out.view(np.int64), dense.reshape(-1).view(np.int64))
    assert buf[offset - 1] == 7.0 if offset else True
    assert buf[offset + n_rows * n_cols] == 7.0        # nothing written past the end


def test_expand_pairs_all_code_paths():
    """Host half of the zero-suppressed fetch: bit-exact dense rows, ragged widths and alignments,
    on the widest instruction set of this machine and on the forced AVX2 / SSE2 paths."""
    import subprocess, sys
    for n_cols, off in ((1, 0), (7, 1), (64, 0), (129, 3), (1000, 5)):
        _expand_case(n_cols, 9, n_cols, off)
    code = ("import sys; sys.path.insert(0, %r); from tests.test_abi import _expand_case\n"
            "for n_cols, off in ((1, 0), (7, 1), (64, 0), (129, 3), (1000, 5)): _expand_case(n_cols, 9, n_cols, off)" % ROOT)
    for path in ("sse2", "avx2"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SCASA_EXPAND=path), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
